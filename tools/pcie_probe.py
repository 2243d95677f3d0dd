"""Host<->device copy rates of this box with page-locked memory (what bounds the
e2e line of bench.py): H2D of one 4096^2 u8 map, D2H of J f32 + action u8, each
alone and both directions at once, on every visible GPU concurrently.
Usage: python tools/pcie_probe.py [n_gpus]"""
import sys
import time

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
cells = 4096 * 4096
bufs = []
for d in range(n):
    torch.cuda.set_device(d)
    bufs.append(dict(
        hm=torch.empty(cells, dtype=torch.uint8).pin_memory(),
        dm=torch.empty(cells, dtype=torch.uint8, device=f"cuda:{d}"),
        hj=torch.empty(cells * 5, dtype=torch.uint8).pin_memory(),
        dj=torch.empty(cells * 5, dtype=torch.uint8, device=f"cuda:{d}"),
        s_up=torch.cuda.Stream(device=d), s_dn=torch.cuda.Stream(device=d)))


def run(up, down, reps=10):
    def once():
        for d, b in enumerate(bufs):
            if up:
                with torch.cuda.stream(b["s_up"]):
                    b["dm"].copy_(b["hm"], non_blocking=True)
            if down:
                with torch.cuda.stream(b["s_dn"]):
                    b["hj"].copy_(b["dj"], non_blocking=True)
    def sync():
        for d in range(n):
            torch.cuda.synchronize(d)
    once(); sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    sync()
    dt = (time.perf_counter() - t0) / reps
    nbytes = n * ((cells if up else 0) + (cells * 5 if down else 0))
    return dt * 1e3, nbytes / dt / 1e9


for name, up, down in (("H2D 16 MiB", True, False), ("D2H 80 MiB", False, True),
                       ("both", True, True)):
    ms, gbs = run(up, down)
    print(f"{n} GPU(s) {name:12s} {ms:8.3f} ms per round  {gbs:7.1f} GB/s aggregate "
          f"({gbs / n:6.1f} per GPU)")
