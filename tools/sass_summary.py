"""Instruction histogram of the fused sweep kernel in libpp2d.so and the SASS of one
steady-state marching step (cuobjdump; runs anywhere, no GPU needed).
usage: python tools/sass_summary.py > profiles/r02_sass_fused.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("PP2D_LIB") or os.path.join(ROOT, "path_planning_2d_b200", "libpp2d.so")
FUN = "_ZN4pp2d16mdp_sweep_kernelILi2ELi2ELb0ELb0ELb0EEEvNS_11SweepParamsE"
txt = subprocess.run(["cuobjdump", "-sass", "-fun", FUN, LIB], capture_output=True, text=True).stdout
ins = []
for line in txt.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((m.group(1), m.group(2).strip()))


def opcode(t):
    t = re.sub(r"^@!?U?P[0-9T]\s+", "", t)
    return t.split()[0].split(".")[0]


hist = collections.Counter(opcode(t) for _, t in ins)
print(f"# cuobjdump -sass of mdp_sweep_kernel<T=2, CW=2, POLICY=0, P2P=0, LIN=0> in {os.path.basename(LIB)}")
print(f"# {len(ins)} instructions in the kernel")
for k, v in hist.most_common():
    print(f"{v:6d}  {k}")
# tcgen05 / TMA evidence (none expected: gather-and-reduce stencil)
for pat in ("UTMALDG", "UTMASTG", "UTCMMA", "UTCHMMA", "LDTM", "STTM", "LDGSTS", "FFMA", "FMNMX3"):
    print(f"# {pat}: {sum(1 for _, t in ins if pat in t)}")
stores = [i for i, (_, t) in enumerate(ins) if "STG" in t]
if len(stores) >= 4:
    a, b = stores[2], stores[3]
    step = ins[a + 1:b + 1]
    h2 = collections.Counter(opcode(t) for _, t in step)
    print(f"\n# one steady-state marching step (between two J stores): {len(step)} instructions = "
          "4 Bellman backups (2 cells x 2 fused sweeps)")
    print("# " + "  ".join(f"{k} {v}" for k, v in h2.most_common()))
    print("# non-FFMA/FMNMX3 instructions of the step, in order:")
    for addr, t in step:
        if not (t.startswith("FFMA") or t.startswith("FMNMX3")):
            print(f"  /*{addr}*/  {t}")
