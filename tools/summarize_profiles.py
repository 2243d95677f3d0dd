"""Turn the raw ncu outputs of tools/gpu_final.sh (gpurun_out/, scratch) into the
committed summaries under profiles/.  usage: python tools/summarize_profiles.py <gpurun tag> [round prefix]
Every header names the gpurun call the numbers of THAT file come from."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r02"   # prefix of the files under profiles/


def launch_summary(src, dst, header):
    rows = list(csv.reader(open(src)))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if d["Metric Unit"] == "ns" else v * 1e3 if d["Metric Unit"] == "ms" else v
        k = d["Kernel Name"][:70]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(header)
        f.write("# %-70s %8s %10s %8s %7s\n" % ("kernel", "launches", "total_us", "avg_us", "share"))
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-72s %8d %10.1f %8.2f %6.2f%%\n" % (k, v[0], v[1], v[1] / v[0], v[1] / tot * 100))
    return agg


def raw_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return dict(zip(rows[0], rows[-1])), dict(zip(rows[0], rows[1]))


WANT = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg",
        "sm__cycles_active.avg"]


def kernel_summary(rep, dst, header):
    vals, units = raw_metrics(rep)
    with open(dst, "w") as f:
        f.write(header)
        f.write("%-86s %s\n" % ("Kernel Name", vals.get("Kernel Name", "?")))
        for w in WANT:
            if w in vals:
                f.write("%-86s %s %s\n" % (w, vals[w], units.get(w, "")))
        f.write("# warp stall reasons (warps per issue-active cycle)\n")
        st = [(k, float(v.replace(",", ""))) for k, v in vals.items()
              if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
              and "not_issued" not in k]
        for k, v in sorted(st, key=lambda x: -x[1])[:11]:
            f.write("%-86s %.3f\n" % (k, v))
    return vals, units


def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


launch_summary(os.path.join(OUT, f"launches_{tag}.csv"), os.path.join(PROF, f"{rnd}_launches_summary.txt"),
               "# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 on: python bench.py --steps 2 "
               "--warmup 3 --no-cpu --no-ref-cuda --no-qv\n# (B200, gpurun call %s; first 400 launches; "
               "per-launch times are cold-cache and serialised: compare SHARES)\n" % tag)
shutil.copy(os.path.join(OUT, f"launches_{tag}.csv"), os.path.join(PROF, f"{rnd}_launches.csv"))
launch_summary(os.path.join(OUT, f"pomdp_launches_{tag}.csv"),
               os.path.join(PROF, f"{rnd}_pomdp_launches_summary.txt"),
               "# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pomdp_ -c 600 on: python "
               "tools/bench_pomdp.py 1250 --fixture\n# (QV-tree batch of 1250 plans x 3 calls: 32-query warm-up + two "
               "full batches; the offline solvers are skipped with --fixture: they are 29 000 launches; gpurun "
               "call %s)\n" % tag)
vals, units = kernel_summary(os.path.join(OUT, f"prof_fused_{tag}.ncu-rep"),
                             os.path.join(PROF, f"{rnd}_ncu_fused_kernel.txt"),
                             "# ncu --set full --clock-control none --import-source on, fused kernel "
                             "mdp_sweep_kernel<2,2,false>, 4096x4096, B200 (gpurun call %s)\n# full report: not "
                             "committed (binary); regenerate with tools/gpu_final.sh\n" % tag)
rd = to_bytes(vals["dram__bytes_read.sum"], units["dram__bytes_read.sum"])
wr = to_bytes(vals["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
json.dump({"kernel": "mdp_sweep_kernel<2,2,false>", "dram_bytes_per_launch": rd + wr, "dram_read": rd,
           "dram_write": wr, "source": f"ncu --set full, gpurun call {tag}, 4096x4096, 2 sweeps per launch",
           "algorithmic_bytes_per_launch": 4096 * 4096 * 2 * 10},
          open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
kernel_summary(os.path.join(OUT, f"prof_values_{tag}.ncu-rep"),
               os.path.join(PROF, f"{rnd}_ncu_values_kernel.txt"),
               "# ncu --set full --clock-control none --import-source on, pomdp_values_kernel (QV-tree bounds), "
               "one expansion round of a 1250-query batch, B200 (gpurun call %s)\n" % tag)
for src, dst in ((f"bench_{tag}.json", f"{rnd}_bench_n1.json"),
                 (f"bench_ref_{tag}.json", f"{rnd}_bench_reference_arm.json")):
    shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
print(open(os.path.join(PROF, f"{rnd}_launches_summary.txt")).read())
print(open(os.path.join(PROF, f"{rnd}_pomdp_launches_summary.txt")).read())
print(open(os.path.join(PROF, f"{rnd}_ncu_fused_kernel.txt")).read())
print(open(os.path.join(PROF, f"{rnd}_ncu_values_kernel.txt")).read())
print(open(os.path.join(PROF, "traffic.json")).read())
