#!/usr/bin/env python
"""Turn gpurun_out/<tag>/{launches.csv,prof.ncu-rep,bench*.json} into the small text summaries kept under profiles/.

    python scripts/summarize_profiles.py <tag> [<round-name>]
"""
import collections
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1]
name = sys.argv[2] if len(sys.argv) > 2 else tag
src = os.path.join("gpurun_out", tag)
dst = "profiles"
os.makedirs(dst, exist_ok=True)

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
           "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
           "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct"]

lines = []
lp = os.path.join(src, "launches.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, cnt, tot, probes = collections.OrderedDict(), collections.Counter(), 0.0, {}
    for r in rows[hi + 2:]:
        if len(r) <= vi:
            continue
        k = r[ki][:100]
        v = float(r[vi].replace(",", ""))
        if "cutlass" in k or "dmma_peak_kernel" in k or "cublas" in k.lower():      # FP64 ceiling probes: outside the timed passes
            probes[k] = probes.get(k, 0.0) + v
            continue
        agg[k] = agg.get(k, 0.0) + v
        cnt[k] += 1
        tot += v
    lines.append("# ncu launch list (gpu__time_duration.sum, --clock-control none): `python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline`")
    lines.append("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  total %.1f us over %d launches (3 passes)" % (tot / 1e3, sum(cnt.values())))
    lines.append("%10s %6s %8s %6s  kernel" % ("total_us", "count", "us/launch", "share"))
    for k, v in sorted(agg.items(), key=lambda x: -x[1]):
        lines.append("%10.1f %6d %8.1f %5.1f%%  %s" % (v / 1e3, cnt[k], v / 1e3 / cnt[k], 100 * v / tot, k))
    for k, v in probes.items():
        lines.append("# not counted (FP64 ceiling probe after the timed passes): %10.1f us  %s" % (v / 1e3, k))
    open(os.path.join(dst, "%s_launches.txt" % name), "w").write("\n".join(lines) + "\n")

for rep, suffix in (("prof", "ncu_full"), ("prof_eig", "ncu_eig")):
    rp = os.path.join(src, rep + ".ncu-rep")
    if not os.path.exists(rp):
        continue
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = ["# ncu --set full --clock-control none, selected raw metrics per captured launch (%s)" % tag]
    for r in rows[2:]:
        out.append("")
        out.append("kernel: " + r[idx["Kernel Name"]][:160])
        for m in METRICS:
            if m in idx:
                out.append("  %-78s %s %s" % (m, r[idx[m]], units[idx[m]]))
    open(os.path.join(dst, "%s_%s.txt" % (name, suffix)), "w").write("\n".join(out) + "\n")

for f in sorted(os.listdir(src)):
    if f.startswith("bench") and f.endswith(".json"):
        try:
            d = json.loads(open(os.path.join(src, f)).read().strip().splitlines()[-1])
            json.dump(d, open(os.path.join(dst, "%s_%s" % (name, f)), "w"), indent=1)
        except Exception:
            pass
print("wrote", [f for f in os.listdir(dst) if f.startswith(name)])
