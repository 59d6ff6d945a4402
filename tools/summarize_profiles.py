#!/usr/bin/env python
"""Turns gpurun_out ncu artefacts into the committed text summaries under profiles/.
usage: python tools/summarize_profiles.py launches <csv> <out.txt> "<title>" [--last N]   (N = launches of one step)
       python tools/summarize_profiles.py kernel <ncu-rep> <out.txt> "<title>" """
import collections
import csv
import re
import subprocess
import sys

mode, src, dst, title = sys.argv[1:5]
if mode == "launches":
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
    if "--last" in sys.argv:
        rows = rows[-int(sys.argv[sys.argv.index("--last") + 1]):]
    for row in rows:
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    out = [f"# {title}", "# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes",
           f"total_ms {tot:.3f}  launches {sum(n for n, _ in agg.values())}"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{t:9.3f} ms {100 * t / tot:5.1f}%  n={n:4d}  avg={t / n * 1000:9.1f} us  {k}")
else:
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_elapsed.avg",
            "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"]
    out = [f"# {title}"]
    for h, u, v in zip(r[0], r[1], r[2]):
        if h in want:
            out.append(f"{h:78s} {u:16s} {v}")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
