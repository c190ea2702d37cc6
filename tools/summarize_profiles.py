"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

usage: python tools/summarize_profiles.py <tag> <launches.csv> <report.ncu-rep>
"""
import collections
import csv
import io
import re
import subprocess
import sys

tag, launches, report = sys.argv[1:4]
out = io.StringIO()

rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
iK, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[iK]).replace("void ", "")
    agg[name].append(float(r[iV].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in agg.values())
out.write(f"# launch list ({launches}): gpu__time_duration.sum per launch, cold-cache and serialised - compare SHARES\n")
out.write(f"{'kernel':58s} {'launches':>8s} {'median_us':>10s} {'total_us':>10s} {'share':>7s}\n")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    v2 = sorted(v)
    out.write(f"{k[:58]:58s} {len(v):8d} {v2[len(v2) // 2]:10.2f} {sum(v):10.1f} {100 * sum(v) / tot:6.1f}%\n")

raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, u = rr[0], rr[1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_global_red.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
out.write(f"\n# full capture ({report}): one launch per kernel, --set full --clock-control none\n")
for row in rr[2:]:
    d = dict(zip(h, row))
    kname = d["Kernel Name"].split("(")[0]
    out.write(f"\n## {kname}\n")
    for k in keys:
        if k in d:
            out.write(f"{k:70s} {d[k]:>18s} {u[h.index(k)]}\n")
open(f"profiles/{tag}_summary.txt", "w").write(out.getvalue())
print(out.getvalue())
