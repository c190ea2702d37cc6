"""profiles/r02_summary.json from the .ncu-rep files a gpurun call brought back: the counters bench.py quotes next to
its live timings (DRAM bytes per launch, FMA-pipe and issue-slot utilisation), with the commit the capture was taken at.
usage: python tools/make_profile_summary.py <commit> key=report.ncu-rep[:kernel-regex] ..."""
import csv, io, json, subprocess, sys

commit = sys.argv[1]
out = {"commit": commit, "how": "ncu --set full --clock-control none, one launch per kernel (tools/ncu_round2.sh); cold-cache, serialised"}
want = {"time_us": "gpu__time_duration.sum", "registers": "launch__registers_per_thread",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "fma_pipe_busy_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "warp_instructions": "smsp__inst_executed.sum", "warps_active_per_sm": "sm__warps_active.avg.per_cycle_active"}
for arg in sys.argv[2:]:
    key, rest = arg.split("=", 1)
    rep, _, rx = rest.partition(":")
    cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"] + (["--kernel-name", "regex:" + rx] if rx else [])
    rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
    h, u, r = rows[0], rows[1], rows[2]
    d, units = dict(zip(h, r)), dict(zip(h, u))
    def num(k):
        return float(d[k].replace(",", "")) if d.get(k) not in (None, "") else None
    def nbytes(k):
        v = num(k)
        if v is None:
            return None
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[k], 1)
    e = {"kernel": d["Kernel Name"].split("(")[0], "report": rep.split("/")[-1]}
    for k, m in want.items():
        e[k] = num(m)
    if units.get(want["time_us"]) == "ms":
        e["time_us"] *= 1e3
    e["dram_bytes"] = (nbytes("dram__bytes_read.sum") or 0) + (nbytes("dram__bytes_write.sum") or 0)
    out[key] = e
json.dump(out, open("profiles/r02_summary.json", "w"), indent=1)
print(json.dumps(out, indent=1))
