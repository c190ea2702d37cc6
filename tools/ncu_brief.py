"""Key raw metrics + per-function / per-opcode instruction shares of one kernel in an .ncu-rep (needs ncu on PATH).
usage: python tools/ncu_brief.py report.ncu-rep [source.cu [kernel-name-regex]]"""
import csv, io, re, subprocess, sys
from collections import Counter

rep = sys.argv[1]
srcfile = sys.argv[2] if len(sys.argv) > 2 else "molearn_b200/csrc/energy.cu"
ksel = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
raw = subprocess.run(["ncu", "-i", rep, *ksel, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block", "sm__warps_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum"]
for k in want:
    print(f"{k:75s} {d.get(k)}")
for k in rows[0]:
    if "issue_stalled" in k and "per_issue_active" in k and float(d[k] or 0) > 0.1:
        print(f"{k.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', ''):75s} {float(d[k]):.2f}")
out = subprocess.run(["ncu", "-i", rep, *ksel, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
src = open(srcfile).read().split("\n")
def region(l):
    for k in range(l, 0, -1):
        t = src[k - 1]
        if "__device__" in t or "__global__" in t or re.match(r"^(bonded_kernel|nb_kernel)", t):
            m = re.search(r"(\w+)\s*\(", t)
            if m: return m.group(1)
    return "?"
hdr = None; cur = None; infile = False
reg = Counter(); ops = Counter(); lines = Counter()
for r in rows:
    if r and r[0] == "File Path": infile = r[1].endswith(srcfile.split("/")[-1]); continue
    if r and r[0] == "Line No": hdr = {h: i for i, h in reversed(list(enumerate(r)))}; continue
    if hdr is None or len(r) < 10: continue
    if r[0]:
        try: cur = int(r[0])
        except ValueError: pass
        continue
    try: n = float(r[hdr["Instructions Executed"]])
    except ValueError: n = 0.0
    sass = r[3].split()
    if sass:
        ops[(sass[1] if sass[0].startswith("@") else sass[0]).split(".")[0]] += n
    if infile and cur:
        reg[region(cur)] += n; lines[cur] += n
    else:
        reg["<headers>"] += n
tot = sum(ops.values())
print("warp instructions", tot)
print("by function:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in reg.most_common(12)))
print("by opcode:  ", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(16)))
for l, c in lines.most_common(14):
    print(f"  {l:5d} {100 * c / tot:4.1f}%  {src[l - 1].strip()[:120]}")
