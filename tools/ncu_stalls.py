"""Warp-stall samples of one kernel in an .ncu-rep (captured with --import-source on), aggregated by opcode.
usage: python tools/ncu_stalls.py report.ncu-rep kernel-name-regex"""
import csv, io, subprocess, sys
from collections import defaultdict

rep, kre = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + kre, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def f(r, h):
    try:
        return float(r[ix[h]])
    except (ValueError, IndexError):
        return 0.0


seen, tot = set(), 0.0
byop, cnt, ex = defaultdict(lambda: defaultdict(float)), defaultdict(float), defaultdict(float)
for r in data:
    if len(r) < len(hdr) or r[ix["Address"]] in seen:      # several launches of the kernel repeat the listing
        continue
    seen.add(r[ix["Address"]])
    w = r[ix["Source"]].split()
    op = (w[1] if w and w[0].startswith("@") else w[0] if w else "?").split(".")[0]
    n = f(r, "# Samples")
    tot += n
    cnt[op] += n
    ex[op] += f(r, "Instructions Executed")
    for s in stalls:
        byop[op][s] += f(r, s)
print(f"total samples {tot:.0f}")
print(f"{'op':8s} {'samples':>8s} {'share':>6s} {'executed':>10s}  top stall reasons")
for op, c in sorted(cnt.items(), key=lambda x: -x[1])[:14]:
    top = sorted(byop[op].items(), key=lambda x: -x[1])[:4]
    print(f"{op:8s} {c:8.0f} {100 * c / tot:5.1f}% {ex[op]:10.0f}  " + ", ".join(f"{k[6:]} {v:.0f}" for k, v in top))
