"""SASS of the pair kernel's mask-free micro-tile (one rotation step) with its instruction counts per pair.
usage: python tools/sass_hotloop.py [lib.so] > profiles/rNN_nb_kernel2_hotloop.sass"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from sass_reads import analyse, blocks_of  # noqa: E402

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "molearn_b200", "_C", "libmolearn_b200.so")
blocks = blocks_of(lib, "nb_kernel2ILb1ELb0ELb1E")


def packed(b):
    return sum(t.startswith(("FFMA2", "FMUL2", "FADD2")) for _, t in b)


hot = min((b for b in blocks if packed(b) >= 150), key=len)      # the masked variant of the same step carries 48 more selects
while hot and hot[0][1].startswith("CS2R"):                      # zeroing of the j-force accumulators in front of the loop
    hot = hot[1:]
cnt = collections.Counter(t.split()[0].split(".")[0] for _, t in hot)
pk = cnt["FFMA2"] + cnt["FMUL2"] + cnt["FADD2"]
n, hist, cyc, _ = analyse(hot)
print("# nb_kernel2<GRAD=1, FULL=0, LUT=1> (sm_100a, cuobjdump -sass of molearn_b200/_C/libmolearn_b200.so, tools/sass_hotloop.py):")
print("# the mask-free micro-tile of one rotation step = 16 pairs per lane = 8 packed iterations (two j-atoms per instruction).")
print(f"# opcode counts of this basic block ({len(hot)} instructions): " + ", ".join(f"{k} {v}" for k, v in cnt.most_common()))
print(f"# per pair: {pk * 2 / 16:.1f} packed FP32 operations (FFMA2/FMUL2/FADD2 = {pk} per 16 pairs, two pairs each), {cnt['MUFU'] / 16:.1f} MUFU, "
      f"{(cnt['FMNMX'] + cnt['FMNMX3']) / 16:.1f} FMNMX(3), {cnt['IADD3'] / 16:.2f} IADD3 + {cnt['LDS'] / 16:.2f} LDS (type-pair table and the step's j-group)")
print(f"# register pairs read by the FFMA2/FMUL2 (tools/sass_reads.py): {sorted(hist.items())}; read-limited cycles >= {cyc:.0f} against {2 * n} at two per operation")
print("# FFMA2 operand forms: Rn.F32x2.HI_LO = a register pair, Rn.F32 = one 32-bit register broadcast to both halves, .reuse = operand reuse cache")
for a, t in hot:
    print(f"/*{a:04x}*/  {t} ;")
