#!/bin/bash
# rebuild with different knobs and print the bonded kernel time at B=4096 and B=64
for defs in "$@"; do
  MLP_NVCC_DEFS="$defs" python -m molearn_b200.build --force > /dev/null 2>&1
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline --bonded-only 2>/dev/null | tail -n 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline_bonded']; print('$defs', 'B4096_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), 'B64_ms', round(d['roofline']['other_kernels_ms']['bonded'],4))"
done
