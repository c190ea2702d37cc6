#!/bin/bash
# rebuild the library with different compile-time knobs and print the NB kernel time of each
for defs in "$@"; do
  MLP_NVCC_DEFS="$defs" python -m molearn_b200.build --force > /dev/null 2>&1
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$defs', 'value', round(d['value']), 'nb_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), 'bonded_ms', round(r['other_kernels_ms']['bonded'],4))"
done
