#!/bin/bash
# rebuild with different knobs and print step / pair kernel / reduce kernel times of the headline batch
for defs in "$@"; do
  MLP_NVCC_DEFS="$defs" python -m molearn_b200.build --force > /dev/null 2>&1
  python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$defs', 'value', round(d['value']), 'step_ms', round(d['ms_per_step'],4), 'nb_ms', round(r['kernel_ms'],4), 'reduce_ms', round(r['other_kernels_ms']['nb_reduce'],4), 'bonded_ms', round(r['other_kernels_ms']['bonded'],4))"
done
