#!/bin/bash
# Runs bench.py once per value of an environment knob and prints the per-kernel times side by side.
# usage (on the GPU box): bash tools/run_variants.sh GATK_XG_ALIGN 4 16
knob=$1; shift
for v in "$@"; do
  env $knob=$v timeout 200 python bench.py --steps 5 --warmup 3 --no-epochs > gpurun_out/bench_${knob}_$v.json 2> gpurun_out/bench_err.log
  python -c "
import json
d=json.load(open('gpurun_out/bench_${knob}_$v.json'))
print('$knob=$v', round(d['ms_per_step'], 3), {k: v['ms'] for k, v in d['kernels'].items()})"
done
