for pf in 0 8; do
GATK_PF=$pf timeout 200 python bench.py --steps 5 --warmup 3 --no-epochs > gpurun_out/bench34_pf$pf.json 2> gpurun_out/bench_err.log
python -c "
import json
d=json.load(open('gpurun_out/bench34_pf$pf.json'))
print('pf=$pf', d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items() if 'gemm' in k})"
done
