for al in 4 16; do
GATK_XG_ALIGN=$al timeout 200 python bench.py --steps 5 --warmup 3 --no-epochs > gpurun_out/bench38_al$al.json 2> gpurun_out/bench_err.log
python -c "
import json
d=json.load(open('gpurun_out/bench38_al$al.json'))
print('align=$al', d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
done
