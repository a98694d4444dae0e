#!/bin/bash
# Single-GPU evidence run: default bench, smoke under ncu (as the driver does), launch list, full capture of the main kernels.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_ncu.log 2>&1; echo "smoke under ncu rc=$?"; grep "smoke" gpurun_out/r02_smoke_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-epochs --no-hidden > gpurun_out/r02_ncu_launch.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"attn_x_bwd_mma_kernel|attn_x_fwd_kernel|logits_pack_mma|gemm_batched_tf32x3|gemm_tn_batched_ta" -s 30 -c 12 -o gpurun_out/prof_r02b python bench.py --steps 2 --warmup 3 --no-hidden --no-epochs > gpurun_out/ncu_r02b.log 2>&1; echo "full capture rc=$?"
