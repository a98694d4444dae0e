#!/bin/bash
# Second multi-GPU lease of round 2: the source-row-shard backward of the hidden layer.  usage: bash tools/multi_gpu_session2.sh <world>
W=$1
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
L=gpurun_out/mg_session2_${W}.log; : > $L
step() { echo "=== $1" >> $L; shift; timeout $1 bash -c "$2" >> $L 2>&1; echo "rc=$? ($SECONDS s)" >> $L; }
step "bench products x$W (check, hidden layer with the source-shard backward, PPI)" 420 "$TR --master-port 29602 bench.py --gpus $W --steps 10 > gpurun_out/r02_bench_${W}gpu_v2.json 2> gpurun_out/err.log; tail -c 300 gpurun_out/err.log"
step "dist_worker world $W (peer push)" 240 "$TR --master-port 29601 tests/dist_worker.py nccl 2>&1 | grep -v 'Warn\|warn' | tail -3"
step "hidden layer, one exchange chunk" 240 "GATK_SHARD_CHUNKS=1 $TR --master-port 29604 bench.py --gpus $W --steps 5 --workload products_hidden --no-epochs --no-check > gpurun_out/r02_hidden_${W}gpu_v2_c1.json 2> gpurun_out/err.log"
step "hidden layer, four exchange chunks" 240 "GATK_SHARD_CHUNKS=4 $TR --master-port 29606 bench.py --gpus $W --steps 5 --workload products_hidden --no-epochs --no-check > gpurun_out/r02_hidden_${W}gpu_v2_c4.json 2> gpurun_out/err.log"
grep -v "Warn\|warn\|sparse_coo" $L | tail -30
