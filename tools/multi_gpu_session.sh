#!/bin/bash
# One multi-GPU lease, everything that needs it: usage  bash tools/multi_gpu_session.sh <world> [quick]
W=$1; QUICK=$2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
L=gpurun_out/mg_session_${W}.log; : > $L
step() { echo "=== $1" >> $L; shift; timeout $1 bash -c "$2" >> $L 2>&1; echo "rc=$? ($SECONDS s)" >> $L; }
step "bench products x$W (check, hidden layer, PPI)" 420 "$TR --master-port 29602 bench.py --gpus $W --steps 10 > gpurun_out/r02_bench_${W}gpu.json 2> gpurun_out/r02_bench_${W}gpu.err; tail -c 300 gpurun_out/r02_bench_${W}gpu.err"
if [ "$QUICK" != "quick" ]; then
step "bench papers_shard x$W" 420 "$TR --master-port 29603 bench.py --gpus $W --steps 5 --workload papers_shard > gpurun_out/r02_papers_${W}gpu.json 2> gpurun_out/r02_papers_${W}gpu.err; tail -c 600 gpurun_out/r02_papers_${W}gpu.err"
step "dist_worker world $W (peer push)" 240 "$TR --master-port 29601 tests/dist_worker.py nccl 2>&1 | grep -v 'Warn\|warn' | tail -5"
step "hidden layer, one exchange chunk" 240 "GATK_SHARD_CHUNKS=1 $TR --master-port 29604 bench.py --gpus $W --steps 5 --workload products_hidden --no-epochs --no-check > gpurun_out/r02_hidden_${W}gpu_c1.json 2> gpurun_out/err.log"
step "dist_worker world $W (NCCL exchanges)" 240 "GATK_PEER_PUSH=0 $TR --master-port 29605 tests/dist_worker.py nccl 2>&1 | grep -v 'Warn\|warn' | tail -5"
else
step "dist_worker world $W (peer push)" 240 "$TR --master-port 29601 tests/dist_worker.py nccl 2>&1 | grep -v 'Warn\|warn' | tail -5"
step "bench papers_tiny x$W" 240 "$TR --master-port 29603 bench.py --gpus $W --steps 5 --workload papers_tiny > gpurun_out/r02_papers_tiny_${W}gpu.json 2> gpurun_out/err.log; tail -c 300 gpurun_out/err.log"
step "hidden layer workload" 240 "$TR --master-port 29604 bench.py --gpus $W --steps 5 --workload products_hidden --no-epochs --no-check > gpurun_out/r02_hidden_${W}gpu.json 2> gpurun_out/err.log; tail -c 300 gpurun_out/err.log"
fi
grep -v "Warn\|warn\|sparse_coo" $L | tail -40
