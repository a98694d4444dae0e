cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; L=gpurun_out/diag_host_mm.log
lscpu | grep -i "model name\|^CPU(s)\|hypervisor\|flags" | cut -c1-2000 > $L
python tools/diag_host_mm.py 15 >> $L 2>&1
# the same under the profiler's process injection (no kernels launch; it only attaches)
ncu --metrics gpu__time_duration.sum python tools/diag_host_mm.py 8 >> $L 2>&1
# and with the other cores busy
for i in $(seq 1 16); do (timeout 45 python -c "import torch; a=torch.randn(512,512)
while True: a=a@a; a/=a.norm()" &) ; done
python tools/diag_host_mm.py 8 >> $L 2>&1
cat $L | cut -c1-600
