#!/bin/bash
# Reproduce "parity moves under ncu" and run the sanitizers over the same tiny case.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/diag.log
: > $L
python tools/diag_nondet.py run plain >> $L 2>&1
python tools/diag_nondet.py run plain2 >> $L 2>&1
GATK_POISON=1 python tools/diag_nondet.py run poison >> $L 2>&1
ncu --metrics gpu__time_duration.sum --log-file gpurun_out/diag_ncu_a.log python tools/diag_nondet.py run ncu >> $L 2>&1
ncu --metrics gpu__time_duration.sum --cache-control none --log-file gpurun_out/diag_ncu_b.log python tools/diag_nondet.py run ncu_nocache >> $L 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --log-file gpurun_out/diag_ncu_c.log python tools/diag_nondet.py run ncu_noclock >> $L 2>&1
for t in plain2 poison ncu ncu_nocache ncu_noclock; do echo "== plain vs $t" >> $L; python tools/diag_nondet.py cmp plain $t >> $L 2>&1; done
for tool in memcheck initcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/san_$tool.log python tools/diag_nondet.py run san_$tool >> $L 2>&1
  echo "sanitizer $tool rc=$?" >> $L
  tail -3 gpurun_out/san_$tool.log >> $L
done
tail -100 $L
