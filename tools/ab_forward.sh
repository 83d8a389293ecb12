#!/bin/bash
# One GPU call: parity of the sum-product kernels on the working-tree library, then their throughput
# (optionally against ab/lib_A.so: pass "ab" as the first argument).
timeout 900 python -m pytest tests/test_forward.py tests/test_pairhmm.py -m gpu -x -q > gpurun_out/pytest_fwd.log 2>&1; tail -3 gpurun_out/pytest_fwd.log
: > gpurun_out/ab_fwd2.log
libs="B"; [ "$1" = ab ] && libs="A B"
for lib in $libs; do
  if [ $lib = A ]; then export DNAB_LIB=$PWD/ab/lib_A.so; else unset DNAB_LIB; fi
  echo "lib=$lib" | tee -a gpurun_out/ab_fwd2.log
  timeout 300 python bench.py --workload cfg5 --mode fwdback --steps 3 --warmup 3 --cpu-sample 0 2>/dev/null | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('fwdback cfg5 reads/s %.1f cells/s %.4g frac %.4f' % (j['reads_per_sec'], j['value'], j['roofline']['frac']))" | tee -a gpurun_out/ab_fwd2.log
  python tools/probe_forward.py cfg5 592 2>&1 | tail -1 | tee -a gpurun_out/ab_fwd2.log
  python tools/probe_forward.py cfg2 148 40 2>&1 | tail -1 | tee -a gpurun_out/ab_fwd2.log
  python tools/probe_forward.py cfg4 592 2>&1 | tail -1 | tee -a gpurun_out/ab_fwd2.log
  python tools/probe_forward.py cfg3 592 2>&1 | tail -1 | tee -a gpurun_out/ab_fwd2.log
  python tools/probe_forward.py cfg1 2960 2>&1 | tail -1 | tee -a gpurun_out/ab_fwd2.log
done
if [ "$2" = ncu ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:forwardKernel -s 1 -c 1 -f -o gpurun_out/r01_forward_cfg5 python bench.py --workload cfg5 --mode fwdback --steps 1 --warmup 1 --reads-per-step 148 --cpu-sample 0 > gpurun_out/ncu_fwd.log 2>&1
  tail -2 gpurun_out/ncu_fwd.log
fi
