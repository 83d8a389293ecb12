#!/bin/bash
# One GPU call: parity of the sum-product kernels, their throughput A/B (ab/lib_A.so vs working tree), and a
# thread-count / cluster sweep of the Viterbi fill on config 2.
timeout 900 python -m pytest tests/test_forward.py tests/test_pairhmm.py -m gpu -x -q > gpurun_out/pytest_fwd.log 2>&1; tail -3 gpurun_out/pytest_fwd.log
: > gpurun_out/ab_fwd.log
for lib in A B; do
  if [ $lib = A ]; then export DNAB_LIB=$PWD/ab/lib_A.so; else unset DNAB_LIB; fi
  echo "lib=$lib" | tee -a gpurun_out/ab_fwd.log
  timeout 300 python bench.py --workload cfg5 --mode fwdback --steps 3 --warmup 3 --cpu-sample 0 2>/dev/null | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('fwdback cfg5 reads/s %.1f cells/s %.4g frac %.4f' % (j['reads_per_sec'], j['value'], j['roofline']['frac']))" | tee -a gpurun_out/ab_fwd.log
  python tools/probe_forward.py cfg5 592 2>&1 | tail -1 | tee -a gpurun_out/ab_fwd.log
  python tools/probe_forward.py cfg2 148 40 2>&1 | tail -1 | tee -a gpurun_out/ab_fwd.log
done
unset DNAB_LIB
for cfgv in "4 1024" "4 800" "4 640" "3 1024" "5 1024"; do
  set -- $cfgv
  timeout 300 python bench.py --workload cfg2 --steps 2 --warmup 3 --reads-per-step 480 --cpu-sample 0 --cluster $1 --threads $2 2>/dev/null | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('cfg2 cluster $1 threads $2: reads/s %.1f (cluster %s threads %s)' % (j['reads_per_sec'], j['config']['cluster_size'], j['config']['threads_per_cta']))" | tee -a gpurun_out/ab_fwd.log
done
