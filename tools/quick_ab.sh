#!/bin/bash
# Smallest useful GPU call: Viterbi parity tests, then config 2 A/B (ab/lib_A.so vs working tree).
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
for lib in A B; do
  if [ $lib = A ]; then export DNAB_LIB=$PWD/ab/lib_A.so; else unset DNAB_LIB; fi
  timeout 100 python bench.py --workload cfg2 --steps 2 --warmup 3 --reads-per-step 480 --cpu-sample 0 2>/dev/null | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('cfg2 lib=$lib reads/s %.1f' % j['reads_per_sec'])"
done
