#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
for wl in cfg1 cfg1; do
timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --cpu-sample 0 2>/dev/null | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('$wl reads/s %.1f clusters %s threads %s' % (j['reads_per_sec'], j['config'].get('reads_in_flight'), j['config']['threads_per_cta']))"
done
