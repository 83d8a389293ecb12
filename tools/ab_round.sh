#!/bin/bash
# One GPU call: parity tests on the working-tree library, then bench A/B (ab/lib_A.so = previous commit, default = working tree).
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/ab.log
run() { # label lib workload
  if [ "$2" = A ]; then export DNAB_LIB=$PWD/ab/lib_A.so; else unset DNAB_LIB; fi
  timeout 300 python bench.py --workload $3 --steps 3 --warmup 3 --cpu-sample 0 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l)
        print('$3 lib=$2 reads/s %.1f cells/s %.4g e2e %.4g' % (j['reads_per_sec'], j['value'], j['e2e']['value']))
" | tee -a gpurun_out/ab.log
}
for rep in 1 2; do run x A cfg2; run x B cfg2; done
for wl in cfg3 cfg4 cfg5 cfg1; do run x A $wl; run x B $wl; done
unset DNAB_LIB; python tools/probe_phases.py cfg2 66 2>&1 | tail -1 | tee gpurun_out/probe_phases_cfg2.log
