#!/bin/bash
# One GPU call: parity tests, then the bench of every workload, then the forward probes (logs under gpurun_out/).
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg2.log 2> gpurun_out/bench_cfg2.err
for wl in cfg3 cfg4 cfg5 cfg1; do
  timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --cpu-sample 8 > gpurun_out/bench_$wl.log 2> gpurun_out/bench_$wl.err
done
for wl in cfg2 cfg3 cfg4 cfg5 cfg1; do tail -c 2500 gpurun_out/bench_$wl.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); c = j.get('cpu_baseline') or {}
        print('$wl', 'reads/s %.0f cells/s %.3g e2e %.3g frac %.4f cluster %s thr %s cpu reads/s %s' % (j['reads_per_sec'], j['value'], j['e2e']['value'], j['roofline']['frac'], j['config']['cluster_size'], j['config']['threads_per_cta'], c.get('reads_per_sec')))
"; done
python tools/probe_forward.py cfg2 148 40 > gpurun_out/forward.log 2>&1
python tools/probe_forward.py cfg5 592 >> gpurun_out/forward.log 2>&1
python tools/probe_forward.py cfg4 592 >> gpurun_out/forward.log 2>&1
python tools/probe_forward.py cfg1 2960 >> gpurun_out/forward.log 2>&1
cat gpurun_out/forward.log
