#!/bin/bash
# One GPU call for the round's evidence: parity tests, the default bench, the ncu launch list and full capture of the
# fill kernel, the other workloads, and the sum-product kernels A/B (ab/lib_B.so, ab/lib_C.so = candidate builds of the forward kernel).
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg2.log 2> gpurun_out/bench_cfg2.err; tail -c 400 gpurun_out/bench_cfg2.log
bash tools/profile_round.sh
: > gpurun_out/ab_fwd.log
# A = the committed library; B = closure evaluated on the frontier; C = B + two states per thread in lockstep
for lib in B C; do
  DNAB_LIB=$PWD/ab/lib_$lib.so timeout 600 python -m pytest tests/test_forward.py -m gpu -x -q > gpurun_out/pytest_fwd_$lib.log 2>&1; echo "lib $lib parity:"; tail -2 gpurun_out/pytest_fwd_$lib.log
done
for lib in A B C; do
  if [ $lib = A ]; then unset DNAB_LIB; else export DNAB_LIB=$PWD/ab/lib_$lib.so; fi
  echo "lib=$lib" | tee -a gpurun_out/ab_fwd.log
  timeout 300 python bench.py --workload cfg5 --mode fwdback --steps 3 --warmup 3 --cpu-sample 0 2>/dev/null | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('fwdback cfg5 reads/s %.1f cells/s %.4g frac %.4f' % (j['reads_per_sec'], j['value'], j['roofline']['frac']))" | tee -a gpurun_out/ab_fwd.log
  python tools/probe_forward.py cfg5 592 2>&1 | tail -1 | tee -a gpurun_out/ab_fwd.log
  python tools/probe_forward.py cfg2 148 40 2>&1 | tail -1 | tee -a gpurun_out/ab_fwd.log
  python tools/probe_forward.py cfg4 592 2>&1 | tail -1 | tee -a gpurun_out/ab_fwd.log
  python tools/probe_forward.py cfg1 2960 2>&1 | tail -1 | tee -a gpurun_out/ab_fwd.log
done
unset DNAB_LIB
for wl in cfg3 cfg4 cfg5 cfg1; do
  timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --cpu-sample 8 > gpurun_out/bench_$wl.log 2> gpurun_out/bench_$wl.err
done
timeout 300 python bench.py --workload cfg5 --mode fwdback --steps 3 --warmup 3 > gpurun_out/bench_cfg5_fwdback.log 2> gpurun_out/bench_cfg5_fwdback.err
for wl in cfg2 cfg3 cfg4 cfg5 cfg1 cfg5_fwdback; do tail -c 3000 gpurun_out/bench_$wl.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); c = j.get('cpu_baseline') or {}
        print('$wl', 'reads/s %.0f cells/s %.3g e2e %.3g frac %.4f cluster %s thr %s cpu reads/s %s' % (j['reads_per_sec'], j['value'], j['e2e']['value'], j['roofline']['frac'], j['config'].get('cluster_size'), j['config'].get('threads_per_cta'), c.get('reads_per_sec')))
"; done
# full capture of the sum-product kernel (forward + backward + counts) on the -l 8 machine
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forwardKernel -s 1 -c 1 -f -o gpurun_out/r01_forward_cfg5 python bench.py --workload cfg5 --mode fwdback --steps 1 --warmup 1 --reads-per-step 148 --cpu-sample 0 > gpurun_out/ncu_fwd.log 2>&1
tail -2 gpurun_out/ncu_fwd.log
