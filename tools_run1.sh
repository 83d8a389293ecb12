timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for wl in cfg3 cfg4; do for tn in 0 9999; do
 echo "$wl tailN=$tn" >> gpurun_out/sweep.log
 DNAB_TAIL_N=$tn timeout 120 python tools_probe.py $wl 296 >> gpurun_out/sweep.log 2>&1
done; echo "$wl default" >> gpurun_out/sweep.log; timeout 120 python tools_probe.py $wl 296 >> gpurun_out/sweep.log 2>&1; done
