timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for cfg in "0 0 0 0 3" "0 640 0 0 3" "0 576 0 0 3" "0 512 0 0 3" "0 1024 0 0 3" "8 0 0 0 3" "0 0 0 0 4"; do
 echo "cfg=$cfg" >> gpurun_out/sweep.log
 timeout 120 python tools_probe.py cfg2 22 $cfg >> gpurun_out/sweep.log 2>&1
done
for cfg in "0 0 0 0 3" "0 512 0 0 3"; do
 echo "cfg4 cfg=$cfg" >> gpurun_out/sweep.log
 timeout 120 python tools_probe.py cfg4 148 $cfg >> gpurun_out/sweep.log 2>&1
 echo "cfg3 cfg=$cfg" >> gpurun_out/sweep.log
 timeout 120 python tools_probe.py cfg3 74 $cfg >> gpurun_out/sweep.log 2>&1
done
