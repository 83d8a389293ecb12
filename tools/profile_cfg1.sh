#!/bin/bash
# ncu --set full of ONE launch of the read-batched kernel on dnastore-l4 (T = 1, 148 CTAs, 4,736 reads), after the same
# command has exited 0 without ncu
cmd="python bench.py --workload cfg1 --steps 1 --warmup 1 --others 0 --cpu-sample 0 --reads-per-step 4736"
timeout 200 $cmd > gpurun_out/prof_plain1.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbiFillBatch -s 1 -c 1 -f -o gpurun_out/r02_batch_cfg1 $cmd > gpurun_out/ncu_cfg1.log 2>&1
tail -1 gpurun_out/ncu_cfg1.log
