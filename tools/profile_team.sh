#!/bin/bash
# ncu --set full of ONE launch of the read-batched kernel on a 148-CTA team (config 2, 32 reads), after the same command has
# exited 0 without ncu; summary + by-phase view -> gpurun_out/r02_batch_cfg2_team_summary.txt
cmd="python bench.py --steps 1 --warmup 1 --others 0 --cpu-sample 0 --reads-per-step 32 --opt kernel=1"
timeout 200 $cmd > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viterbiFillBatch -s 1 -c 1 -f -o gpurun_out/r02_batch_cfg2_team $cmd > gpurun_out/ncu_team.log 2>&1
tail -1 gpurun_out/ncu_team.log
