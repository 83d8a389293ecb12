#!/bin/bash
# The ncu evidence committed under profiles/: launch list of the default bench command, full capture of the fill kernel.
python bench.py --steps 2 --warmup 1 > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_l.log 2>&1
python bench.py --steps 1 --warmup 1 --reads-per-step 66 --cpu-sample 0 > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:viterbiFill -s 1 -c 1 -f -o gpurun_out/r01_fill_push_cfg2 python bench.py --steps 1 --warmup 1 --reads-per-step 66 --cpu-sample 0 > gpurun_out/ncu_f.log 2>&1
tail -1 gpurun_out/ncu_f.log
