#!/bin/bash
# the round-end sequence on one B200: GPU tests, smoke, the default bench line, the reference arm, then the launch list of the
# same bench command under ncu (only after the command has exited 0 without it)
o=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $o/final_pytest.log 2>&1; echo rc=$? >> $o/final_pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > $o/final_smoke.log 2>&1; echo rc=$? >> $o/final_smoke.log
date +%s > $o/final_t0
timeout 900 python bench.py > $o/final_bench.json 2> $o/final_bench.err; rc=$?; echo rc=$rc >> $o/final_bench.err
date +%s > $o/final_t1
timeout 600 python bench.py --impl reference > $o/final_bench_ref.json 2> $o/final_bench_ref.err; echo rc=$? >> $o/final_bench_ref.err
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/final_launches.csv python bench.py --steps 1 --warmup 3 > $o/final_ncu.log 2>&1; echo rc=$? >> $o/final_ncu.log
fi
