#!/bin/bash
# one GPU iteration of the batch-kernel bring-up: parity of a few golden cases, then short benches
# usage: tools/gpu_iter.sh <tag> "<opts for try>" "<bench opt sets separated by ;>" "<workloads>"
tag=$1; tryopts=$2; benchsets=$3; wls=${4:-"cfg1 cfg4 cfg2"}
timeout 200 python tools/gpu_try_batch.py kat147 l4c4_global_mixed l4c4_local_mixed l4c4_len12_local mr2l4c4_local cfg4_global_dels kat185 cfg5_l8_local cfg3_global_indels cfg2_global_subs $tryopts > gpurun_out/try_$tag.log 2>&1; echo rc=$? >> gpurun_out/try_$tag.log
for spec in "cfg1 200" "cfg4 300" "cfg2 70"; do
  timeout 120 python tools/gpu_try_batch.py --agree $spec $tryopts >> gpurun_out/try_$tag.log 2>&1; echo rc=$? >> gpurun_out/try_$tag.log
done
grep -q "rc=124" gpurun_out/try_$tag.log && { echo "HANG in parity; skipping benches" >> gpurun_out/try_$tag.log; exit 1; }
IFS=';' read -ra sets <<< "$benchsets"
i=0
for set in "${sets[@]}"; do
  for w in $wls; do
    timeout 120 python bench.py --workload $w --steps 3 --warmup 1 --cpu-sample 0 $set > gpurun_out/bench_${tag}_${i}_$w.log 2> gpurun_out/bench_${tag}_${i}_$w.err; echo "rc=$? opts=$set" >> gpurun_out/bench_${tag}_${i}_$w.log
  done
  i=$((i+1))
done
