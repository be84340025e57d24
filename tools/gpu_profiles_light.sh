#!/bin/bash
# Cheap refresh of profiles/: launch list, kineto breakdown, layer table, bench lines (no ncu --set full).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; OUT=gpurun_out
R=${ROUND_TAG:-r1}
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file $OUT/launches_$R.csv python tools/profile_step.py > $OUT/ncu_launch.log 2>&1
tail -1 $OUT/ncu_launch.log
timeout 300 python tools/kineto_step.py > $OUT/kineto.txt 2>&1
timeout 300 python tools/layer_table.py > $OUT/layer_table.txt 2> $OUT/layer_table.err
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.txt 2> $OUT/bench_n1.err; tail -c 300 $OUT/bench_n1.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.txt 2> $OUT/bench_ref.err; tail -c 300 $OUT/bench_ref.txt
