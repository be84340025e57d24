#!/bin/bash
# One gpurun call that produces the raw material of profiles/ (see tools/make_profiles.py for the post-processing).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; OUT=gpurun_out
R=${ROUND_TAG:-r1}
# every launch of one device-resident 768x768 decode step: duration + DRAM bytes (cold-cache, serialised)
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file $OUT/launches_$R.csv python tools/profile_step.py > $OUT/ncu_launch.log 2>&1
tail -1 $OUT/ncu_launch.log
# ncu --set full of the implicit-GEMM launches of the VAE's last levels (the biggest layers) and of one UNet self-attention
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:igemm_tc_kernel -s 404 -c 14 \
  -f -o $OUT/igemm_$R python tools/profile_step.py > $OUT/ncu_full.log 2>&1
tail -1 $OUT/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tc -s 0 -c 1 \
  -f -o $OUT/attn_$R python tools/profile_step.py > $OUT/ncu_attn.log 2>&1
tail -1 $OUT/ncu_attn.log
# raw pages as CSV (small); keep a report only when it is small enough to travel back
for n in igemm_$R attn_$R; do
  ncu -i $OUT/$n.ncu-rep --page raw --csv > $OUT/${n}_raw.csv 2>/dev/null
  sz=$(stat -c %s $OUT/$n.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 20000000 ]; then rm -f $OUT/$n.ncu-rep; fi
done
ls -la $OUT
timeout 300 python tools/kineto_step.py > $OUT/kineto.txt 2>&1
timeout 300 python tools/layer_table.py > $OUT/layer_table.txt 2> $OUT/layer_table.err
timeout 300 python tools/bench_elementwise.py > $OUT/elementwise_roofline.jsonl 2> $OUT/elementwise.err
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.txt 2> $OUT/bench_n1.err; tail -c 600 $OUT/bench_n1.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.txt 2> $OUT/bench_ref.err; tail -c 400 $OUT/bench_ref.txt
