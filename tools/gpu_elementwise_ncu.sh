#!/bin/bash
# HBM roofline evidence of the elementwise / entropy-index kernels on >= 256 MB instances: event timings (jsonl) and an
# ncu pass with dram throughput / bytes per launch (csv).  ROUND_TAG names the outputs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
R=${ROUND_TAG:-r2}
timeout 300 python tools/bench_elementwise.py > gpurun_out/elementwise_roofline_$R.jsonl 2> gpurun_out/elementwise.err
cat gpurun_out/elementwise_roofline_$R.jsonl | cut -c1-170
EW_ITERS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --csv --log-file gpurun_out/elementwise_ncu_$R.csv python tools/bench_elementwise.py > /dev/null 2>> gpurun_out/elementwise.err
tail -2 gpurun_out/elementwise.err
