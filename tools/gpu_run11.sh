#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; OUT=gpurun_out
for f in test_elementwise_gpu test_pipeline_gpu test_configs_gpu; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --tb=short > $OUT/$f.txt 2>&1; echo "rc=$?" >> $OUT/$f.txt; tail -3 $OUT/$f.txt
done
timeout 300 python tools/bench_elementwise.py 2> $OUT/elementwise.err | tee $OUT/elementwise_roofline.jsonl | cut -c1-140
