#!/bin/bash
# One gpurun call of round 2: every GPU test file in its own process under a timeout, smoke, the default bench line.
# TESTS / BENCH / SMOKE select parts:  TESTS="test_a test_b" BENCH="--steps 20 --warmup 5" SMOKE=1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/fullsize_parity.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
for f in ${TESTS-test_elementwise_gpu test_igemm_gpu test_attention_gpu test_pipeline_gpu test_configs_gpu test_fullsize_parity_gpu test_dropin_api}; do
  echo "== $f"
  timeout ${TEST_TIMEOUT:-1200} python -m pytest tests/$f.py -q -m gpu -s --tb=short > $OUT/$f.txt 2>&1
  echo "rc=$?" >> $OUT/$f.txt
  tail -4 $OUT/$f.txt
done
if [ -n "$SMOKE" ]; then
  echo "== smoke"
  timeout 600 python __graft_entry__.py smoke > $OUT/smoke.txt 2>&1; echo "rc=$?" >> $OUT/smoke.txt; tail -2 $OUT/smoke.txt
fi
if [ -n "$BENCH" ]; then
  echo "== bench"
  timeout 1200 python bench.py $BENCH > $OUT/bench.txt 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.txt
  tail -c 3000 $OUT/bench.txt; tail -5 $OUT/bench.err
fi
