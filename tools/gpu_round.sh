#!/bin/bash
# One gpurun call: probes, GPU test files (each in its own process under a timeout), a short bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi > $OUT/smi.txt 2>&1
python -c "import torch;print(torch.cuda.get_device_name(0), torch.version.cuda)" > $OUT/env.txt 2>&1
echo "== probe igemm";      timeout 120 python tools/tc_probe.py igemm      > $OUT/probe_igemm.txt 2>&1; echo "rc=$?" >> $OUT/probe_igemm.txt
echo "== probe attention";  timeout 120 python tools/tc_probe.py attention  > $OUT/probe_attn.txt 2>&1;  echo "rc=$?" >> $OUT/probe_attn.txt
tail -3 $OUT/probe_igemm.txt; tail -3 $OUT/probe_attn.txt
for f in ${TESTS:-test_elementwise_gpu test_igemm_gpu test_attention_gpu test_pipeline_gpu}; do
  echo "== $f"
  timeout ${TEST_TIMEOUT:-900} python -m pytest tests/$f.py -q -m gpu -s --tb=short > $OUT/$f.txt 2>&1
  echo "rc=$?" >> $OUT/$f.txt
  tail -4 $OUT/$f.txt
done
if [ -n "$BENCH" ]; then
  echo "== bench"
  timeout 900 python bench.py $BENCH > $OUT/bench.txt 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.txt
  tail -2 $OUT/bench.txt; tail -5 $OUT/bench.err
fi
