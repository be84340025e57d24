#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; OUT=gpurun_out
timeout 600 python -m pytest tests/test_igemm_gpu.py -q -m gpu -x --tb=short > $OUT/test_igemm_gpu.txt 2>&1; echo "rc=$?" >> $OUT/test_igemm_gpu.txt; tail -25 $OUT/test_igemm_gpu.txt
ONEDC_COLMODE=2 timeout 600 python -m pytest tests/test_igemm_gpu.py -q -m gpu -x --tb=short 2>&1 | tail -3
for f in test_elementwise_gpu test_pipeline_gpu test_configs_gpu; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --tb=short > $OUT/$f.txt 2>&1; echo "rc=$?" >> $OUT/$f.txt; tail -3 $OUT/$f.txt
done
for tr in 0 1; do ONEDC_TRANSPOSED=$tr timeout 120 python tools/one_layer.py 1 768 768 128 128 3 10; ONEDC_TRANSPOSED=$tr timeout 120 python tools/one_layer.py 1 768 768 256 128 3 10; done
timeout 300 python tools/kineto_step.py > $OUT/kineto.txt 2>&1; grep -v "Warn\|_warn" $OUT/kineto.txt | head -6
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench.txt 2> $OUT/bench.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench.txt").read().strip().splitlines()[-1])
    print("value", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "rans", round(d["host_rans_ms_per_step"],3), "igemm ms", round(d["roofline"]["kernel_ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), "attn", round(d["roofline"]["attention"]["ms_per_step"],3), "launches", d["gpu_launches"])
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench.err").read()[-1500:])
PY
