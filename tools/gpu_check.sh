#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; OUT=gpurun_out
for f in test_pipeline_gpu test_configs_gpu; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --tb=short > $OUT/$f.txt 2>&1; echo "rc=$?" >> $OUT/$f.txt; tail -3 $OUT/$f.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench.txt 2> $OUT/bench.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench.txt").read().strip().splitlines()[-1])
    print("value", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "ms", round(d["e2e"]["ms_per_step"],3), "p50", round(d["p50_ms_per_image_e2e"],3), "pipelined", round(d["e2e_pipelined"]["value"],2))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench.err").read()[-2500:])
PY
