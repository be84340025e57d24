#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; OUT=gpurun_out
for f in test_igemm_gpu test_pipeline_gpu; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --tb=short > $OUT/$f.txt 2>&1; echo "rc=$?" >> $OUT/$f.txt; tail -4 $OUT/$f.txt
done
for shape in "1 384 384 128 128 3" "1 384 384 256 256 3" "1 96 96 320 320 3"; do
  for cm in 0 1; do
    ONEDC_COLMODE=$cm timeout 120 python tools/igemm_roles.py $shape
  done
done 2>&1 | tee $OUT/igemm_roles.txt
for cm in 0 1; do
  ONEDC_COLMODE=$cm timeout 300 python tools/layer_table.py > $OUT/layer_table_cm$cm.txt 2> $OUT/layer_table.err
  ONEDC_COLMODE=$cm timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_cm$cm.txt 2> $OUT/bench_cm$cm.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_cm$cm.txt").read().strip().splitlines()[-1])
    print("colmode=$cm value", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "igemm ms", round(d["roofline"]["kernel_ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), "attn", round(d["roofline"]["attention"]["ms_per_step"],3))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_cm$cm.err").read()[-1500:])
PY
done
