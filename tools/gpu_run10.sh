#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; OUT=gpurun_out
for tr in 1 0; do
ONEDC_TRANSPOSED=$tr timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_tr$tr.txt 2> $OUT/bench.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_tr$tr.txt").read().strip().splitlines()[-1])
    print("transposed=$tr value", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "igemm ms", round(d["roofline"]["kernel_ms_per_step"],3), "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench.err").read()[-1500:])
PY
done
