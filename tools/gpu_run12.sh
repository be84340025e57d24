#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; OUT=gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -x --tb=short > $OUT/test_pipeline_gpu.txt 2>&1; echo "rc=$?" >> $OUT/test_pipeline_gpu.txt; tail -12 $OUT/test_pipeline_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for d in 2 3; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --pipeline $d > $OUT/bench_p$d.txt 2> $OUT/bench.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_p$d.txt").read().strip().splitlines()[-1])
    print("value", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "ms", round(d["e2e"]["ms_per_step"],3), "pipelined", d.get("e2e_pipelined"))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench.err").read()[-2500:])
PY
done
