#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; OUT=gpurun_out
for b in 4 8; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch $b > $OUT/bench_b$b.txt 2> $OUT/bench_b$b.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_b$b.txt").read().strip().splitlines()[-1])
    print("batch $b value", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_b$b.err").read()[-2500:])
PY
done
timeout 600 python bench.py --steps 24 --warmup 5 --no-cpu-baseline --pipeline 4 > $OUT/bench_p4.txt 2> $OUT/bench.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_p4.txt").read().strip().splitlines()[-1])
print("pipelined 4:", d.get("e2e_pipelined"))
PY
