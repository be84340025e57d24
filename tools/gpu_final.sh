#!/bin/bash
# Round-end check: the whole GPU suite exactly as the driver runs it, smoke(), and the default bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; OUT=gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.txt; tail -4 $OUT/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $OUT/bench_default.txt 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -c 700 $OUT/bench_default.txt
