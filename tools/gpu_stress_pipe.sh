#!/bin/bash
# Does the pipelined decoder (three streams of graphs in flight) survive?  kodak64 (64 x 768x512, depth 3) a few times, each
# under its own timeout: a deadlock between concurrently running kernels shows up as a timeout here.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  timeout 120 python bench.py --workload kodak64 --steps 8 --warmup 2 --no-cpu-baseline > gpurun_out/stress_$i.txt 2> gpurun_out/stress_$i.err
  echo "run $i rc=$? $(tail -c 160 gpurun_out/stress_$i.txt)"
done
