#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for shape in "1 384 384 128 128 3" "1 384 384 256 256 3" "1 96 96 320 320 3" "1 1 9216 320 960 1"; do
  for cm in 0 1; do
    ONEDC_COLMODE=$cm python tools/igemm_roles.py $shape
  done
done 2>&1 | tee gpurun_out/igemm_roles.txt
