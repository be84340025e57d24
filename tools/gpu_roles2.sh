#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for shape in "1 12 12 1280 1280 3" "1 24 24 1280 1280 3"; do
  timeout 120 python tools/igemm_roles.py $shape
done 2>&1 | tee gpurun_out/igemm_roles_small.txt
