#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; OUT=gpurun_out
timeout 600 python -m pytest tests/test_igemm_gpu.py -q -m gpu -x --tb=short -k "transposed or column" 2>&1 | tail -3
for tr in 0 1; do ONEDC_TRANSPOSED=$tr timeout 120 python tools/one_layer.py 1 768 768 128 128 3 10; ONEDC_TRANSPOSED=$tr timeout 120 python tools/one_layer.py 1 768 768 256 128 3 10; ONEDC_TRANSPOSED=$tr timeout 120 python tools/igemm_roles.py 1 768 768 128 128 3;  done
