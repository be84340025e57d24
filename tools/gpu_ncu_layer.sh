#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; OUT=gpurun_out
for cm in 0 1; do
  ONEDC_COLMODE=$cm python tools/one_layer.py 1 384 384 128 128 3 20
  ONEDC_COLMODE=$cm python tools/one_layer.py 1 384 384 256 256 3 20
  ONEDC_COLMODE=$cm timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_tc -s 2 -c 1 -f -o $OUT/layer128_cm$cm python tools/one_layer.py 1 384 384 128 128 3 2 > $OUT/ncu_layer_cm$cm.log 2>&1
  tail -2 $OUT/ncu_layer_cm$cm.log
done
