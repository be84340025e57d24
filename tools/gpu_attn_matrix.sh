#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for ns in 2 1; do for ks in 1 2 3 4; do
  echo -n "nsbuf=$ns kvsplit=$ks: "; ONEDC_ATTN_NSBUF=$ns ONEDC_ATTN_KVSPLIT=$ks python tools/attn_time.py 2>&1 | head -1
done; done
