#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
for v in u2 u2mb2; do NBMF_LIB_VARIANT=$v timeout 300 python tests/tools/gibbs_bench.py > $O/r2_gibbs_bench9_$v.log 2>&1; done
echo done
