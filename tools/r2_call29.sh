#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_configs.py tests/test_gpu_entrypoints.py -m gpu -q -x --timeout=300 -k "gibbs or Gibbs or config or sample" > $O/r2_tests29.log 2>&1
echo "pytest rc=$?" >> $O/r2_tests29.log
timeout 300 python tests/tools/gibbs_bench.py > $O/r2_gibbs_bench10.log 2>&1
echo done
