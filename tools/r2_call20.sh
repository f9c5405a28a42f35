#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_entrypoints.py tests/test_gpu_configs.py tests/test_gpu_parity.py "tests/test_gpu_horizon.py::test_auto_precision_holds_the_gate_at_config_4" -m gpu -q -x -s --timeout=600 -k "gibbs or Gibbs or sample or config or draws or gate" > $O/r2_tests20.log 2>&1
echo "pytest rc=$?" >> $O/r2_tests20.log
timeout 300 python tests/tools/gibbs_bench.py > $O/r2_gibbs_bench8.log 2>&1
timeout 300 python bench.py --workload c4 --precision 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c20_c4_hilo.json 2> $O/r2_c20_c4_hilo.err
echo done
