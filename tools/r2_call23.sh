#!/bin/bash
# five epilogue groups at K = 64 (hi only, hi + f16 lo)
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_pipeline.py "tests/test_gpu_horizon.py::test_auto_precision_holds_the_gate_at_config_4" -m gpu -q -x -s --timeout=600 > $O/r2_tests23.log 2>&1
echo "pytest rc=$?" >> $O/r2_tests23.log
timeout 300 python bench.py --workload c4 --precision 3 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c23_c4_fast.json 2> $O/r2_c23_c4_fast.err
timeout 300 python bench.py --workload c4 --precision 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c23_c4_hilo.json 2> $O/r2_c23_c4_hilo.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c23_c4_auto.json 2> $O/r2_c23_c4_auto.err
echo done
