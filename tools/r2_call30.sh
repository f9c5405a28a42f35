#!/bin/bash
# final tree: error growth of the final kernels (packed-fp32 epilogue) at 65536^2, K = 128 beside the round's table; whole GPU suite
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python tools/error_growth.py 65536 128 1,10,20,40 lo8,fast > $O/r2_error_growth_64k_final.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $O/r2_gpu_tests_final3.log 2>&1
echo "pytest rc=$?" >> $O/r2_gpu_tests_final3.log
timeout 300 python bench.py --workload c3 --steps 200 --warmup 5 > $O/r2h_bench_c3.json 2> $O/r2h_bench_c3.err
echo done
