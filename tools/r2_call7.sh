#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_pipeline.py tests/test_gpu_round2.py tests/test_gpu_horizon.py -m gpu -q -x --timeout=600 -s > $O/r2_tests7.log 2>&1
echo "pytest rc=$?" >> $O/r2_tests7.log
timeout 600 python tools/error_growth.py 65536 128 1,2,5,10,20,40,60 lo8,fast > $O/r2_eg_64k_fold.log 2>&1
for p in 3 4; do
  timeout 600 python bench.py --workload c5 --precision $p --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_bench_c5_p${p}_fold.json 2> $O/r2_bench_c5_p${p}_fold.err
  NBMF_TP_SUB_STEPS=0 timeout 600 python bench.py --workload c5 --precision $p --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_bench_c5_p${p}_nofold.json 2> $O/r2_bench_c5_p${p}_nofold.err
done
timeout 300 python bench.py --workload c4 --precision 4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_bench_c4_p4_fold.json 2>/dev/null
timeout 1500 python tests/tools/config3_gibbs.py 2000 400 > $O/r2_config3_gibbs.log 2>&1
echo done
