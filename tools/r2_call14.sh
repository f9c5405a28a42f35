#!/bin/bash
# setmaxnreg + x64 TMEM load + early S-slot release (+ 4 S slots at K=64); Gibbs sweep with atomics / NB=64
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_pipeline.py tests/test_gpu_round2.py tests/test_gpu_entrypoints.py tests/test_gpu_configs.py -m gpu -q -x --timeout=600 > $O/r2_tests14.log 2>&1
echo "pytest rc=$?" >> $O/r2_tests14.log
timeout 300 python tests/tools/gibbs_bench.py > $O/r2_gibbs_bench6.log 2>&1
NBMF_LIB_VARIANT=mb2 timeout 300 python tests/tools/gibbs_bench.py > $O/r2_gibbs_bench6_mb2.log 2>&1
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c14_c4_auto.json 2> $O/r2_c14_c4_auto.err
timeout 300 python bench.py --workload c4 --precision 3 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c14_c4_fast.json 2> $O/r2_c14_c4_fast.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_c14_c5_auto.json 2> $O/r2_c14_c5_auto.err
timeout 600 python bench.py --precision 4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_c14_c5_lo8.json 2> $O/r2_c14_c5_lo8.err
export NBMF_LIB_VARIANT=prof
timeout 300 python tools/role_profile.py 65536 3 64 > $O/r2_roles14_c4_fast.log 2>&1
timeout 300 python tools/role_profile.py 65536 3 128 > $O/r2_roles14_64k_fast.log 2>&1
echo done
