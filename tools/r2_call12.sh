#!/bin/bash
# packed-fp32 epilogue (FMUL2/FFMA2) of the tensor pass and the column-pair Gibbs sweep: correctness subset, benches, role cycles
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_pipeline.py tests/test_gpu_round2.py tests/test_gpu_entrypoints.py tests/test_gpu_configs.py -m gpu -q -x --timeout=600 > $O/r2_tests12.log 2>&1
echo "pytest rc=$?" >> $O/r2_tests12.log
timeout 300 python tests/tools/gibbs_bench.py > $O/r2_gibbs_bench5.log 2>&1
NBMF_LIB_VARIANT=nb64 timeout 300 python tests/tools/gibbs_bench.py > $O/r2_gibbs_bench5_nb64.log 2>&1
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c12_c4_auto.json 2> $O/r2_c12_c4_auto.err
timeout 300 python bench.py --workload c4 --precision 3 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c12_c4_fast.json 2> $O/r2_c12_c4_fast.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_c12_c5_auto.json 2> $O/r2_c12_c5_auto.err
timeout 600 python bench.py --precision 4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_c12_c5_lo8.json 2> $O/r2_c12_c5_lo8.err
export NBMF_LIB_VARIANT=prof
timeout 300 python tools/role_profile.py 65536 3 64 > $O/r2_roles_c4_fast.log 2>&1
timeout 300 python tools/role_profile.py 65536 4 64 > $O/r2_roles_c4_lo8.log 2>&1
timeout 300 python tools/role_profile.py 65536 3 128 > $O/r2_roles_64k_fast.log 2>&1
timeout 300 python tools/role_profile.py 65536 4 128 > $O/r2_roles_64k_lo8.log 2>&1
echo done
