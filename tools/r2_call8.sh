#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_pipeline.py tests/test_gpu_round2.py tests/test_gpu_configs.py -m gpu -q -x --timeout=600 -s > $O/r2_tests8.log 2>&1
echo "pytest rc=$?" >> $O/r2_tests8.log
for p in 3 4; do
  timeout 600 python bench.py --workload c5 --precision $p --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_bench_c5_p${p}_fold2.json 2> $O/r2_bench_c5_p${p}_fold2.err
done
NBMF_TP_SUB_STEPS=0 timeout 600 python bench.py --workload c5 --precision 4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_bench_c5_p4_nofold2.json 2>/dev/null
timeout 300 python tests/tools/gibbs_bench.py > $O/r2_gibbs_bench3.log 2>&1
timeout 300 python tools/error_growth.py 65536 128 1,20,40 lo8 > $O/r2_eg_64k_fold2.log 2>&1
export NBMF_TP_SPIN_LIMIT=2000000000
S="python tools/smallk_bench.py quick shapes=2,4,5 iters=2"
$S > $O/r2_plain_smallk2.log 2>&1 && timeout 900 ncu --set full --clock-control none -k regex:pass_mma64\|pass_simt -s 12 -c 12 -o $O/r2_prof_smallk2 $S > $O/r2_ncu_f_smallk2.log 2>&1
ncu -i $O/r2_prof_smallk2.ncu-rep --page raw --csv > $O/r2_prof_smallk2_raw.csv 2>/dev/null; rm -f $O/r2_prof_smallk2.ncu-rep
echo done
