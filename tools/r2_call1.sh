#!/bin/bash
# first GPU call of round 2: tests, error growth curves, quick timings
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r2_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > $O/r2_tests1.log 2>&1
echo "pytest rc=$?" >> $O/r2_tests1.log
timeout 600 python tools/error_growth.py 16384 128 1,2,5,10,20,40,60,80,100 tensor,fast,acc 0,128 > $O/r2_eg_16k.log 2>&1
timeout 900 python tools/error_growth.py 65536 128 1,2,5,10,20,40,60,80,100 tensor,fast,acc 0,128 > $O/r2_eg_64k.log 2>&1
for p in 3 2 4; do
  timeout 600 python bench.py --workload c5 --precision $p --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_bench_c5_p$p.json 2> $O/r2_bench_c5_p$p.err
done
timeout 600 python tests/tools/gibbs_bench.py > $O/r2_gibbs_bench1.log 2>&1
echo done
