#!/bin/bash
# third GPU call: LO8 variant (fp8 lo term) correctness + timing, red throughput microbenchmark
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_tensor.py tests/test_gpu_pipeline.py -m gpu -q --timeout=300 -s -k "tensor or pipelined" > $O/r2_tests3.log 2>&1
echo "pytest rc=$?" >> $O/r2_tests3.log
timeout 120 tools/ubench/red_throughput 1024 > $O/r2_red_throughput.log 2>&1
timeout 600 python tools/error_growth.py 16384 128 1,2,5,10,20,40,60 tensor,lo8,fast > $O/r2_eg_16k_lo8.log 2>&1
timeout 600 python tools/error_growth.py 65536 128 1,2,5,10,20,40,60 tensor,lo8,fast 128 > $O/r2_eg_64k_lo8.log 2>&1
for p in 4 2; do
  timeout 600 python bench.py --workload c5 --precision $p --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_bench_c5_p${p}_b.json 2> $O/r2_bench_c5_p${p}_b.err
done
timeout 300 python bench.py --workload c4 --precision 4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_bench_c4_p4.json 2> $O/r2_bench_c4_p4.err
timeout 300 python bench.py --workload c4 --precision 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_bench_c4_p2.json 2> $O/r2_bench_c4_p2.err
timeout 300 python bench.py --workload c4 --precision 3 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_bench_c4_p3.json 2> $O/r2_bench_c4_p3.err
echo done
