#!/bin/bash
# final tree: whole GPU suite, smoke, the bench lines that changed since r2_call18 (C4, C3) and the default line
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $O/r2_gpu_tests_final2.log 2>&1
echo "pytest rc=$?" >> $O/r2_gpu_tests_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke_final2.log 2>&1
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > $O/r2g_bench_c4_auto.json 2> $O/r2g_bench_c4_auto.err
timeout 300 python bench.py --workload c4 --precision 3 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2g_bench_c4_fast.json 2> $O/r2g_bench_c4_fast.err
timeout 300 python bench.py --workload c3 --steps 200 --warmup 5 > $O/r2g_bench_c3.json 2> $O/r2g_bench_c3.err
timeout 900 python bench.py > $O/r2g_bench_default.json 2> $O/r2g_bench_default.err
echo done
