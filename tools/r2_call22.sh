#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 python bench.py --workload c4 --precision 3 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c22_c4_fast_main.json 2> $O/r2_c22_c4_fast_main.err
NBMF_LIB_VARIANT=epi88 timeout 300 python bench.py --workload c4 --precision 3 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c22_c4_fast_epi88.json 2> $O/r2_c22_c4_fast_epi88.err
echo done
