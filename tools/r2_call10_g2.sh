#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_glue.py -m gpu -q --timeout=300 -s -k "group or glue" > $O/r2_tests10.log 2>&1
echo "pytest rc=$?" >> $O/r2_tests10.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > $O/r2_dist_check_g2b.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > $O/r2_bench_c5_g2b.json 2> $O/r2_bench_c5_g2b.err
echo done
