#!/bin/bash
# 8-GPU call with the round's final kernels: sharded == unsharded, C5 scaling line (AUTO and hi + e5m2 lo)
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
G=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 600 $T --master-port 29602 tools/dist_check.py > $O/r2f_dist_check_g$G.log 2>&1
timeout 900 $T --master-port 29603 bench.py --gpus $G --steps 20 --warmup 5 > $O/r2f_bench_c5_g$G.json 2> $O/r2f_bench_c5_g$G.err
timeout 900 $T --master-port 29605 bench.py --gpus $G --steps 20 --warmup 5 --precision 4 --no-e2e > $O/r2f_bench_c5_g${G}_lo8.json 2> $O/r2f_bench_c5_g${G}_lo8.err
echo done
