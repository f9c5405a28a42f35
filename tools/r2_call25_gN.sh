#!/bin/bash
# C5 scaling line at G GPUs with the round's final kernels
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
G=${1:-4}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 900 $T --master-port 29613 bench.py --gpus $G --steps 20 --warmup 5 > $O/r2f_bench_c5_g$G.json 2> $O/r2f_bench_c5_g$G.err
echo done
