#!/bin/bash
# 8-GPU call: aggregate host->device bandwidth, sharded == unsharded, C5 scaling line with and without row slicing
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
G=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/r2_topo_g$G.log 2>&1
timeout 300 $T --master-port 29601 tools/h2d_probe.py 1.2 > $O/r2_h2d_probe_g$G.log 2>&1
timeout 600 $T --master-port 29602 tools/dist_check.py > $O/r2_dist_check_g$G.log 2>&1
timeout 900 $T --master-port 29603 bench.py --gpus $G --steps 20 --warmup 5 > $O/r2_bench_c5_g$G.json 2> $O/r2_bench_c5_g$G.err
NBMF_ROW_SLICING=0 timeout 900 $T --master-port 29604 bench.py --gpus $G --steps 20 --warmup 5 --no-e2e > $O/r2_bench_c5_g${G}_noslice.json 2> $O/r2_bench_c5_g${G}_noslice.err
timeout 900 $T --master-port 29605 bench.py --gpus $G --steps 20 --warmup 5 --precision 4 --no-e2e > $O/r2_bench_c5_g${G}_lo8.json 2> $O/r2_bench_c5_g${G}_lo8.err
echo done
