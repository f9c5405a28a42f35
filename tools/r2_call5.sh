#!/bin/bash
# 2-GPU call: group / communicator tests, sharded == unsharded, scaling of bench at N=2
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi topo -m > $O/r2_topo.log 2>&1
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_glue.py tests/test_gpu_entrypoints.py tests/test_gpu_tensor.py -m gpu -q --timeout=600 -s > $O/r2_tests5.log 2>&1
echo "pytest rc=$?" >> $O/r2_tests5.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > $O/r2_dist_check_g2.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py hook > $O/r2_dist_check_g2_hook.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_bench_c5_g2.json 2> $O/r2_bench_c5_g2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --exchange hook --no-e2e > $O/r2_bench_c5_g2_hook.json 2> $O/r2_bench_c5_g2_hook.err
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2_bench_c5_g1.json 2> $O/r2_bench_c5_g1.err
timeout 600 python -m pytest tests/test_gpu_horizon.py -m gpu -q --timeout=600 -s > $O/r2_tests5_horizon.log 2>&1
echo done
