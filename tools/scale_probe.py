"""torchrun tools/scale_probe.py [noop]: per-rank phase times of the sharded C5 iteration."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbmf_b200 import Engine
from nbmf_b200.dist import init_nccl, make_allreduce_hook, shard_columns
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
init_nccl(dev)
F = N = 262144; K = 128
lo, hi = shard_columns(N, rank, world)
e = Engine(device=local, n_global=N, n_offset=lo)
e.generate_synthetic(F, hi - lo, 8, 1.0, 1.0, 5)
mode = sys.argv[1] if len(sys.argv) > 1 else "nccl"
if mode == "noop":
    e.set_allreduce_hook(lambda p, c, s: None)
else:
    e.set_allreduce_hook(make_allreduce_hook(dev))
e.init_random_labels(K, 6)
e.vb_begin(K, 1.0, 1.0, 1.0 / K)
for _ in range(4): e.vb_step(False)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(6): e.vb_step(False)
e.synchronize(); dt = (time.perf_counter() - t0) / 6
e.vb_end(0, want_outputs=False)
t = e.timing()
msg = f"[{mode}] rank {rank}: wall/iter {1e3*dt:.2f} ms rows {t['pass_rows_ms']:.2f} exch {t['exchange_ms']:.2f} cols {t['pass_cols_ms']:.2f} params {t['params_ms']:.2f}"
outs = [None] * world
dist.all_gather_object(outs, msg)
if rank == 0: print("\n".join(outs), flush=True)
e.close(); dist.destroy_process_group()
