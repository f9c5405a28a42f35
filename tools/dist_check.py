"""torchrun --nproc-per-node G tools/dist_check.py [hook] : column-sharded VB equals the single-GPU run.
Default: the library's own NCCL communicator (nbmf_comm_init_rank; torch.distributed only ships the id);
`hook`: the all-reduce hook served by torch.distributed."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbmf_b200 import Engine, _lib  # noqa: E402
from nbmf_b200.dist import combine_lowerbounds, init_library_comm, init_nccl, make_allreduce_hook, shard_columns  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
init_nccl(dev)
USE_HOOK = len(sys.argv) > 1 and sys.argv[1] == "hook"


def attach(e):
    if USE_HOOK:
        e.set_allreduce_hook(make_allreduce_hook(dev))
    else:
        init_library_comm(e, rank, world)


rel = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)
for (F, N, K, prec, iters) in [(300, 1000, 5, _lib.PREC_FP64, 12), (2048, 4096, 128, _lib.PREC_TENSOR, 8), (1500, 3000, 64, _lib.PREC_AUTO, 6),
                               (4096, 8192, 128, _lib.PREC_TENSOR_LO8, 5), (4096, 8192, 128, _lib.PREC_TENSOR_FAST, 5)]:
    lo, hi = shard_columns(N, rank, world)
    e = Engine(device=local, precision=prec, n_global=N, n_offset=lo)
    e.generate_synthetic(F, hi - lo, 4, 1.0, 1.0, 21)
    attach(e)
    e.init_random_labels(K, 22)
    W, H, _, _ = e.vb_aspect(K, 1.0, 1.0, 0.5, iter=iters, checklb=True)
    terms = e.vb_lb_terms(iters)

    def ar(x):
        t = torch.tensor(x, device=dev, dtype=torch.float64); dist.all_reduce(t); return t.cpu().numpy()
    lb = combine_lowerbounds(terms, ar)
    Hs = [torch.zeros((N * (r + 1) // world - N * r // world, K), dtype=torch.float64, device=dev) for r in range(world)]
    dist.all_gather(Hs, torch.tensor(np.ascontiguousarray(H), device=dev))
    Hall = torch.cat(Hs).cpu().numpy()
    e.close()
    if rank == 0:
        e1 = Engine(device=local, precision=prec)
        e1.generate_synthetic(F, N, 4, 1.0, 1.0, 21)
        e1.init_random_labels(K, 22)
        W1, H1, lb1, _ = e1.vb_aspect(K, 1.0, 1.0, 0.5, iter=iters, checklb=True)
        e1.close()
        print(f"F={F} N={N} K={K} prec={prec} world={world} exchange={'hook' if USE_HOOK else 'library nccl'}: relF E_W {rel(W, W1):.2e} E_H {rel(Hall, H1):.2e} "
              f"ELBO {np.max(np.abs(lb - lb1) / np.abs(lb1)):.2e}", flush=True)
        assert rel(W, W1) < 5e-6 and rel(Hall, H1) < 5e-6 and np.max(np.abs(lb - lb1) / np.abs(lb1)) < 5e-6
    dist.barrier()
# column-sharded Gibbs: counter-based draws keyed by (f, global n, sweep) => the chain does not
# depend on the sharding; E_W must be identical, E_H the matching block
F, N, K = 200, 900, 8
lo, hi = shard_columns(N, rank, world)
e = Engine(device=local, n_global=N, n_offset=lo)
e.generate_synthetic(F, hi - lo, 4, 0.7, 0.7, 31)
attach(e)
e.init_random_labels(K, 5)
Wg, Hg, _, _ = e.gibbs_dirber(K, 1.0, 1.0, 0.5, iter=60, burnin=0.5, seed=13)
e.close()
if rank == 0:
    e1 = Engine(device=local)
    e1.generate_synthetic(F, N, 4, 0.7, 0.7, 31)
    e1.init_random_labels(K, 5)
    W1, H1, _, _ = e1.gibbs_dirber(K, 1.0, 1.0, 0.5, iter=60, burnin=0.5, seed=13)
    e1.close()
    print(f"sharded Gibbs world={world}: max|dE_W| {np.max(np.abs(Wg - W1)):.2e} max|dE_H block| {np.max(np.abs(Hg - H1[lo:hi])):.2e}", flush=True)
    assert np.allclose(Wg, W1, rtol=0, atol=1e-12) and np.allclose(Hg, H1[lo:hi], rtol=0, atol=1e-12)
dist.barrier()
if rank == 0:
    print("dist_check ok")
dist.destroy_process_group()
