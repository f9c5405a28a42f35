"""world_size-2 `gloo` tests (CPU) of the column-sharding host logic:
the shard plan, the exchange of the F x K Dirichlet-side statistic and the ELBO combination.
The per-rank compute is the oracle's batch form standing in for the engine (no GPU here); the
GPU path itself is checked on two B200s by tools/dist_check.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_plan_covers_all_columns():
    from nbmf_b200.dist import shard_columns
    for N in (1, 7, 300, 5429, 262144):
        for w in (1, 2, 3, 8):
            blocks = [shard_columns(N, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == N
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_columns(10, 2, 2)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nbmf_b200.dist import combine_lowerbounds, shard_columns
    from oracle import oracle
    F, N, K, iters = 40, 66, 4, 6
    V, _, _ = oracle.generate(F, N, 3, 0.4, 0.4, 5)
    V = V.copy(order="F"); V[np.random.default_rng(1).random(V.shape) < 0.15] = oracle.NA
    Z = oracle.random_labels(V, K, 2)
    lo, hi = shard_columns(N, rank, world)
    Vl, Zl = np.asfortranarray(V[:, lo:hi]), np.asfortranarray(Z[:, lo:hi])

    def allreduce(x):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()
    # initial statistics: local histograms, the row statistic summed across ranks
    obs = Vl != oracle.NA
    nt = np.stack([((Zl == k) & obs).sum(1) for k in range(K)], 1).astype(float)
    fo = np.stack([((Zl == k) & obs & (Vl == 1)).sum(0) for k in range(K)], 1).astype(float)
    fz = np.stack([((Zl == k) & obs & (Vl == 0)).sum(0) for k in range(K)], 1).astype(float)
    nt = allreduce(nt)
    lbs = np.zeros((iters, 3))
    import scipy.special as sp
    for i in range(iters):
        # one sharded iteration = oracle batch step on the local columns with the replicated
        # row statistic, then the exchange (the engine's allreduce hook) of its partial update
        r = oracle.vb_aspect_batch(Vl, None, K, 1.0, 1.0, 0.7, iter=1, checklb=True, stats=(nt, fo, fz))
        nt_partial, fo, fz = r["stats"]
        g = 0.7 + nt
        elw = sp.digamma(g) - sp.digamma(g.sum(1, keepdims=True))
        row_term = F * (sp.gammaln(K * 0.7) - K * sp.gammaln(0.7)) + (0.7 - 1) * elw.sum() - \
            (sp.gammaln(g.sum(1)).sum() - sp.gammaln(g).sum() + ((g - 1) * elw).sum())
        lbs[i] = [r["lowerbounds"][0] - row_term, row_term, 0.0]   # local part, replicated part
        E_W = r["E_W"]; E_Hl = r["E_H"]
        nt = allreduce(nt_partial)
    lb = combine_lowerbounds(lbs, allreduce)
    Hs = [torch.zeros((shard_columns(N, r_, world)[1] - shard_columns(N, r_, world)[0], K), dtype=torch.float64)
          for r_ in range(world)]
    dist.all_gather(Hs, torch.from_numpy(np.ascontiguousarray(E_Hl)))
    if rank == 0:
        full = oracle.vb_aspect_batch(V, Z, K, 1.0, 1.0, 0.7, iter=iters, checklb=True)
        E_H = torch.cat(Hs).numpy()
        out["dW"] = float(np.max(np.abs(E_W - full["E_W"])))
        out["dH"] = float(np.max(np.abs(E_H - full["E_H"])))
        out["dlb"] = float(np.max(np.abs(lb - full["lowerbounds"]) / np.abs(full["lowerbounds"])))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_iteration_equals_unsharded_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["dW"] < 1e-12 and out["dH"] < 1e-12 and out["dlb"] < 1e-12, dict(out)
