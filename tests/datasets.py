"""Loaders for the committed real-data fixtures (tests/golden/make_datasets.py)."""
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NA = np.iinfo(np.int32).min


def unvotes100():
    v = np.load(os.path.join(HERE, "unvotes100.npz"))["V"].astype(np.int32)
    v[v < 0] = NA
    return np.asfortranarray(v)


def mnistbinary_2000():
    d = np.load(os.path.join(HERE, "mnistbinary_2000.npz"))
    F, N = (int(x) for x in d["shape"])
    bits = np.unpackbits(d["bits"], axis=1, bitorder="little")[:, :F]
    return np.asfortranarray(bits.T.astype(np.int32))


def mnistbinary_full():
    """all of data/mnistbinary: 784 x 60000 (BASELINE configs[2] at its size)"""
    d = np.load(os.path.join(HERE, "mnistbinary_full.npz"))
    F, N = (int(x) for x in d["shape"])
    bits = np.unpackbits(d["bits"], axis=1, bitorder="little")[:, :F]
    return np.asfortranarray(bits.T.astype(np.int32))


def hold_out(V, frac, seed):
    """BASELINE config 2: hide `frac` of the observed entries (the experiments use
    runif(1) > 0.75, fig10:57; BASELINE.json says 20 %).  Returns (train, test)."""
    rng = np.random.default_rng(seed)
    obs = V != NA
    test = obs & (rng.random(V.shape) < frac)
    tr = V.copy(order="F"); tr[test] = NA
    te = np.full_like(V, NA); te[test] = V[test]
    return np.asfortranarray(tr), np.asfortranarray(te)
