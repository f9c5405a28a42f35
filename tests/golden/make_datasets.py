"""Exports the two real datasets BASELINE.json's configs name into small fixtures, with the
R-free reader nbmf_b200/rdata.py (no R in the image; /root/reference is not on the GPU box):
  unvotes100.npz       -- data/unvotes100.rda, 100 x 5429, int8, -1 = NA            (config 2)
  mnistbinary_2000.npz -- data/mnistbinary.rda, first 2000 of 60000 columns, bit-packed
                          (config 3; the reference itself can only hold ~2000 columns)
Run in the build container:  python tests/golden/make_datasets.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nbmf_b200 import rdata  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
u = rdata.read_rda("/root/reference/data/unvotes100.rda")["unvotes100"]
assert u.shape == (100, 5429) and int(np.isnan(u).sum()) == 50614
u8 = np.where(np.isnan(u), -1, u).astype(np.int8)
np.savez_compressed(os.path.join(HERE, "unvotes100.npz"), V=u8)
m = rdata.read_rda("/root/reference/data/mnistbinary.rda")["mnistbinary"]
assert m.shape == (784, 60000) and not np.isnan(m).any()
sub = np.ascontiguousarray(m[:, :2000].T.astype(np.uint8))          # [n][f]
np.savez_compressed(os.path.join(HERE, "mnistbinary_2000.npz"), bits=np.packbits(sub, axis=1, bitorder="little"),
                    shape=np.array([784, 2000]), density_full=np.array([m.mean()]))
print("unvotes100 mean", np.nanmean(u), "mnist density", m.mean(), "subset", sub.mean())
# all 60000 columns (BASELINE configs[2] at its size; 2 MB): tests/tools/config3_gibbs.py
full = np.ascontiguousarray(m.T.astype(np.uint8))
np.savez_compressed(os.path.join(HERE, "mnistbinary_full.npz"), bits=np.packbits(full, axis=1, bitorder="little"),
                    shape=np.array([784, 60000]))
