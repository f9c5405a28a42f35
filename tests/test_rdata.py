"""The R-free .rda reader (nbmf_b200/rdata.py): a hand-built RDX2 stream, and the reference's
own data files when they are present (build container)."""
import bz2
import os
import struct

import numpy as np
import pytest

from nbmf_b200 import rdata
from tests import datasets


def _i(x):
    return struct.pack(">i", x)


def _charsxp(s):
    b = s.encode()
    return _i(9 | (1 << 16)) + _i(len(b)) + b  # CHARSXP with an encoding flag in the gp bits


def _sym(s):
    return _i(1) + _charsxp(s)


def make_rdx2(name, mat):
    """Minimal writer: pairlist{ tag=name : REALSXP with attributes pairlist{dim: INTSXP} }."""
    F, N = mat.shape
    body = b"RDX2\nX\n" + _i(2) + _i(0x030600) + _i(0x020300)
    body += _i(2 | 0x400)                       # LISTSXP with tag
    body += _sym(name)
    body += _i(14 | 0x200) + _i(F * N) + np.asfortranarray(mat).astype(">f8").tobytes(order="F")
    body += _i(2 | 0x400) + _sym("dim") + _i(13) + _i(2) + _i(F) + _i(N) + _i(254)   # attributes
    body += _i(254)
    return bz2.compress(body)


def test_reader_on_handbuilt_stream(tmp_path):
    m = np.arange(12, dtype=np.float64).reshape(3, 4)
    m[1, 2] = np.nan
    p = tmp_path / "x.rda"
    p.write_bytes(make_rdx2("toy", m))
    d = rdata.read_rda(str(p))
    assert list(d) == ["toy"] and d["toy"].shape == (3, 4)
    assert np.array_equal(d["toy"], m, equal_nan=True)
    im = rdata.to_imat(d["toy"])
    assert im.dtype == np.int32 and im[1, 2] == rdata.NA_INTEGER and im[2, 3] == 11


def test_fixtures_have_the_reference_shapes():
    u = datasets.unvotes100()
    assert u.shape == (100, 5429) and int((u == datasets.NA).sum()) == 50614
    assert abs(u[u != datasets.NA].mean() - 0.773) < 1e-3
    m = datasets.mnistbinary_2000()
    assert m.shape == (784, 2000) and set(np.unique(m)) == {0, 1}
    tr, te = datasets.hold_out(u, 0.2, 2)
    n_obs = (u != datasets.NA).sum()
    assert abs((te != datasets.NA).sum() / n_obs - 0.2) < 0.01
    assert ((tr != datasets.NA) & (te != datasets.NA)).sum() == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/data"), reason="reference data not present")
def test_reader_on_reference_files():
    d = rdata.read_rda("/root/reference/data/unvotes100.rda")["unvotes100"]
    assert np.array_equal(rdata.to_imat(d), datasets.unvotes100())
    a = rdata.read_rda("/root/reference/data/animals.rda")["animals"]
    assert a.shape == (50, 85) and a.dtype == np.int32 and abs(a.mean() - 0.368) < 1e-3
    z = rdata.read_rda("/root/reference/data/zachary.rda")["zachary"]
    assert z.shape == (34, 34)
