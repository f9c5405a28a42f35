"""R-free reader for the reference's datasets (`data/*.rda`): bzip2/gzip/xz-compressed
`RDX2` XDR serialisations holding one numeric/integer matrix with `dim` (+`dimnames`)
attributes.  Returns column-major NumPy arrays; NA -> NaN for
doubles, INT_MIN for integers, exactly as R stores them.

Only the SEXP types those files contain are implemented (NILVALUE, REFSXP, SYMSXP, LISTSXP,
CHARSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP).
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct

import numpy as np

NA_INTEGER = -2147483648


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.p = 0
        self.refs = []

    def i32(self) -> int:
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def raw(self, n: int) -> bytes:
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if t == 254:   # NILVALUE
            return None
        if t == 255:   # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 1:     # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 2:     # LISTSXP (pairlist) -> list of (tag, value)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                del attr
                nflags = struct.unpack_from(">i", self.b, self.p)[0]
                nt = nflags & 0xFF
                if nt == 254:
                    self.p += 4
                    return out
                if nt != 2:
                    cdr = self.item()
                    out.append((None, cdr))
                    return out
                self.p += 4
                has_attr = bool(nflags & 0x200)
                has_tag = bool(nflags & 0x400)
        if t == 9:     # CHARSXP
            n = self.i32()
            return None if n == -1 else self.raw(n).decode("utf-8", "replace")
        if t in (10, 13):  # LGLSXP, INTSXP
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int32)
            self.p += 4 * n
            return self._attrs(v, has_attr)
        if t == 14:    # REALSXP
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
            self.p += 8 * n
            return self._attrs(v, has_attr)
        if t == 16:    # STRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
            return self._attrs(v, has_attr)
        if t == 19:    # VECSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
            return self._attrs(v, has_attr)
        raise NotImplementedError(f"SEXP type {t} at offset {self.p}")

    def _attrs(self, v, has_attr):
        if not has_attr:
            return v
        attrs = dict((k, x) for k, x in (self.item() or []))
        if isinstance(v, np.ndarray) and "dim" in attrs:
            v = v.reshape(tuple(int(d) for d in attrs["dim"]), order="F")
            v = np.asfortranarray(v)
        if isinstance(v, np.ndarray):
            return _Tagged(v, attrs)
        return v


class _Tagged:
    def __init__(self, array, attrs):
        self.array = array
        self.attrs = attrs


def _decompress(raw: bytes) -> bytes:
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


def read_rda(path: str) -> dict:
    """{object name: ndarray (column-major)} for every matrix/vector in the .rda file."""
    buf = _decompress(open(path, "rb").read())
    if buf[:5] != b"RDX2\n" or buf[5:7] != b"X\n":
        raise ValueError("not an RDX2 XDR .rda file")
    r = _Reader(buf)
    r.p = 7
    r.i32(); r.i32(); r.i32()   # format version, writer R version, min R version
    top = r.item()
    out = {}
    for tag, val in top or []:
        if isinstance(val, _Tagged):
            out[tag] = val.array
        elif isinstance(val, np.ndarray):
            out[tag] = val
    return out


def to_imat(a: np.ndarray) -> np.ndarray:
    """What the Rcpp glue does to an R matrix passed as `const arma::imat&`: int32,
    NA_real_/NA_integer_ -> NA_INTEGER (src/RcppExports.cpp:262)."""
    if a.dtype.kind == "f":
        out = np.full(a.shape, NA_INTEGER, dtype=np.int32)
        m = ~np.isnan(a)
        out[m] = a[m].astype(np.int32)
        return np.asfortranarray(out)
    return np.asfortranarray(a.astype(np.int32))
