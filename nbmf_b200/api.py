"""Host-side mirror of the reference's operator interface for the hot path.

Same names, argument meaning and error behaviour as the Rcpp functions and their
R wrappers, so the parity tests read like calls into the reference:

    VB_Aspect_Rcpp            src/collapsed_gibbs_z_VB.cpp:373-526
    VB_DirBer_Rcpp            src/collapsed_gibbs_z_VB.cpp:97-168
    Gibbs_DirBer_finite_Rcpp  src/collapsed_gibbs_z.cpp:363-437 (+ :565-605)
    lower_bound_aspect        src/collapsed_gibbs_z_VB.cpp:339-369
    VB_aspectRcpp / VB_DirBer / Gibbs_DirBer   R/generic_*.R (0-basing of Z, checks)
    BernoulliNMF              R/BernoulliNMF.R:12-106
    loglikelihood / expectation   R/commons.R:2-33

All compute goes through the C ABI of libnbmf_cuda (include/nbmf_cuda.h); there
is no NumPy/PyTorch/CPU implementation of the path in this package.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import NbmfError, NbmfOpts, NbmfTiming

NA_INTEGER = np.iinfo(np.int32).min
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint32)
_lp = C.POINTER(C.c_int64)


def as_imat(a) -> np.ndarray:
    """What `Rcpp::traits::input_parameter<const arma::imat&>` does to an R
    matrix (RcppExports.cpp:262): REAL -> int32 copy with NA_real_ -> NA_INTEGER,
    column-major."""
    a = np.asarray(a)
    if a.dtype.kind == "f":
        out = np.full(a.shape, NA_INTEGER, dtype=np.int32)
        m = ~np.isnan(a)
        out[m] = a[m].astype(np.int32)
        a = out
    return np.asfortranarray(a, dtype=np.int32)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class Engine:
    """One libnbmf_cuda handle = one GPU's share of V (all of it unless sharded)."""

    def __init__(self, device: int = 0, precision: int = _lib.PREC_AUTO, flush_interval: int = 0,
                 n_global: int = 0, n_offset: int = 0, pipe_block_cols: int = 0, planned_iters: int = 0):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        opts = NbmfOpts(device, precision, flush_interval, pipe_block_cols, n_global, n_offset, planned_iters, 0)
        rc = self.lib.nbmf_create(C.byref(self.h), C.byref(opts))
        if rc != 0:
            raise NbmfError(rc, "nbmf_create failed (no CUDA device? there is no CPU fallback)")
        self.F = self.N = 0
        self._hook_ref = None

    # -- plumbing --------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise NbmfError(rc, self.lib.nbmf_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.lib.nbmf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return int(self.lib.nbmf_stream(self.h) or 0)

    def synchronize(self):
        self._ck(self.lib.nbmf_synchronize(self.h))

    def timing(self) -> dict:
        t = NbmfTiming()
        self._ck(self.lib.nbmf_get_timing(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in NbmfTiming._fields_}

    def set_allreduce_hook(self, fn):
        """fn(device_ptr:int, count:int, stream:int) -> None, sums fp64 in place across ranks."""
        if fn is None:
            self._hook_ref = None
            self._ck(self.lib.nbmf_set_allreduce_hook(self.h, C.cast(None, _lib.ALLREDUCE_FN), None))
            return

        def tramp(ctx, buf, count, stream):
            try:
                fn(int(buf or 0), int(count), int(stream or 0))
                return 0
            except Exception:  # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1

        self._hook_ref = _lib.ALLREDUCE_FN(tramp)
        self._ck(self.lib.nbmf_set_allreduce_hook(self.h, self._hook_ref, None))

    # -- collectives owned by the library (one process per GPU) ----------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """128 bytes rank 0 ships to the other ranks (nbmf_comm_unique_id)."""
        buf = C.create_string_buffer(128)
        rc = _lib.load().nbmf_comm_unique_id(buf)
        if rc != 0:
            raise NbmfError(rc, "nbmf_comm_unique_id failed")
        return buf.raw

    def comm_init(self, id128: bytes, world: int, rank: int):
        buf = C.create_string_buffer(bytes(id128), 128)
        self._ck(self.lib.nbmf_comm_init_rank(self.h, buf, world, rank))

    def comm_destroy(self):
        self._ck(self.lib.nbmf_comm_destroy(self.h))

    def comm_info(self):
        r, w = C.c_int(0), C.c_int(1)
        self._ck(self.lib.nbmf_comm_info(self.h, C.byref(r), C.byref(w)))
        return r.value, w.value

    # -- data ------------------------------------------------------------
    def set_data(self, V):
        V = as_imat(V)
        self.F, self.N = V.shape
        self._ck(self.lib.nbmf_set_data_i32(self.h, _ptr(V, _ip), self.F, self.N))

    def set_data_bits(self, vbits, mbits, F, N):
        vbits = np.ascontiguousarray(vbits, dtype=np.uint32)
        mb = np.ascontiguousarray(mbits, dtype=np.uint32) if mbits is not None else None
        self.F, self.N = F, N
        self._ck(self.lib.nbmf_set_data_bits(self.h, _ptr(vbits, _up), _ptr(mb, _up), F, N))

    def generate_synthetic(self, F, N, Ktrue, a, b, seed):
        self.F, self.N = F, N
        self._ck(self.lib.nbmf_generate_synthetic(self.h, F, N, Ktrue, a, b, seed))

    def get_data_bits(self, want_mask=True):
        fw = (self.F + 31) // 32
        v = np.zeros((self.N, fw), dtype=np.uint32)
        m = np.zeros((self.N, fw), dtype=np.uint32) if want_mask else None
        self._ck(self.lib.nbmf_get_data_bits(self.h, _ptr(v, _up), _ptr(m, _up)))
        return v, m

    def get_data(self) -> np.ndarray:
        """V back as an int32 F x N matrix (NA = INT_MIN), from the device bit planes."""
        v, m = self.get_data_bits()
        vb = np.unpackbits(v.view(np.uint8), axis=1, bitorder="little")[:, :self.F]
        mb = np.unpackbits(m.view(np.uint8), axis=1, bitorder="little")[:, :self.F]
        out = np.where(mb == 1, vb.astype(np.int32), NA_INTEGER).astype(np.int32)
        return np.asfortranarray(out.T)

    def dims(self):
        F, N, nobs = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self.lib.nbmf_get_dims(self.h, C.byref(F), C.byref(N), C.byref(nobs)))
        return F.value, N.value, nobs.value

    # -- statistics ------------------------------------------------------
    def init_from_labels(self, Z, Kc):
        Z = as_imat(Z)
        if Z.shape != (self.F, self.N):
            raise ValueError("Z and V dimensions differ")
        self._ck(self.lib.nbmf_init_from_labels(self.h, _ptr(Z, _ip), Kc))

    def init_random_labels(self, Kc, seed):
        self._ck(self.lib.nbmf_init_random_labels(self.h, Kc, seed))

    def set_stats(self, n_total, f_ones, f_zeros):
        nt = np.asfortranarray(n_total, dtype=np.float64)
        fo = np.asfortranarray(f_ones, dtype=np.float64)
        fz = np.asfortranarray(f_zeros, dtype=np.float64)
        self._ck(self.lib.nbmf_set_stats(self.h, nt.shape[1], _ptr(nt, _dp), _ptr(fo, _dp), _ptr(fz, _dp)))

    def get_stats(self, Kc):
        nt = np.zeros((self.F, Kc), order="F"); fo = np.zeros((self.N, Kc), order="F")
        fz = np.zeros((self.N, Kc), order="F")
        self._ck(self.lib.nbmf_get_stats(self.h, Kc, _ptr(nt, _dp), _ptr(fo, _dp), _ptr(fz, _dp)))
        return nt, fo, fz

    # -- VB --------------------------------------------------------------
    def vb_aspect(self, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, checklb=False, return_E_Z=False, out=None):
        """`out=(E_W, E_H)`: caller-owned column-major fp64 buffers (e.g. pinned memory)."""
        if out is not None:
            E_W, E_H = out
            assert E_W.shape == (self.F, K) and E_H.shape == (self.N, K) and E_W.flags.f_contiguous and E_H.flags.f_contiguous
        else:
            E_W = np.zeros((self.F, K), order="F"); E_H = np.zeros((self.N, K), order="F")
        lbs = np.zeros(max(iter, 1))
        E_Z = np.zeros((self.F, K, self.N), order="F") if return_E_Z else None
        self._ck(self.lib.nbmf_vb_aspect(self.h, K, alpha, beta, gamma, iter, int(checklb), _ptr(E_W, _dp),
                                         _ptr(E_H, _dp), _ptr(lbs, _dp), _ptr(E_Z, _dp)))
        return E_W, E_H, lbs[:iter], E_Z

    def vb_aspect_bits(self, vbits, F, N, K, stats, alpha=1.0, beta=1.0, gamma=1.0, iter=100, checklb=False, out=None):
        """VB_Aspect_Rcpp from host memory in one call (nbmf_vb_aspect_bits): `vbits` the column-packed
        bit plane [N][ceil(F/32)] (no NA), `stats` = (n_total F x K, f_ones N x K, f_zeros N x K)
        column-major fp64.  The upload overlaps the first iteration when `vbits` is pinned.  With a
        library communicator the row-side arrays belong to rank 0: the other ranks neither upload
        n_total (it may be None) nor download E_W (returned as None)."""
        vbits = np.ascontiguousarray(vbits, dtype=np.uint32)
        assert vbits.shape == (N, (F + 31) // 32)
        rank, world = self.comm_info()
        root = world <= 1 or rank == 0
        nt, fo, fz = stats
        nt = np.asfortranarray(nt, dtype=np.float64) if (root or nt is not None) else None
        fo = np.asfortranarray(fo, dtype=np.float64); fz = np.asfortranarray(fz, dtype=np.float64)
        assert (nt is None or nt.shape == (F, K)) and fo.shape == (N, K) and fz.shape == (N, K)
        self.F, self.N = F, N
        if out is not None:
            E_W, E_H = out
            assert (E_W is None or (E_W.shape == (F, K) and E_W.flags.f_contiguous)) and E_H.shape == (N, K) and E_H.flags.f_contiguous
        else:
            E_W = np.zeros((F, K), order="F"); E_H = np.zeros((N, K), order="F")
        if not root:
            nt = None; E_W = None
        lbs = np.zeros(max(iter, 1))
        self._ck(self.lib.nbmf_vb_aspect_bits(self.h, _ptr(vbits, _up), F, N, K, alpha, beta, gamma, iter, int(checklb),
                                              _ptr(nt, _dp), _ptr(fo, _dp), _ptr(fz, _dp), _ptr(E_W, _dp), _ptr(E_H, _dp),
                                              _ptr(lbs, _dp)))
        return E_W, E_H, lbs[:iter]

    def vb_begin(self, K, alpha=1.0, beta=1.0, gamma=1.0, mean_params=False):
        self._vbK = K
        self._ck(self.lib.nbmf_vb_begin(self.h, K, alpha, beta, gamma, int(mean_params)))

    def vb_step(self, checklb=False):
        self._ck(self.lib.nbmf_vb_step(self.h, int(checklb)))

    def vb_end(self, n_lb=0, want_outputs=True):
        K = self._vbK
        E_W = np.zeros((self.F, K), order="F") if want_outputs else None
        E_H = np.zeros((self.N, K), order="F") if want_outputs else None
        lbs = np.zeros(max(n_lb, 1))
        self._ck(self.lib.nbmf_vb_end(self.h, _ptr(E_W, _dp), _ptr(E_H, _dp), _ptr(lbs, _dp), n_lb, None))
        return E_W, E_H, lbs[:n_lb]

    def vb_lb_terms(self, n_iter):
        t = np.zeros((max(n_iter, 1), 3))
        self._ck(self.lib.nbmf_vb_get_lb_terms(self.h, _ptr(t, _dp), n_iter))
        return t[:n_iter]

    def argmax_labels(self, want_host=False):
        """which.max(E_Z[f,,n]) of the last VB iteration, kept on the device (BernoulliNMF.R:45-52)."""
        Z = np.zeros((self.F, self.N), dtype=np.int32, order="F") if want_host else None
        self._ck(self.lib.nbmf_vb_argmax_labels(self.h, _ptr(Z, _ip)))
        return Z

    def lower_bound(self, K, alpha=1.0, beta=1.0, gamma=1.0):
        lb = C.c_double(0.0)
        terms = np.zeros(7)
        self._ck(self.lib.nbmf_lower_bound(self.h, K, alpha, beta, gamma, C.byref(lb), _ptr(terms, _dp)))
        return lb.value, terms

    def vb_dirber(self, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, mode=_lib.DIRBER_EXACT_WAVEFRONT,
                  return_E_Z=False):
        Kc = K + 1
        E_W = np.zeros((self.F, Kc), order="F"); E_H = np.zeros((self.N, Kc), order="F")
        E_Z = np.zeros((self.F, Kc, self.N), order="F") if return_E_Z else None
        self._ck(self.lib.nbmf_vb_dirber(self.h, K, alpha, beta, gamma, iter, mode, _ptr(E_W, _dp),
                                         _ptr(E_H, _dp), _ptr(E_Z, _dp)))
        return E_W, E_H, E_Z

    # -- Gibbs -----------------------------------------------------------
    def gibbs_dirber(self, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, burnin=0.5, seed=1, test_idx=None,
                     want_Z=False):
        E_W = np.zeros((self.F, K), order="F"); E_H = np.zeros((self.N, K), order="F")
        idx = np.ascontiguousarray(test_idx, dtype=np.int64) if test_idx is not None else None
        nt = 0 if idx is None else idx.size
        E_V = np.zeros(max(nt, 1))
        Z = np.zeros((self.F, self.N), dtype=np.int32, order="F") if want_Z else None
        self._ck(self.lib.nbmf_gibbs_dirber(self.h, K, alpha, beta, gamma, iter, burnin, seed, _ptr(E_W, _dp),
                                            _ptr(E_H, _dp), _ptr(idx, _lp), nt, _ptr(E_V, _dp), _ptr(Z, _ip)))
        return E_W, E_H, (E_V[:nt] if nt else None), Z

    def gibbs_dirber_collapsed(self, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, burnin=0.5, seed=1, want_samples=True):
        """The reference's own collapsed sweep (wavefront order, uniforms keyed by entry and sweep)."""
        ns = iter - int(np.floor(iter * burnin))
        Zs = np.zeros((self.F, self.N, max(ns, 1)), dtype=np.int32, order="F") if want_samples else None
        Zl = np.zeros((self.F, self.N), dtype=np.int32, order="F")
        nw = C.c_int(0)
        self._ck(self.lib.nbmf_gibbs_dirber_collapsed(self.h, K, alpha, beta, gamma, iter, burnin, seed, _ptr(Zs, _ip),
                                                      C.byref(nw), _ptr(Zl, _ip)))
        return (Zs[:, :, :ns] if want_samples else None), nw.value, Zl

    def gibbs_dirber_dp_collapsed(self, gamma=1.0, iter=100, burnin=0.5, seed=1, want_samples=True):
        """Gibbs_DirBer_DP_Rcpp's sweep (Kmax = 200; labels installed with init_from_labels(Z, 200))."""
        ns = iter - int(np.floor(iter * burnin))
        Zs = np.zeros((self.F, self.N, max(ns, 1)), dtype=np.int32, order="F") if want_samples else None
        Zl = np.zeros((self.F, self.N), dtype=np.int32, order="F")
        nw = C.c_int(0)
        self._ck(self.lib.nbmf_gibbs_dirber_dp_collapsed(self.h, 1.0, 1.0, gamma, iter, burnin, seed, _ptr(Zs, _ip),
                                                         C.byref(nw), _ptr(Zl, _ip)))
        return (Zs[:, :, :ns] if want_samples else None), nw.value, Zl

    def gibbs_dirdir_collapsed(self, K, alpha=1.0, gamma=1.0, iter=100, burnin=0.5, seed=1, want_samples=True):
        """Gibbs_DirDir_finite_Rcpp's sweep: assignments Z and C."""
        ns = iter - int(np.floor(iter * burnin))
        mk = lambda: np.zeros((self.F, self.N, max(ns, 1)), dtype=np.int32, order="F")
        Zs, Cs = (mk(), mk()) if want_samples else (None, None)
        Zl = np.zeros((self.F, self.N), dtype=np.int32, order="F"); Cl = np.zeros_like(Zl, order="F")
        nw = C.c_int(0)
        self._ck(self.lib.nbmf_gibbs_dirdir_collapsed(self.h, K, alpha, gamma, iter, burnin, seed, _ptr(Zs, _ip),
                                                      _ptr(Cs, _ip), C.byref(nw), _ptr(Zl, _ip), _ptr(Cl, _ip)))
        if want_samples:
            Zs, Cs = Zs[:, :, :ns], Cs[:, :, :ns]
        return Zs, Cs, nw.value, Zl, Cl

    def gibbs_sweep_given_W(self, v_n, W, alpha=1.0, iter=100, burnin=0.5, seed=1):
        v = np.ascontiguousarray(v_n, dtype=np.int32)
        W = np.asfortranarray(W, dtype=np.float64)
        F, K = W.shape
        Cm = np.zeros((F, K), order="F")
        self._ck(self.lib.nbmf_gibbs_sweep_given_W(self.h, _ptr(v, _ip), _ptr(W, _dp), F, K, alpha, iter,
                                                   burnin, seed, _ptr(Cm, _dp)))
        return Cm

    def sample_gibbs_cpp(self, v_n, W, Cinit, alpha=1.0, iter=100, burnin=0.5, seed=1, as_coded=True):
        """sample_gibbs_cpp (sample_gibbs.cpp:22-80); as_coded=False stores the draw (R sample_gibbs_Z)."""
        v = np.ascontiguousarray(v_n, dtype=np.int32)
        W = np.asfortranarray(W, dtype=np.float64)
        Ci = np.asfortranarray(Cinit, dtype=np.int32)
        F, K = W.shape
        assert Ci.shape == (F, K) and v.shape == (F,)
        Cm = np.zeros((F, K), order="F")
        self._ck(self.lib.nbmf_sample_gibbs_cpp(self.h, _ptr(v, _ip), _ptr(W, _dp), _ptr(Ci, _ip), F, K, alpha, iter,
                                                burnin, seed, int(as_coded), _ptr(Cm, _dp)))
        return Cm

    # -- posterior-mean reducers / conditional draws of the collapsed samplers ----
    def compute_expectation_W(self, Z_samples, gamma, Kq=None):
        Zs = np.asfortranarray(Z_samples, dtype=np.int32)
        F, N, J = Zs.shape
        if Kq is None:
            Kq = int(Zs.max()) + 1           # Z_samples.max() + 1 (collapsed_gibbs_z.cpp:180)
        E = np.zeros((F, Kq), order="F")
        self._ck(self.lib.nbmf_compute_expectation_W(self.h, _ptr(Zs, _ip), F, N, J, gamma, Kq, _ptr(E, _dp)))
        return E

    def compute_expectation_H(self, Z_samples, V, alpha, beta, Kq=None):
        Zs = np.asfortranarray(Z_samples, dtype=np.int32)
        Vi = as_imat(V)
        F, N, J = Zs.shape
        if Kq is None:
            Kq = int(Zs.max()) + 1
        E = np.zeros((N, Kq), order="F")
        self._ck(self.lib.nbmf_compute_expectation_H(self.h, _ptr(Zs, _ip), _ptr(Vi, _ip), F, N, J, alpha, beta, Kq,
                                                     _ptr(E, _dp)))
        return E

    def samples_VWH_DirBer(self, V, Z_samples, K, gamma, alpha, beta, seed=1, want_V=True):
        Zs = np.asfortranarray(Z_samples, dtype=np.int32)
        Vi = as_imat(V)
        F, N, J = Zs.shape
        Ws = np.zeros((F, K, J), order="F"); Hs = np.zeros((N, K, J), order="F")
        Vs = np.zeros((F, N, J), dtype=np.int32, order="F") if want_V else None
        self._ck(self.lib.nbmf_samples_VWH_DirBer(self.h, _ptr(Vi, _ip), _ptr(Zs, _ip), F, N, J, K, gamma, alpha, beta,
                                                  seed, _ptr(Ws, _dp), _ptr(Hs, _dp), _ptr(Vs, _ip)))
        return {"W_samples": Ws, "H_samples": Hs, "V_samples": Vs}

    def lower_bound_aspect_full(self, E_Z, E_log_W, E_log_H, E_log_1_H, gamma_vb, alpha_vb, beta_vb, alpha, beta,
                                gamma, V):
        f = lambda a: np.asfortranarray(a, dtype=np.float64)
        E_Z, E_log_W, E_log_H, E_log_1_H, gamma_vb, alpha_vb, beta_vb = map(f, (E_Z, E_log_W, E_log_H, E_log_1_H,
                                                                                   gamma_vb, alpha_vb, beta_vb))
        Vi = as_imat(V)
        F, N = Vi.shape
        K = E_log_W.shape[1]
        assert E_Z.shape == (F, K, N)
        lb = C.c_double(0.0)
        t7 = np.zeros(7)
        self._ck(self.lib.nbmf_lower_bound_aspect_full(self.h, _ptr(E_Z, _dp), _ptr(E_log_W, _dp), _ptr(E_log_H, _dp),
                                                       _ptr(E_log_1_H, _dp), _ptr(gamma_vb, _dp), _ptr(alpha_vb, _dp),
                                                       _ptr(beta_vb, _dp), alpha, beta, gamma, _ptr(Vi, _ip), F, N, K,
                                                       C.byref(lb), _ptr(t7, _dp)))
        return lb.value, t7

    # -- measurement aids -----------------------------------------------------
    def rng_probe(self, which, a, b, seed, n):
        out = np.zeros(n)
        self._ck(self.lib.nbmf_debug_rng_probe(self.h, which, a, b, seed, n, _ptr(out, _dp)))
        return out

    def ffma_peak_tflops(self):
        t = C.c_double(0.0)
        self._ck(self.lib.nbmf_debug_ffma_peak(self.h, C.byref(t)))
        return t.value

    # -- predictive LL -----------------------------------------------------
    def predictive_ll(self, E_W, E_H, V_test):
        Vt = as_imat(V_test)
        E_W = np.asfortranarray(E_W, dtype=np.float64); E_H = np.asfortranarray(E_H, dtype=np.float64)
        F, N = Vt.shape
        ll = C.c_double(0.0); n = C.c_int64(0)
        self._ck(self.lib.nbmf_predictive_ll(self.h, _ptr(E_W, _dp), _ptr(E_H, _dp), E_W.shape[1],
                                             _ptr(Vt, _ip), F, N, C.byref(ll), C.byref(n)))
        return ll.value, n.value


class Group:
    """One process, several GPUs (nbmf_group_*): the columns of V sharded over `devices`, the F x K
    statistic exchanged by the library's own NCCL communicator.  Calls take and return whole matrices."""

    def __init__(self, devices, precision: int = _lib.PREC_AUTO, flush_interval: int = 0, pipe_block_cols: int = 0,
                 planned_iters: int = 0):
        self.lib = _lib.load()
        self.g = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        opts = NbmfOpts(0, precision, flush_interval, pipe_block_cols, 0, 0, planned_iters, 0)
        rc = self.lib.nbmf_group_create(C.byref(self.g), C.byref(opts), devs, len(devices))
        if rc != 0:
            raise NbmfError(rc, "nbmf_group_create failed")
        self.size = len(devices)
        self.F = self.N = 0

    def _ck(self, rc):
        if rc != 0:
            raise NbmfError(rc, self.lib.nbmf_group_last_error(self.g).decode())

    def close(self):
        if getattr(self, "g", None) and self.g:
            self.lib.nbmf_group_destroy(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def columns(self, rank):
        lo, hi = C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.nbmf_group_columns(self.g, rank, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def timing(self, rank=0) -> dict:
        t = NbmfTiming()
        h = self.lib.nbmf_group_handle(self.g, rank)
        rc = self.lib.nbmf_get_timing(h, C.byref(t))
        if rc != 0:
            raise NbmfError(rc, "nbmf_get_timing")
        return {k: getattr(t, k) for k, _ in NbmfTiming._fields_}

    def set_data(self, V):
        V = as_imat(V)
        self.F, self.N = V.shape
        self._ck(self.lib.nbmf_group_set_data_i32(self.g, _ptr(V, _ip), self.F, self.N))

    def generate_synthetic(self, F, N, Ktrue, a, b, seed):
        self.F, self.N = F, N
        self._ck(self.lib.nbmf_group_generate_synthetic(self.g, F, N, Ktrue, a, b, seed))

    def init_from_labels(self, Z, Kc):
        Z = as_imat(Z)
        self._ck(self.lib.nbmf_group_init_from_labels(self.g, _ptr(Z, _ip), Kc))

    def init_random_labels(self, Kc, seed):
        self._ck(self.lib.nbmf_group_init_random_labels(self.g, Kc, seed))

    def vb_begin(self, K, alpha=1.0, beta=1.0, gamma=1.0, mean_params=False):
        self._vbK = K
        self._ck(self.lib.nbmf_group_vb_begin(self.g, K, alpha, beta, gamma, int(mean_params)))

    def vb_step(self, checklb=False):
        self._ck(self.lib.nbmf_group_vb_step(self.g, int(checklb)))

    def synchronize(self):
        self._ck(self.lib.nbmf_group_synchronize(self.g))

    def vb_end(self, n_lb=0, want_outputs=True):
        K = self._vbK
        E_W = np.zeros((self.F, K), order="F") if want_outputs else None
        E_H = np.zeros((self.N, K), order="F") if want_outputs else None
        lbs = np.zeros(max(n_lb, 1))
        self._ck(self.lib.nbmf_group_vb_end(self.g, _ptr(E_W, _dp), _ptr(E_H, _dp), _ptr(lbs, _dp), n_lb))
        return E_W, E_H, lbs[:n_lb]

    def vb_aspect(self, V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, checklb=False):
        """VB_Aspect_Rcpp over all devices of the group, one call."""
        V = as_imat(V); Z = as_imat(Z)
        self.F, self.N = V.shape
        E_W = np.zeros((self.F, K), order="F"); E_H = np.zeros((self.N, K), order="F")
        lbs = np.zeros(max(iter, 1))
        self._ck(self.lib.nbmf_group_vb_aspect(self.g, _ptr(V, _ip), _ptr(Z, _ip), self.F, self.N, K, alpha, beta, gamma,
                                               iter, int(checklb), _ptr(E_W, _dp), _ptr(E_H, _dp), _ptr(lbs, _dp)))
        return E_W, E_H, lbs[:iter]

    def vb_aspect_bits(self, vbits, F, N, K, stats, alpha=1.0, beta=1.0, gamma=1.0, iter=100, checklb=False, out=None):
        vbits = np.ascontiguousarray(vbits, dtype=np.uint32)
        assert vbits.shape == (N, (F + 31) // 32)
        nt, fo, fz = (np.asfortranarray(x, dtype=np.float64) for x in stats)
        self.F, self.N = F, N
        if out is not None:
            E_W, E_H = out
        else:
            E_W = np.zeros((F, K), order="F"); E_H = np.zeros((N, K), order="F")
        lbs = np.zeros(max(iter, 1))
        self._ck(self.lib.nbmf_group_vb_aspect_bits(self.g, _ptr(vbits, _up), F, N, K, alpha, beta, gamma, iter,
                                                    int(checklb), _ptr(nt, _dp), _ptr(fo, _dp), _ptr(fz, _dp),
                                                    _ptr(E_W, _dp), _ptr(E_H, _dp), _ptr(lbs, _dp)))
        return E_W, E_H, lbs[:iter]

    def gibbs_dirber(self, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, burnin=0.5, seed=1):
        E_W = np.zeros((self.F, K), order="F"); E_H = np.zeros((self.N, K), order="F")
        self._ck(self.lib.nbmf_group_gibbs_dirber(self.g, K, alpha, beta, gamma, iter, burnin, seed, _ptr(E_W, _dp),
                                                  _ptr(E_H, _dp)))
        return E_W, E_H


# ---------------------------------------------------------------------------
# the reference's function names
# ---------------------------------------------------------------------------
_E_Z_LIMIT = 1 << 27  # doubles (1 GiB): above this the cube is not returned


def _engine_for(V, engine, **kw):
    if engine is not None:
        return engine, False
    e = Engine(**kw)
    e.set_data(V)
    return e, True


def VB_Aspect_Rcpp(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, checklb=False, *, engine=None,
                   device=0, precision=_lib.PREC_AUTO):
    """collapsed_gibbs_z_VB.cpp:373-526.  Returns the same named list; `likelihoods`
    and `kldivergences` are NaN as in the reference (:389-391)."""
    e, own = _engine_for(V, engine, device=device, precision=precision)
    try:
        e.init_from_labels(Z, K)
        want_ez = e.F * e.N * K <= _E_Z_LIMIT
        E_W, E_H, lbs, E_Z = e.vb_aspect(K, alpha, beta, gamma, iter, checklb, return_E_Z=want_ez)
    finally:
        if own:
            e.close()
    nan = np.full(iter, np.nan)
    return {"E_Z": E_Z, "E_W": E_W, "E_H": E_H, "lowerbounds": lbs, "likelihoods": nan, "kldivergences": nan.copy()}


def VB_DirBer_Rcpp(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, *, engine=None, device=0,
                   mode=_lib.DIRBER_EXACT_WAVEFRONT):
    """collapsed_gibbs_z_VB.cpp:97-168: K+1 columns, iter-1 sweeps, stale-row outputs."""
    e, own = _engine_for(V, engine, device=device, precision=_lib.PREC_FP64)
    try:
        e.init_from_labels(Z, K + 1)
        want_ez = mode == _lib.DIRBER_EXACT_WAVEFRONT and e.F * e.N * (K + 1) <= _E_Z_LIMIT
        E_W, E_H, E_Z = e.vb_dirber(K, alpha, beta, gamma, iter, mode, return_E_Z=want_ez)
    finally:
        if own:
            e.close()
    return {"E_Z": E_Z, "E_W": E_W, "E_H": E_H}


def Gibbs_DirBer_finite_Rcpp(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, burnin=0.5, *, seed=1,
                             test_idx=None, engine=None, device=0, collapsed=False):
    """collapsed_gibbs_z.cpp:363-437.  By default the engine runs the uncollapsed blocked chain
    (same stationary posterior); instead of the F x N x J `Z_samples` cube it
    returns the Rao-Blackwell means the R wrapper derives from it
    (generic_Gibbs_DirBer.R:19-20,54-101) plus the last assignment.
    `collapsed=True` runs the reference's own sweep (wavefront order, keyed uniforms) and
    returns the reference's `Z_samples` (F x N x J int; its `E_Z` is a function of it,
    compute_expectation_Z :157-173) plus the last assignment."""
    e, own = _engine_for(V, engine, device=device, precision=_lib.PREC_FP64)
    if collapsed:
        try:
            e.init_from_labels(Z, K)
            Zs, nw, Zl = e.gibbs_dirber_collapsed(K, alpha, beta, gamma, iter, burnin, seed)
        finally:
            if own:
                e.close()
        return {"Z_samples": Zs, "Z_last": Zl, "written": nw}
    try:
        e.init_from_labels(Z, K)
        E_W, E_H, E_V, Zl = e.gibbs_dirber(K, alpha, beta, gamma, iter, burnin, seed, test_idx, want_Z=True)
    finally:
        if own:
            e.close()
    return {"E_W": E_W, "E_H": E_H, "E_V_test": E_V, "Z_last": Zl}


def Gibbs_DirBer_DP_Rcpp(V, Z, alpha=1.0, beta=1.0, gamma=1.0, iter=100, burnin=0.5, *, seed=1, engine=None, device=0):
    """collapsed_gibbs_z.cpp:275-359 (Kmax = 200, `alpha` / `beta` unused there).  Returns `Z_samples`."""
    e, own = _engine_for(V, engine, device=device, precision=_lib.PREC_FP64)
    try:
        e.init_from_labels(Z, 200)
        Zs, nw, Zl = e.gibbs_dirber_dp_collapsed(gamma, iter, burnin, seed)
    finally:
        if own:
            e.close()
    return {"Z_samples": Zs, "Z_last": Zl, "written": nw}


def Gibbs_DirDir_finite_Rcpp(V, Z, K, alpha=1.0, gamma=1.0, iter=100, burnin=0.5, *, seed=1, engine=None, device=0):
    """collapsed_gibbs_z.cpp:441-557.  Returns `Z_samples` and `C_samples`."""
    e, own = _engine_for(V, engine, device=device, precision=_lib.PREC_FP64)
    try:
        e.init_from_labels(Z, K)
        Zs, Cs, nw, Zl, Cl = e.gibbs_dirdir_collapsed(K, alpha, gamma, iter, burnin, seed)
    finally:
        if own:
            e.close()
    return {"Z_samples": Zs, "C_samples": Cs, "Z_last": Zl, "C_last": Cl, "written": nw}


def lower_bound_aspect(V, n_total, f_ones, f_zeros, alpha, beta, gamma, *, device=0):
    """ELBO of the parameters implied by the statistics with q(Z) at its optimum
    (collapsed_gibbs_z_VB.cpp:339-369 evaluated right after the q(Z) update)."""
    e = Engine(device=device, precision=_lib.PREC_FP64)
    try:
        e.set_data(V)
        e.set_stats(n_total, f_ones, f_zeros)
        return e.lower_bound(np.asarray(n_total).shape[1], alpha, beta, gamma)[0]
    finally:
        e.close()


def _zero_base(Z):
    """`Z <- Z - min(Z, na.rm=TRUE)` (generic_VB_aspect.R:26, generic_VB_DirBer.R:8)."""
    Z = as_imat(Z).copy(order="F")
    m = Z != NA_INTEGER
    if m.any():
        Z[m] -= Z[m].min()
    return Z


def _check_nonneg(res):
    if not np.all(res["E_W"][~np.isnan(res["E_W"])] >= 0):
        raise ValueError("E_W contains negative values")
    if not np.all(res["E_H"][~np.isnan(res["E_H"])] >= 0):
        raise ValueError("E_W contains negative values")  # sic: generic_VB_aspect.R:31


def VB_aspectRcpp(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, **kw):
    """R/generic_VB_aspect.R:21-35."""
    res = VB_Aspect_Rcpp(V, _zero_base(Z), K, alpha, beta, gamma, iter, **kw)
    _check_nonneg(res)
    res["class"] = ("VBaspect", "VBBernoulli", "Bernoulli")
    return res


def VB_DirBer(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, **kw):
    """R/generic_VB_DirBer.R:5-17."""
    res = VB_DirBer_Rcpp(V, _zero_base(Z), K, alpha, beta, gamma, iter, **kw)
    _check_nonneg(res)
    res["class"] = ("VBDirBer", "VBBernoulli", "Bernoulli")
    return res


def Gibbs_DirBer(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, burnin=0.5, **kw):
    """R/generic_Gibbs_DirBer.R:1-38."""
    res = Gibbs_DirBer_finite_Rcpp(V, _zero_base(Z), K, alpha, beta, gamma, iter, burnin, **kw)
    _check_nonneg(res)
    res.update(V=as_imat(V), alpha=alpha, beta=beta, gamma=gamma)
    res["class"] = ("GibbsDirBer", "Bernoulli")
    return res


# -- the posterior-mean helpers the R wrappers apply to the sample cubes, on the device ----------------
# (compute_expectation_W_Rcpp :178-197, compute_expectation_H_Rcpp :230-272,
#  compute_expectation_H_dirdir_Rcpp :205-225; O(F N J K) in the reference)
def _with_engine(device, fn):
    e = Engine(device=device, precision=_lib.PREC_FP64)
    try:
        return fn(e)
    finally:
        e.close()


def compute_expectation_W_Rcpp(Z_samples, gamma, *, device=0):
    """normalise(gamma + mean_j n_total_j) over max(label)+1 columns; NA labels count for nothing."""
    return _with_engine(device, lambda e: e.compute_expectation_W(Z_samples, gamma))


def compute_expectation_H_Rcpp(Z_samples, V, alpha, beta, *, device=0):
    """mean_j (alpha + ones_j) / (alpha + beta + total_j); the reference's k loop stops at max(label), so
    the last column stays zero (:236-242) -- kept."""
    return _with_engine(device, lambda e: e.compute_expectation_H(Z_samples, V, alpha, beta))


def compute_expectation_H_dirdir_Rcpp(C_samples, alpha, *, device=0):
    """normalise(alpha + mean_j column counts of C) (:205-225, generic_Gibbs_DirDir.R:40-43): the W reducer
    applied to the cube with rows and columns exchanged."""
    Ct = np.asfortranarray(np.transpose(np.asarray(C_samples), (1, 0, 2)))
    return _with_engine(device, lambda e: e.compute_expectation_W(Ct, alpha))


def Gibbs_DirBer_DP(V, Z, alpha=1.0, beta=1.0, gamma=1.0, iter=100, burnin=0.5, **kw):
    """R/generic_Gibbs_DirBer_DP.R:1-28."""
    res = Gibbs_DirBer_DP_Rcpp(V, _zero_base(Z), alpha, beta, gamma, iter, burnin, **kw)
    dev = kw.get("device", 0)
    res["E_W"] = compute_expectation_W_Rcpp(res["Z_samples"], gamma, device=dev)
    res["E_H"] = compute_expectation_H_Rcpp(res["Z_samples"], V, alpha, beta, device=dev)
    _check_nonneg(res)
    res.update(alpha=alpha, beta=beta, gamma=gamma)
    res["class"] = ("GibbsDirBer", "Bernoulli")
    return res


def Gibbs_DirDir(V, Z, K, alpha=1.0, gamma=1.0, iter=100, burnin=0.5, **kw):
    """R/generic_Gibbs_DirDir.R:1-34."""
    res = Gibbs_DirDir_finite_Rcpp(V, _zero_base(Z), K, alpha, gamma, iter, burnin, **kw)
    dev = kw.get("device", 0)
    res["E_W"] = compute_expectation_W_Rcpp(res["Z_samples"], gamma, device=dev)
    res["E_H"] = compute_expectation_H_dirdir_Rcpp(res["C_samples"], alpha, device=dev)
    _check_nonneg(res)
    Vi = as_imat(V)
    if res["E_W"].shape[0] != Vi.shape[0] or res["E_H"].shape[0] != Vi.shape[1]:
        raise ValueError("Dimensions E_H and V do not match")
    res.update(V=Vi, alpha=alpha, gamma=gamma)
    res["class"] = ("GibbsDirDir", "Bernoulli")
    return res


def expectation(model):
    """R/commons.R:31-33 (VB) -- E_W %*% t(E_H).  For Gibbs results the engine
    already accumulated mean_j E_W^(j) E_H^(j)T on the held-out entries."""
    return model["E_W"] @ model["E_H"].T


def loglikelihood(model, V_test, *, device=0):
    """R/commons.R:17-24: sum over non-NA test entries of log(v p + (1-v)(1-p));
    evaluated on the device without forming the F x N matrix."""
    e = Engine(device=device, precision=_lib.PREC_FP64)
    try:
        ll, n = e.predictive_ll(model["E_W"], model["E_H"], V_test)
    finally:
        e.close()
    return {"loglikelihood": ll, "n": n}


def BernoulliNMF(V, K, K_init=None, model="DirBer", method="Gibbs", alpha=1.0, beta=1.0, gamma=1.0,
                 iter=1000, burnin=0.5, Z_init=None, *, seed=0, device=0):
    """R/BernoulliNMF.R:12-106 (VB: aspectRcpp, DirBer; Gibbs: DirBer, DirBerDP, DirDir).  Z initialisation: uniform labels on observed entries
    (:23-33); Gibbs warm-starts with 100 VB iterations at K_init=10 and takes the
    arg-max of E_Z (:41-52) -- computed on the device from the final operands."""
    Vi = as_imat(V)
    if K_init is None:
        K_init = K
    if method == "Gibbs":
        K_init = 10  # BernoulliNMF.R:18-20
    rng = np.random.default_rng(seed)
    if Z_init is None:
        Z = rng.integers(1, K_init + 1, size=Vi.shape).astype(np.int32)  # sample(K_init, 1), 1-based
        Z = np.asfortranarray(Z)
        Z[Vi == NA_INTEGER] = NA_INTEGER
    else:
        Z = as_imat(Z_init)
        if Z.shape != Vi.shape:
            raise ValueError("dim(Z) != dim(V)")
    if method == "Gibbs":
        if model not in ("DirBer", "DirBerDP", "DirDir"):
            raise ValueError("model must be 'DirBer', 'DirBerDP' or 'DirDir' for method='Gibbs'")
        if model != "DirBerDP" and K < 10:
            raise ValueError("K < K_init=10: the reference would write counters out of bounds "
                             "(BernoulliNMF.R:18-20, SURVEY.md Appendix B.9)")
        # warm start and which.max hand-off on the device: no E_Z cube crosses the boundary
        e = Engine(device=device, precision=_lib.PREC_FP64 if model != "DirBer" else _lib.PREC_AUTO)
        try:
            e.set_data(Vi)
            e.init_from_labels(_zero_base(Z), K_init)
            e.vb_aspect(K_init, alpha, beta, gamma, iter=100)
            Zw = e.argmax_labels(want_host=(model != "DirBer"))
            if model == "DirBer":
                E_W, E_H, E_V, Zl = e.gibbs_dirber(K, alpha, beta, gamma, iter, burnin, seed, None, want_Z=True)
        finally:
            e.close()
        if model == "DirBer":
            res = {"E_W": E_W, "E_H": E_H, "E_V_test": E_V, "Z_last": Zl, "V": Vi, "alpha": alpha, "beta": beta,
                   "gamma": gamma, "class": ("GibbsDirBer", "Bernoulli")}
            _check_nonneg(res)
        elif model == "DirBerDP":       # BernoulliNMF.R:60-64
            res = Gibbs_DirBer_DP(Vi, Zw, alpha, beta, gamma, iter, burnin, seed=seed, device=device)
        else:                           # BernoulliNMF.R:66-70
            res = Gibbs_DirDir(Vi, Zw, K, alpha, gamma, iter, burnin, seed=seed, device=device)
    elif method == "VB":
        if model == "aspectRcpp" or model == "aspect":
            res = VB_aspectRcpp(Vi, Z, K, alpha, beta, gamma, iter, device=device)
        elif model == "DirBer":
            res = VB_DirBer(Vi, Z, K, alpha, beta, gamma, iter, device=device)
        else:
            raise NotImplementedError("Collapsed version. Not implemented")
    else:
        raise ValueError("method must be 'Gibbs' or 'VB'")
    if res["E_W"].shape[0] != Vi.shape[0]:
        raise ValueError("Dimensions E_W and V do not match")
    if res["E_H"].shape[0] != Vi.shape[1]:
        raise ValueError("Dimensions E_H and V do not match")
    return res
