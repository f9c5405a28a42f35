"""ctypes binding of libnbmf_cuda.so (include/nbmf_cuda.h).

The shared library is built in-tree by `make -C nbmf_b200/csrc` (or
`__graft_entry__.build()`).  There is no fallback: if the library is missing, or
no CUDA device is present when a handle is created, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# NBMF_LIB_VARIANT=prof selects the instrumented build (libnbmf_cuda_prof.so, `make prof`): debug aid of tools/role_profile.py
SO_PATH = os.path.join(HERE, "libnbmf_cuda" + ("_" + os.environ["NBMF_LIB_VARIANT"] if os.environ.get("NBMF_LIB_VARIANT") else "") + ".so")

NBMF_OK = 0
PREC_AUTO, PREC_FP64, PREC_TENSOR, PREC_TENSOR_FAST, PREC_TENSOR_LO8, PREC_FP32 = 0, 1, 2, 3, 4, 5
PATH_FP64, PATH_TENSOR, PATH_TENSOR_FAST, PATH_TENSOR_LO8, PATH_FP32 = 0, 1, 2, 3, 4
DIRBER_EXACT_WAVEFRONT, DIRBER_CONTRACTION = 0, 1


class NbmfOpts(C.Structure):
    _fields_ = [("device", C.c_int), ("precision", C.c_int), ("flush_interval", C.c_int),
                ("pipe_block_cols", C.c_int), ("n_global", C.c_int64), ("n_offset", C.c_int64),
                ("planned_iters", C.c_int), ("reserved", C.c_int)]


class NbmfTiming(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("params_ms", C.c_double), ("pass_rows_ms", C.c_double),
                ("pass_cols_ms", C.c_double), ("exchange_ms", C.c_double), ("launches", C.c_int64),
                ("iters", C.c_int64), ("path", C.c_int64)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint32)
_lp = C.POINTER(C.c_int64)
_H = C.c_void_p

# name -> (restype, argtypes): every symbol include/nbmf_cuda.h declares
SYMBOLS = {
    "nbmf_version": (C.c_char_p, []),
    "nbmf_device_count": (C.c_int, []),
    "nbmf_create": (C.c_int, [C.POINTER(_H), C.POINTER(NbmfOpts)]),
    "nbmf_destroy": (None, [_H]),
    "nbmf_last_error": (C.c_char_p, [_H]),
    "nbmf_stream": (C.c_void_p, [_H]),
    "nbmf_synchronize": (C.c_int, [_H]),
    "nbmf_set_allreduce_hook": (C.c_int, [_H, ALLREDUCE_FN, C.c_void_p]),
    "nbmf_get_timing": (C.c_int, [_H, C.POINTER(NbmfTiming)]),
    "nbmf_set_data_i32": (C.c_int, [_H, _ip, C.c_int64, C.c_int64]),
    "nbmf_set_data_bits": (C.c_int, [_H, _up, _up, C.c_int64, C.c_int64]),
    "nbmf_generate_synthetic": (C.c_int, [_H, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_uint64]),
    "nbmf_get_dims": (C.c_int, [_H, _lp, _lp, _lp]),
    "nbmf_get_data_bits": (C.c_int, [_H, _up, _up]),
    "nbmf_init_from_labels": (C.c_int, [_H, _ip, C.c_int]),
    "nbmf_init_random_labels": (C.c_int, [_H, C.c_int, C.c_uint64]),
    "nbmf_set_stats": (C.c_int, [_H, C.c_int, _dp, _dp, _dp]),
    "nbmf_get_stats": (C.c_int, [_H, C.c_int, _dp, _dp, _dp]),
    "nbmf_vb_aspect": (C.c_int, [_H, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, _dp]),
    "nbmf_vb_aspect_bits": (C.c_int, [_H, _up, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double,
                                      C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]),
    "nbmf_vb_begin": (C.c_int, [_H, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]),
    "nbmf_vb_step": (C.c_int, [_H, C.c_int]),
    "nbmf_vb_end": (C.c_int, [_H, _dp, _dp, _dp, C.c_int, _dp]),
    "nbmf_vb_get_lb_terms": (C.c_int, [_H, _dp, C.c_int]),
    "nbmf_vb_argmax_labels": (C.c_int, [_H, _ip]),
    "nbmf_lower_bound": (C.c_int, [_H, C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp]),
    "nbmf_vb_dirber": (C.c_int, [_H, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp]),
    "nbmf_gibbs_dirber": (C.c_int, [_H, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double,
                                    C.c_uint64, _dp, _dp, _lp, C.c_int64, _dp, _ip]),
    "nbmf_gibbs_dirber_collapsed": (C.c_int, [_H, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double,
                                              C.c_uint64, _ip, C.POINTER(C.c_int), _ip]),
    "nbmf_gibbs_dirber_dp_collapsed": (C.c_int, [_H, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double,
                                                 C.c_uint64, _ip, C.POINTER(C.c_int), _ip]),
    "nbmf_gibbs_dirdir_collapsed": (C.c_int, [_H, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_uint64,
                                              _ip, _ip, C.POINTER(C.c_int), _ip, _ip]),
    "nbmf_gibbs_sweep_given_W": (C.c_int, [_H, _ip, _dp, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_double,
                                           C.c_uint64, _dp]),
    "nbmf_predictive_ll": (C.c_int, [_H, _dp, _dp, C.c_int, _ip, C.c_int64, C.c_int64, _dp, _lp]),
    "nbmf_compute_expectation_W": (C.c_int, [_H, _ip, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_int, _dp]),
    "nbmf_compute_expectation_H": (C.c_int, [_H, _ip, _ip, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_int, _dp]),
    "nbmf_samples_VWH_DirBer": (C.c_int, [_H, _ip, _ip, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double,
                                          C.c_double, C.c_uint64, _dp, _dp, _ip]),
    "nbmf_lower_bound_aspect_full": (C.c_int, [_H, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double,
                                               _ip, C.c_int64, C.c_int64, C.c_int, _dp, _dp]),
    "nbmf_sample_gibbs_cpp": (C.c_int, [_H, _ip, _dp, _ip, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_double, C.c_uint64,
                                        C.c_int, _dp]),
    "nbmf_comm_unique_id": (C.c_int, [C.c_void_p]),
    "nbmf_comm_init_rank": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int]),
    "nbmf_comm_destroy": (C.c_int, [_H]),
    "nbmf_comm_info": (C.c_int, [_H, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nbmf_group_create": (C.c_int, [C.POINTER(_H), C.POINTER(NbmfOpts), C.POINTER(C.c_int), C.c_int]),
    "nbmf_group_destroy": (None, [_H]),
    "nbmf_group_size": (C.c_int, [_H]),
    "nbmf_group_handle": (C.c_void_p, [_H, C.c_int]),
    "nbmf_group_last_error": (C.c_char_p, [_H]),
    "nbmf_group_columns": (C.c_int, [_H, C.c_int, _lp, _lp]),
    "nbmf_group_set_data_i32": (C.c_int, [_H, _ip, C.c_int64, C.c_int64]),
    "nbmf_group_generate_synthetic": (C.c_int, [_H, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_uint64]),
    "nbmf_group_init_from_labels": (C.c_int, [_H, _ip, C.c_int]),
    "nbmf_group_init_random_labels": (C.c_int, [_H, C.c_int, C.c_uint64]),
    "nbmf_group_vb_begin": (C.c_int, [_H, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]),
    "nbmf_group_vb_step": (C.c_int, [_H, C.c_int]),
    "nbmf_group_vb_end": (C.c_int, [_H, _dp, _dp, _dp, C.c_int]),
    "nbmf_group_synchronize": (C.c_int, [_H]),
    "nbmf_group_vb_aspect": (C.c_int, [_H, _ip, _ip, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double,
                                       C.c_int, C.c_int, _dp, _dp, _dp]),
    "nbmf_group_vb_aspect_bits": (C.c_int, [_H, _up, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double,
                                            C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]),
    "nbmf_group_gibbs_dirber": (C.c_int, [_H, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double,
                                          C.c_uint64, _dp, _dp]),
}

# debug / measurement aids exported by the library but not part of include/nbmf_cuda.h
DEBUG_SYMBOLS = {
    "nbmf_debug_model_error": (C.c_double, [C.c_int, C.c_double, C.c_int]),
    "nbmf_debug_rng_probe": (C.c_int, [_H, C.c_int, C.c_double, C.c_double, C.c_uint64, C.c_int64, _dp]),
    "nbmf_debug_ffma_peak": (C.c_int, [_H, _dp]),
}

_lib = None


class NbmfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libnbmf_cuda error {code}: {msg}")
        self.code = code


def _preload_nccl():
    """libnbmf_cuda needs libnccl.so.2.  A process that also runs torch.distributed must end up with ONE
    copy of NCCL: load the wheel's libnccl (the one torch uses) first when it is there, so that the
    library's dependency resolves to it by SONAME; otherwise the system libnccl serves."""
    import sys
    for sp in sys.path:
        cand = os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
            return


def load():
    """Load libnbmf_cuda.so and bind every declared symbol.  Raises if the
    library has not been built -- there is no Python / CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} not found: build it with `make -C nbmf_b200/csrc` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
    _preload_nccl()
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in DEBUG_SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
