"""ctypes binding of libshgo_b200.so -- the C ABI declared in include/shgo_b200.h.

There is deliberately no fallback: if the library has not been built, or a call fails (no CUDA
device, bad arguments), a Python exception carrying shgo_b200_last_error() is raised.
"""
import ctypes
import os

from . import build as _build

_c = ctypes
_vp, _i, _u64, _sz = _c.c_void_p, _c.c_int, _c.c_uint64, _c.c_size_t

# name -> (restype, argtypes); mirrors include/shgo_b200.h one to one
SIGNATURES = {
    "shgo_b200_version": (_i, []),
    "shgo_b200_last_error": (_c.c_char_p, []),
    "shgo_b200_device_info": (_i, [_i, _vp, _vp, _vp, _vp]),
    "shgo_b200_sobol_fill": (_i, [_vp, _i, _u64, _u64, _vp, _vp, _vp, _u64, _i, _vp]),
    "shgo_b200_eval_sobol": (_i, [_i, _i, _i, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "shgo_b200_eval_points": (_i, [_i, _i, _i, _u64, _vp, _u64, _vp, _vp, _vp]),
    "shgo_b200_mask_infeasible": (_i, [_vp, _u64, _vp, _i, _vp, _vp]),
    "shgo_b200_csr_workspace": (_sz, [_u64, _i, _u64]),
    "shgo_b200_csr_from_simplices": (_i, [_vp, _u64, _i, _u64, _vp, _vp, _u64, _vp, _vp, _sz, _vp]),
    "shgo_b200_csr_edges3": (_i, [_vp, _vp, _u64, _vp, _u64, _vp, _vp]),
    "shgo_b200_row_present": (_i, [_vp, _u64, _vp, _vp]),
    "shgo_b200_first_seen": (_i, [_vp, _u64, _i, _u64, _vp, _vp]),
    "shgo_b200_morton_workspace": (_sz, [_u64]),
    "shgo_b200_morton_order": (_i, [_vp, _u64, _i, _u64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "shgo_b200_relabel3": (_i, [_vp, _u64, _i, _u64, _vp, _vp, _vp]),
    "shgo_b200_gather_f64": (_i, [_vp, _vp, _u64, _vp, _vp]),
    "shgo_b200_gather_u8": (_i, [_vp, _vp, _u64, _vp, _vp]),
    "shgo_b200_scatter_u8": (_i, [_vp, _vp, _u64, _vp, _vp]),
    "shgo_b200_star_csr": (_i, [_vp, _vp, _vp, _u64, _u64, _u64, _vp, _vp]),
    "shgo_b200_star_csr_nfev": (_i, [_vp, _vp, _vp, _vp, _u64, _u64, _u64, _vp, _vp, _vp]),
    "shgo_b200_star_csr_local": (_i, [_vp, _vp, _vp, _vp, _u64, _u64, _u64, _u64, _u64, _i, _vp, _vp, _sz, _vp,
                                      _vp]),
    "shgo_b200_star_local_workspace": (_sz, [_u64]),
    "shgo_b200_star_remote_workspace": (_sz, [_u64, _u64]),
    "shgo_b200_star_csr_remote": (_i, [_vp, _vp, _vp, _vp, _i, _u64, _i, _u64, _u64, _i, _vp, _vp, _u64, _vp, _sz, _vp]),
    "shgo_b200_star_csr_remote_filtered": (_i, [_vp, _vp, _vp, _vp, _i, _u64, _i, _u64, _u64, _i, _vp, _vp, _u64, _vp,
                                                _sz, _vp, _i, _vp]),
    "shgo_b200_blockmin_layout": (_i, [_i, _i]),
    "shgo_b200_eval_sobol_blockmin": (_i, [_i, _i, _i, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "shgo_b200_star_csr_fused": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _u64, _i, _u64, _u64, _u64, _vp, ctypes.c_uint32,
                                      _vp, _vp, _sz, _vp, _vp]),
    "shgo_b200_publish_ready": (_i, [_vp, _i, _i, ctypes.c_uint32, _vp]),
    "shgo_b200_ipc_alloc": (_i, [_sz, _vp, _vp]),
    "shgo_b200_ipc_open": (_i, [_vp, _vp]),
    "shgo_b200_copy_async": (_i, [_vp, _vp, _sz, _vp]),
    "shgo_b200_ipc_close": (_i, [_vp]),
    "shgo_b200_ipc_free": (_i, [_vp]),
    "shgo_b200_compact_workspace": (_sz, [_u64]),
    "shgo_b200_compact_pool": (_i, [_vp, _u64, _c.c_int64, _vp, _u64, _vp, _vp, _sz, _vp]),
    "shgo_b200_argmin": (_i, [_vp, _vp, _u64, _vp, _vp, _vp, _sz, _vp]),
    "shgo_b200_subspace_mask": (_i, [_vp, _i, _u64, _vp, _vp]),
    "shgo_b200_gather_rows": (_i, [_vp, _u64, _i, _vp, _u64, _vp, _vp]),
    "shgo_b200_pool_rank_workspace": (_sz, [_u64]),
    "shgo_b200_pool_rank": (_i, [_vp, _u64, _vp, _vp, _u64, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "shgo_b200_argmin_ranked": (_i, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _sz, _vp]),
    "shgo_b200_count_nfev": (_i, [_vp, _vp, _u64, _vp, _vp]),
    "shgo_b200_lattice_eval": (_i, [_i, _i, _i, _i, _vp, _u64, _u64, _vp, _vp, _vp]),
    "shgo_b200_lattice_eval_push": (_i, [_i, _i, _i, _i, _vp, _u64, _u64, _vp, _vp, _vp, _i, _i, _i, ctypes.c_uint32, _vp]),
    "shgo_b200_lattice_coords": (_i, [_i, _i, _vp, _u64, _u64, _vp, _u64, _vp]),
    "shgo_b200_lattice_star_workspace": (_sz, [_u64]),
    "shgo_b200_lattice_star": (_i, [_i, _i, _vp, _u64, _u64, _vp, _vp, _sz, _vp]),
    "shgo_b200_lattice_star_part": (_i, [_i, _i, _vp, _u64, _u64, _i, _i, ctypes.c_uint32, _vp, _vp, _sz, _vp]),
    "shgo_b200_local_bounds_csr": (_i, [_vp, _vp, _vp, _u64, _i, _vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    "shgo_b200_lattice_local_bounds": (_i, [_i, _i, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    "shgo_b200_ctx_create": (_i, [_i, _vp]),
    "shgo_b200_ctx_destroy": (None, [_vp]),
    "shgo_b200_iterate_sobol_host": (_i, [_vp, _i, _i, _i, _u64, _u64, _vp, _vp, _vp, _vp, _vp,
                                          _vp, _u64, _vp, _vp, _vp]),
    "shgo_b200_ctx_shard_init": (_i, [_vp, _i, _i, _u64, _vp]),
    "shgo_b200_ctx_shard_connect": (_i, [_vp, _vp]),
    "shgo_b200_iterate_sobol_host_shard_begin": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "shgo_b200_iterate_sobol_host_shard_finish": (_i, [_vp, _vp, _u64, _vp, _vp]),
    "shgo_b200_fp64_peak": (_i, [_i, _vp, _vp, _vp]),
}

LIB_PATH = _build.LIB
_lib = None


class Shgob200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def lib():
    """Load (once) and return the ctypes handle; raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libshgo_b200.so is missing (%s). Build it with `python -m shgo_b200.build`; "
                "shgo_b200 has no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().shgo_b200_last_error()
        raise Shgob200Error("shgo_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def call(name, *args):
    check(getattr(lib(), name)(*args))
