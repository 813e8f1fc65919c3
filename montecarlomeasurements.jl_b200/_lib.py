"""ctypes binding of libmcm_b200.so (include/mcm_b200.h).

This is the Python stand-in for the Julia `ccall` layer:
same entry points, same argument meaning.  There is NO CPU fallback: if the
shared library or a CUDA device is missing, importing works (so the host logic
can be unit-tested) but the first device call raises `MCMError`.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCM_B200_LIB") or os.path.join(HERE, "libmcm_b200.so")   # override: A/B builds in scripts/
HEADER_PATH = os.path.join(HERE, "..", "include", "mcm_b200.h")


class MCMError(RuntimeError):
    """Any non-zero status from the C ABI."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[mcm_b200 status {code}] {msg}")
        self.code = code


class ArgumentError(MCMError, ValueError):
    """MCM_ERR_INVALID — the reference throws ArgumentError (types.jl:38,88; bymap.jl:34-35)."""


class MethodError(MCMError, TypeError):
    """MCM_ERR_UNSUPPORTED — no method for these argument types (e.g. N mismatch, `//` on floats)."""


MCM_OK, MCM_ERR_INVALID, MCM_ERR_CUDA, MCM_ERR_NOMEM, MCM_ERR_UNSUPPORTED, MCM_ERR_NCCL, MCM_ERR_NOT_INIT = range(7)
F32, F64 = 0, 1


def _parse_enums() -> dict:
    """Read every enumerator of include/mcm_b200.h so Python never drifts from the ABI."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for body in re.findall(r"enum\s*\w*\s*\{(.*?)\}", txt, flags=re.S):
        val = -1
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, expr = [s.strip() for s in item.split("=", 1)]
                expr = expr.rstrip("u")
                val = int(expr, 0)
            else:
                name = item
                val += 1
            out[name] = val
    for name, v in re.findall(r"#define\s+(MCM_\w+)\s+(0x[0-9a-fA-F]+|\d+)\s", txt):
        out[name] = int(v, 0)
    return out


ENUMS = _parse_enums()


class TapeIns(C.Structure):
    _fields_ = [("op", C.c_uint8), ("dst", C.c_uint8), ("a", C.c_uint8), ("b", C.c_uint8),
                ("c", C.c_uint8), ("pad", C.c_uint8 * 3)]


_vp, _sz, _i, _d = C.c_void_p, C.c_size_t, C.c_int, C.c_double
_dp = C.POINTER(C.c_double)
_vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); the complete export list of include/mcm_b200.h
SIGNATURES = {
    "mcm_version": (_i, []),
    "mcm_init": (_i, [_i]),
    "mcm_finalize": (_i, []),
    "mcm_set_stream": (_i, [_vp]),
    "mcm_get_stream": (_vp, []),
    "mcm_sync": (_i, []),
    "mcm_last_error": (C.c_char_p, []),
    "mcm_check_flags": (_i, [C.POINTER(C.c_uint32), _i]),
    "mcm_clear_flags": (_i, [C.c_uint32]),
    "mcm_device_info": (_i, [C.POINTER(_i), C.POINTER(_sz), C.POINTER(_sz)]),
    "mcm_alloc": (_i, [_vpp, _sz]),
    "mcm_alloc_many": (_i, [_vpp, _i, _sz]),
    "mcm_free": (_i, [_vp]),
    "mcm_host_alloc": (_i, [_vpp, _sz]),
    "mcm_host_alloc_wc": (_i, [_vpp, _sz]),
    "mcm_host_free": (_i, [_vp]),
    "mcm_upload": (_i, [_vp, _vp, _sz]),
    "mcm_download": (_i, [_vp, _vp, _sz]),
    "mcm_copy": (_i, [_vp, _vp, _sz]),
    "mcm_fill": (_i, [_i, _vp, _sz, _d]),
    "mcm_convert": (_i, [_i, _vp, _i, _vp, _sz]),
    "mcm_unary": (_i, [_i, _i, _vp, _vp, _sz]),
    "mcm_binary": (_i, [_i, _i, _vp, _d, _vp, _d, _vp, _sz]),
    "mcm_ternary": (_i, [_i, _i, _vp, _d, _vp, _d, _vp, _d, _vp, _sz]),
    "mcm_powi": (_i, [_i, _vp, _i, _vp, _sz]),
    "mcm_polyval": (_i, [_i, _vp, _dp, _i, _i, _vp, _sz]),
    "mcm_affine": (_i, [_i, _vp, _d, _d, _vp, _sz]),
    "mcm_lincomb": (_i, [_i, _i, _vpp, _dp, _vp, _sz]),
    "mcm_chain": (_i, [_i, C.POINTER(TapeIns), _i, _i, _i, _i, _dp, _i, _vpp, _i, _vpp, _i, _sz, _i, _dp]),
    "mcm_chain_stats_fetch": (_i, [_dp]),
    "mcm_chain_ticket": (C.c_uint64, []),
    "mcm_chain_stats_wait": (_i, [C.c_uint64, _dp]),
    "mcm_chain_sharded": (_i, [_i, C.POINTER(TapeIns), _i, _i, _i, _i, _dp, _i, _vpp, _i, _vpp, _i, _sz, _i, _dp]),
    "mcm_chain_host": (_i, [_i, C.POINTER(TapeIns), _i, _i, _i, _i, _dp, _i, _vpp, _i, _vpp, _i, _sz, _i, _dp]),
    "mcm_chain_last_engine": (_i, []),
    "mcm_chain_force_interpreter": (_i, [_i]),
    "mcm_chain_jit_mode": (_i, [_i]),
    "mcm_chain_force_engine": (_i, [_i]),
    "mcm_chain_jit_compiles": (_i, []),
    "mcm_chain_jit_option": (_i, [_i, _i]),
    "mcm_chain_jit_source": (_i, [_i, C.POINTER(TapeIns), _i, _i, _i, _dp, _i, _i, _i, C.c_char_p, _sz, C.POINTER(_sz), C.c_char_p]),
    "mcm_moments": (_i, [_i, _vp, _sz, _dp]),
    "mcm_moments_enqueue": (_i, [_i, _vp, _sz]),
    "mcm_moments_collect": (_i, [_dp]),
    "mcm_moments_use_bulk": (_i, [_i]),
    "mcm_reduce": (_i, [_i, _i, _vp, _sz, _dp]),
    "mcm_count": (_i, [_i, _i, _vp, _sz, _d, C.POINTER(C.c_uint64)]),
    "mcm_count2": (_i, [_i, _i, _vp, _vp, _sz, C.POINTER(C.c_uint64)]),
    "mcm_round_int": (_i, [_i, _vp, _i, _vp, _sz]),
    "mcm_moments_merge": (_i, [_dp, _dp, _dp]),
    "mcm_gram": (_i, [_i, _vpp, _i, _sz, _i, _dp, _vp, _vp]),
    "mcm_gram_use_tensor": (_i, [_i]),
    "mcm_gram_last_engine": (_i, []),
    "mcm_gram_last_digits": (_i, []),
    "mcm_colmeans": (_i, [_i, _vpp, _i, _sz, _dp]),
    "mcm_sort": (_i, [_i, _vp, _vp, _sz]),
    "mcm_sortperm": (_i, [_i, _vp, _vp, _vp, _sz]),
    "mcm_quantile": (_i, [_i, _vp, _sz, _dp, _i, _dp]),
    "mcm_quantile_use_select": (_i, [_i]),
    "mcm_sort_use_fused": (_i, [_i]),
    "mcm_norminvcdf": (_i, [_i, _vp, _vp, _sz]),
    "mcm_systematic_normal": (_i, [_i, _vp, _sz, _d, _d, _vp, C.c_uint64, C.c_uint64]),
    "mcm_systematic_uniform": (_i, [_i, _vp, _sz, _d, _d, _vp, C.c_uint64, C.c_uint64]),
    "mcm_systematic_dist": (_i, [_i, _i, _d, _d, _d, _vp, _sz, _vp, C.c_uint64, C.c_uint64]),
    "mcm_rand_dist": (_i, [_i, _i, _d, _d, _d, _vp, _sz, C.c_uint64, C.c_uint64]),
    "mcm_quantile_dist": (_i, [_i, _i, _d, _d, _d, _vp, _vp, _sz]),
    "mcm_randperm": (_i, [_vp, _sz, C.c_uint64]),
    "mcm_resample_use_search": (_i, [_i]),
    "mcm_rand_real": (_i, [_i, _i, _vp, _sz, _d, _d, C.c_uint64, C.c_uint64]),
    "mcm_autocor": (_i, [_i, _vp, _sz, _i, _dp]),
    "mcm_window_create": (_i, [_sz, C.POINTER(_i), _vpp]),
    "mcm_window_free": (_i, [_i]),
    "mcm_comm_fence": (_i, []),
    "mcm_gather_peer": (_i, [_i, _i, _i, C.c_uint64, C.c_uint64, _vp, _sz, _vpp]),
    "mcm_window_fill": (_i, [_i, _i, _vpp, _i, _sz, C.c_uint64]),
    "mcm_window_bytes": (_sz, [_i, C.c_uint64, _i]),
    "mcm_resample_systematic_sharded": (_i, [_vp, _sz, C.c_uint64, C.c_uint64, C.c_uint64, _d, _i, C.c_uint64]),
    "mcm_complex_op": (_i, [_i, _i, _vp, _vp, _d, _d, _vp, _vp, _d, _d, _vp, _vp, _sz]),
    "mcm_dense_use_v2": (_i, [_i]),
    "mcm_batched_linalg": (_i, [_i, _i, _i, _i, _vpp, _vpp, _vpp, _vpp, _sz]),
    "mcm_batched_linalg_complex": (_i, [_i, _i, _i, _vpp, _vpp, _vpp, _vpp, _sz]),
    "mcm_argmin_absdiff": (_i, [_i, _vp, _sz, _d, C.POINTER(C.c_uint64)]),
    "mcm_argmin_sqdist": (_i, [_i, _vpp, _i, _sz, _dp, C.POINTER(C.c_uint64)]),
    "mcm_shuffle_insert": (_i, [_i, _vp, _sz, C.c_uint64, _d, _vp]),
    "mcm_filter_range": (_i, [_i, _vp, _sz, _d, _d, _vp, _sz, C.POINTER(C.c_size_t)]),
    "mcm_index_bijection": (_i, [_vp, _sz, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    "mcm_systematic_batch": (_i, [_i, _i, _vpp, _i, _sz, _dp, _dp, _i, C.c_uint64, C.c_uint64, C.c_uint64]),
    "mcm_systematic_use_scatter": (_i, [_i]),
    "mcm_quantile_batch": (_i, [_i, _vpp, _i, _sz, _dp, _i, _dp]),
    "mcm_quantile_batch_fallbacks": (_i, []),
    "mcm_quantile_sharded": (_i, [_i, _vpp, _i, _sz, _dp, _i, _dp]),
    "mcm_order_statistics": (_i, [_i, _vp, _sz, C.POINTER(C.c_uint64), _i, _dp]),
    "mcm_pgemm": (_i, [_i, _dp, _i, _i, _vpp, _dp, _dp, _vpp, _sz]),
    "mcm_potrf": (_i, [_vp, _i, C.POINTER(C.c_int)]),
    "mcm_trsm_right_lower": (_i, [_vp, _vp, _i, _i]),
    "mcm_transform_moments": (_i, [_i, _vpp, _i, _sz, _dp, _dp, _i, _vpp]),
    "mcm_pgemm_use_tensor": (_i, [_i]),
    "mcm_pgemm_last_engine": (_i, []),
    "mcm_rand_indices": (_i, [_vp, _sz, C.c_uint64, C.c_uint64, C.c_uint64]),
    "mcm_gather": (_i, [_i, _vp, _vp, _vp, _sz, _sz]),
    "mcm_gather_multi": (_i, [_i, _vpp, _vpp, _i, _vp, _sz, _sz]),
    "mcm_resample_systematic": (_i, [_vp, _sz, _d, _vp]),
    "mcm_comm_unique_id": (_i, [_vp]),
    "mcm_comm_init": (_i, [_vp, _i, _i]),
    "mcm_comm_destroy": (_i, []),
    "mcm_comm_info": (_i, [C.POINTER(_i), C.POINTER(_i)]),
    "mcm_comm_allreduce_sum_f64": (_i, [_vp, _sz]),
    "mcm_moments_sharded": (_i, [_i, _vp, _sz, _dp]),
    "mcm_moments_allmerge": (_i, [_dp]),
    "mcm_gram_sharded": (_i, [_i, _vpp, _i, _sz, _i, _vp]),
}

_lib = None
_inited_device = None


def load():
    """dlopen the library and bind every symbol the header declares (fails loudly)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MCMError(MCM_ERR_NOT_INIT,
                       f"{LIB_PATH} is missing: run `python montecarlomeasurements.jl_b200/build.py` "
                       "(or __graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc == MCM_OK:
        return
    msg = (load().mcm_last_error() or b"").decode(errors="replace")
    if rc == MCM_ERR_INVALID:
        raise ArgumentError(rc, msg)
    if rc == MCM_ERR_UNSUPPORTED:
        raise MethodError(rc, msg)
    raise MCMError(rc, msg)


def init(device: int | None = None):
    """Bind this process to one GPU (one process per GPU; LOCAL_RANK by default)."""
    global _inited_device
    lib = load()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _inited_device is not None:
        if _inited_device != device:
            raise ArgumentError(MCM_ERR_INVALID, f"already initialised on device {_inited_device}")
        return lib
    check(lib.mcm_init(device))
    _inited_device = device
    return lib


def lib():
    """The initialised library (initialises on device LOCAL_RANK/0 on first use)."""
    if _inited_device is None:
        init()
    return _lib


def is_initialised() -> bool:
    return _inited_device is not None
