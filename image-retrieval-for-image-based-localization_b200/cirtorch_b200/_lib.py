"""ctypes binding of libcir_b200.so (include/cir_b200.h).

The library is the product: there is NO CPU or eager-PyTorch fallback behind these calls.
If the shared object is missing, or an op is invoked on a non-CUDA tensor, the call
raises -- loudly -- instead of silently computing somewhere else.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CIR_LIB_PATH: an A/B build of the same ABI (experiments); the product is the in-tree library
LIB_PATH = os.environ.get("CIR_LIB_PATH") or os.path.join(_HERE, "libcir_b200.so")

_c_int = C.c_int
_c_i64 = C.c_int64
_c_f = C.c_float
_vp = C.c_void_p
_szp = C.POINTER(C.c_size_t)

# name -> (restype, argtypes); must list every symbol include/cir_b200.h declares
SIGNATURES = {
    "cir_last_error": (C.c_char_p, []),
    "cir_version": (_c_int, []),
    "cir_launch_count": (_c_i64, [_c_int]),
    "cir_tail_workspace_bytes": (_c_int, [_c_int, _c_int, _c_int, _szp]),
    "cir_tail_fwd": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _vp, _c_int, _c_f, _c_f, _c_int,
                              _vp, _vp, _c_int, _vp, _c_int, _vp, _vp, C.c_size_t, C.c_uint, _vp]),
    "cir_tail_fwd_train": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _vp, _c_int, _c_f, _c_f, _c_int,
                                    _vp, _vp, _c_int, _vp, _c_int, _vp, _vp, _vp, C.c_size_t, C.c_uint, _vp]),
    "cir_gem_bwd": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _vp, _c_int, _c_f, _vp, _vp, _vp, _vp, _vp]),
    "cir_bias_l2n_rows": (_c_int, [_vp, _c_i64, _c_int, _c_i64, _vp, _c_f, _vp, _c_i64, _vp]),
    "cir_powerlaw": (_c_int, [_vp, _c_i64, _c_f, _vp, _vp]),
    "cir_region_pool": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _vp, _c_int, _vp, _c_int, _c_f, _c_int, _vp, _vp]),
    "cir_l2n_rows": (_c_int, [_vp, _c_i64, _c_int, _c_i64, _c_f, _vp, _c_i64, _vp]),
    "cir_pack_bf16": (_c_int, [_vp, _c_i64, _c_int, _c_i64, _vp, _c_i64, _c_int, _c_int, _vp]),
    "cir_search_plan": (_c_int, [_c_int, _c_i64, _c_int, _c_int, _vp]),
    "cir_search_workspace_bytes": (_c_int, [_c_int, _c_i64, _c_int, _c_int, _szp]),
    "cir_search_topk": (_c_int, [_vp, _c_int, _vp, _c_i64, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp,
                                 C.c_int32, _vp, C.c_size_t, C.c_uint, _vp]),
    "cir_search_topk_exchange": (_c_int, [_vp, _c_int, _vp, _c_i64, _c_int, _c_int, C.c_int32, C.POINTER(C.c_void_p), _c_int,
                                          _c_int, _vp, C.c_size_t, C.c_uint, _vp]),
    "cir_search_topk_exchange_merge": (_c_int, [_vp, _c_int, _vp, _c_i64, _c_int, _c_int, C.c_int32, C.POINTER(C.c_void_p), _c_int,
                                                _c_int, C.c_uint32, _vp, _vp, _vp, C.c_size_t, C.c_uint, _vp]),
    "cir_scores_dense": (_c_int, [_vp, _c_int, _vp, _c_i64, _c_int, _vp, _c_i64, _vp]),
    "cir_sort_rows_workspace_bytes": (_c_int, [_c_int, _c_i64, _szp]),
    "cir_sort_rows_desc": (_c_int, [_vp, _c_int, _c_i64, _c_i64, _vp, _vp, _vp, C.c_size_t, _vp]),
    "cir_topk_merge": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_i64, _vp, _vp, _c_int, _vp]),
    "cir_rescore_topk": (_c_int, [_vp, _c_int, _vp, _c_i64, _c_int, _vp, _c_int, C.c_int32, _vp, _vp, _c_int, _vp]),
    "cir_qe_aggregate": (_c_int, [_vp, _c_int, _vp, _c_i64, _c_int, _vp, _vp, _c_int, _c_int, _c_int, _c_f, _c_i64,
                                  _c_f, _vp, _vp]),
    "cir_eval_ap": (_c_int, [_vp, _c_int, _c_i64, _c_i64, _vp, _vp, _vp, _vp, _vp, _c_int, _vp, _vp, _vp]),
    "cir_tuple_loss_workspace_bytes": (_c_int, [_c_int, _szp]),
    "cir_tuple_loss": (_c_int, [_vp, _c_i64, _c_int, _c_int, _c_int, _vp, _c_int, _c_f, _c_f, _vp, _vp, _c_i64, _vp,
                                C.c_size_t, _vp]),
    "cir_l2n_bwd_rows": (_c_int, [_vp, _vp, _c_i64, _c_int, _c_f, _vp, _vp, _vp]),
    "cir_colsum_rows": (_c_int, [_vp, _c_i64, _c_int, _vp, _vp]),
    "cir_gem_dp": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp, C.c_size_t, _vp]),
    "cir_mine_filter": (_c_int, [_vp, _c_int, _c_int, _vp, _c_i64, _vp, _c_int, _vp, _vp, _c_int, _vp, _vp, _vp, _vp, _vp, _c_f,
                                 _vp, _vp]),
}

CIR_POOL_GEM, CIR_POOL_MAC, CIR_POOL_SPOC = 0, 1, 2
CIR_TAIL_NO_WHITEN, CIR_TAIL_POOL_ONLY, CIR_TAIL_ACCUMULATE, CIR_TAIL_HINT_INTEGER_P = 1, 2, 4, 8

_lib = None
_lock = threading.Lock()


class CirError(RuntimeError):
    pass


def load():
    """Load the shared library and bind every declared entry point (raises if any is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise CirError(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "cirtorch_b200 has no CPU / eager fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().cir_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise CirError(f"{what} failed ({rc}): {msg}")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_of(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise CirError("cirtorch_b200 ops run on CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)


def launch_count(reset: bool = False) -> int:
    return int(load().cir_launch_count(1 if reset else 0))


# ----------------------------------------------------------------------------- workspaces
_ws = {}


def workspace(device: torch.device, nbytes: int, tag: str = "default") -> torch.Tensor:
    """Caller-owned scratch the library asks for; cached per (device, stream, tag) and grown on demand."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def release_workspaces():
    _ws.clear()
