"""ctypes binding of libsafevla_b200.so (the C ABI declared in include/safevla_b200.h).

There is no CPU fallback: `get_ctx()` raises if the library is missing, if no CUDA device is
visible, or if the device is not sm_100.  Loading the library itself (symbol checks) works
without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsafevla_b200.so")

F32, BF16 = 0, 1
EPI_NONE, EPI_RELU, EPI_RELU_MASK, EPI_GELU, EPI_RELU_BITS, EPI_MASK_BITS = 0, 1, 2, 3, 4, 5
ATTN_FULL, ATTN_TRAJ_CAUSAL, ATTN_T5_BIAS = 0, 1, 2
PPO_NSCALARS = 16

c_p = C.c_void_p
c_ll = C.c_longlong


class PpoHparams(C.Structure):
    _fields_ = [("clip_param", C.c_float), ("w_action", C.c_float), ("w_value", C.c_float),
                ("w_entropy", C.c_float), ("w_cvalue", C.c_float), ("inv_count", C.c_float),
                ("grad_scale", C.c_float), ("use_clipped_value_loss", C.c_int), ("use_lagrangian", C.c_int)]


class AdamHparams(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("max_grad_norm", C.c_float), ("grad_prescale", C.c_float), ("step", C.c_int),
                ("zero_grad", C.c_int)]


class Dropout(C.Structure):
    """svla_dropout: counter-based mask spec (p, seed, site, step, row0)."""
    _fields_ = [("p", C.c_float), ("seed", C.c_ulonglong), ("site", C.c_uint), ("step", C.c_uint), ("row0", C.c_uint),
                ("row_stride", C.c_uint)]


class GemmDesc(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("A", c_p), ("lda", c_ll), ("transA", C.c_int),
                ("B", c_p), ("ldb", c_ll), ("transB", C.c_int),
                ("C", c_p), ("ldc", c_ll),
                ("dtypeA", C.c_int), ("dtypeB", C.c_int), ("dtypeC", C.c_int),
                ("bias", c_p),
                ("residual", c_p), ("ldr", c_ll), ("dtypeR", C.c_int),
                ("aux", c_p), ("ldaux", c_ll), ("dtypeAux", C.c_int),
                ("epilogue", C.c_int), ("accumulate", C.c_int), ("alpha", C.c_float), ("impl", C.c_int),
                ("colsum_a", c_p), ("dropout", C.POINTER(Dropout))]


class RowMap(C.Structure):
    _fields_ = [("group", C.c_int), ("group_stride", C.c_int), ("group_offset", C.c_int)]


IDENT = RowMap(0, 0, 0)

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/safevla_b200.h
PROTOTYPES: Dict[str, list] = {
    "svla_ctx_create": [C.c_int, C.POINTER(c_p)],
    "svla_ctx_destroy": [c_p],
    "svla_last_error": [],
    "svla_version": [],
    "svla_sm_count": [c_p],
    "svla_launch_count": [],
    "svla_launch_count_add": [C.c_ulonglong],
    "svla_gae_dual": [c_p] + [c_p] * 9 + [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, c_p],
    "svla_discounted_returns_dual": [c_p] + [c_p] * 9 + [C.c_int, C.c_int, C.c_double, c_p],
    "svla_normalize_advantage": [c_p, c_p, c_p, c_p, c_ll, c_p],
    "svla_advantage_sums": [c_p, c_p, c_ll, c_p, c_p],
    "svla_normalize_advantage_from_sums": [c_p, c_p, c_p, c_p, c_p, c_ll, c_p],
    "svla_ppo_lag_fwd_bwd": [c_p] + [c_p] * 12 + [C.POINTER(PpoHparams)] + [c_p] * 4 + [c_ll, C.c_int, c_p],
    "svla_lagrange_update": [c_p, c_p, c_p, c_p, C.c_float, C.c_float, C.c_float, c_p],
    "svla_sq_norm": [c_p, c_p, c_ll, c_p, c_p],
    "svla_clip_adam": [c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_p, C.POINTER(AdamHparams), c_p],
    "svla_gemm": [c_p, C.POINTER(GemmDesc), c_p],
    "svla_gemm_which": [C.POINTER(GemmDesc)],
    "svla_colsum": [c_p, c_p, C.c_int, c_ll, C.c_int, c_ll, c_p, C.c_int, c_p],
    "svla_layernorm_fwd": [c_p, c_p, c_p, C.c_int, c_p, c_p, c_p, C.c_int, C.c_float, c_p, C.c_int, RowMap,
                           c_p, c_p, c_ll, C.c_int, c_p],
    "svla_layernorm_bwd": [c_p, c_p, C.c_int, RowMap, c_p, c_p, C.c_int, c_p, c_p, C.c_int, c_p, c_p, c_p,
                           C.c_int, c_p, c_p, c_p, c_ll, C.c_int, c_p],
    "svla_rmsnorm_fwd": [c_p, c_p, C.c_int, c_p, C.c_float, c_p, C.c_int, c_p, c_ll, C.c_int, c_p],
    "svla_rmsnorm_bwd": [c_p, c_p, C.c_int, c_p, C.c_int, c_p, c_p, c_p, C.c_int, C.c_int, c_p, c_ll, C.c_int, c_p],
    "svla_attn_fwd": [c_p, C.c_int, c_p, c_p, c_p, c_ll, c_p, c_ll, C.c_int, c_p, c_p, c_p, c_p,
                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_p],
    "svla_attn_bwd": [c_p, C.c_int, c_p, c_p, c_p, c_ll, c_p, c_p, c_ll, c_p, c_p, c_p, c_ll, C.c_int, c_p, c_p,
                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_p],
    "svla_set_attn_impl": [C.c_int],
    "svla_attn_cls_fwd": [c_p, c_p, c_ll, c_p, c_p, c_ll, c_p, c_ll, C.c_int, c_p, C.c_int, C.c_int, C.c_int, C.c_int,
                          C.c_float, c_p, c_p],
    "svla_attn_cls_bwd": [c_p, c_p, c_ll, c_p, c_p, c_ll, c_p, c_p, c_ll, c_p, c_ll, c_p, c_p, c_ll, C.c_int, c_p,
                          C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_p, c_p],
    "svla_patchify_u8": [c_p, c_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_p, c_p, c_p, C.c_int, C.c_int,
                         c_p],
    "svla_vit_assemble": [c_p, c_p, c_p, c_p, c_p, C.c_int, C.c_int, C.c_int, C.c_int, c_p],
    "svla_tokens_pool": [c_p, c_p, C.c_int, c_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_p],
    "svla_hl_gauss_fwd_bwd": [c_p, c_p, c_ll, c_p, c_p, C.c_int, C.c_float, C.c_float, c_p, c_p, c_p, c_ll, c_p],
    "svla_attn_decode": [c_p, c_p, c_ll, c_p, c_p, c_ll, c_ll, c_p, C.c_int, c_p, c_ll, C.c_int, C.c_int, C.c_int,
                         C.c_int, C.c_float, C.c_int, c_p],
    "svla_swiglu_fwd": [c_p, c_p, c_p, C.c_int, c_ll, C.c_int, c_p],
    "svla_swiglu_bwd": [c_p, c_p, c_p, c_p, C.c_int, c_ll, C.c_int, c_p],
    "svla_embed_time_fwd": [c_p, c_p, C.c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                            C.c_int, C.c_int, C.c_int, C.c_int, c_p],
    "svla_embed_time_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, C.c_int, c_p, c_p, C.c_int, C.c_int, C.c_int, C.c_int, c_p],
    "svla_nchw_to_tokens": [c_p, c_p, c_p, C.c_int, c_ll, C.c_int, C.c_int, c_p],
    "svla_copy_rows": [c_p, c_p, C.c_int, c_ll, RowMap, c_p, c_p, C.c_int, c_ll, RowMap, c_ll, C.c_int, C.c_int, c_p],
    "svla_fill_rows": [c_p, c_p, c_p, C.c_int, c_ll, RowMap, c_ll, C.c_int, c_p],
    "svla_scale_by": [c_p, c_p, c_ll, c_p, c_p],
    "svla_cast_bf16": [c_p, c_p, c_p, c_ll, c_p],
    "svla_attn_split_fwd": [c_p, C.c_int, c_p, c_p, c_p, c_ll, c_ll, c_p, c_ll, c_p, c_p, C.c_int, C.c_int, C.c_int, C.c_int,
                            C.c_float, c_p],
    "svla_attn_split_bwd": [c_p, C.c_int, c_p, c_p, c_p, c_ll, c_ll, c_p, c_ll, c_ll, c_p, c_p, c_p, c_ll, c_p, c_p,
                            C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_p],
    "svla_dropout_rows": [c_p, c_p, C.c_int, c_ll, c_p, C.c_int, c_ll, c_ll, C.c_int, C.POINTER(Dropout), c_p],
    "svla_attn_drop_fwd": [c_p, C.c_int, c_p, c_p, c_p, c_ll, c_p, c_ll, c_p, c_p, C.c_int, C.c_int, C.c_int, C.c_int,
                           C.c_float, C.POINTER(Dropout), c_p],
    "svla_attn_drop_bwd": [c_p, C.c_int, c_p, c_p, c_p, c_ll, c_p, c_p, c_ll, c_p, c_p, c_p, c_ll, c_p, c_p, C.c_int,
                           C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(Dropout), c_p],
    "svla_split_concat": [c_p, c_p, c_ll, c_ll, C.c_int, c_p, c_ll, C.c_int, C.c_int, C.POINTER(C.c_int), c_p],
    "svla_hash_rows": [c_p, c_p, c_ll, C.c_int, c_p, c_p],
    "svla_episode_cost_step": [c_p, c_p, c_p, c_p, c_p, C.c_int, c_p],
    "svla_combine_cost_advantages": [c_p, c_p, c_p, C.c_int, c_ll, c_p, c_p, c_p],
}
_RESTYPES = {"svla_last_error": C.c_char_p, "svla_launch_count": C.c_ulonglong}

_lib: Optional[C.CDLL] = None
_ctxs: Dict[int, int] = {}


def load_library() -> C.CDLL:
    """dlopen the in-tree library and bind every prototype; raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m safevla_b200.build` "
            "(safevla_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


class _Profiled:
    """Wraps one C entry point with CUDA events on the current stream (bench/profiling only)."""

    def __init__(self, name, fn, sink):
        self.name, self.fn, self.sink = name, fn, sink

    def __call__(self, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = self.fn(*args)
        e1.record()
        self.sink.append((self.name, e0, e1))
        return rc


_prof_sink = None
_SKIP_PROFILE = {"svla_ctx_create", "svla_ctx_destroy", "svla_last_error", "svla_version", "svla_sm_count",
                 "svla_launch_count", "svla_launch_count_add", "svla_gemm_which", "svla_set_attn_impl"}


def profile_start():
    """Start recording per-entry-point device time (CUDA events around every C-ABI call)."""
    global _prof_sink
    lib = load_library()
    _prof_sink = []
    for name in PROTOTYPES:
        if name in _SKIP_PROFILE:
            continue
        fn = getattr(lib, name)
        if not isinstance(fn, _Profiled):
            setattr(lib, name, _Profiled(name, fn, _prof_sink))
        else:
            fn.sink = _prof_sink


def profile_stop():
    """Stop recording; returns {entry point: (total ms, calls)} (synchronises the device)."""
    global _prof_sink
    lib = load_library()
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in _prof_sink or []:
        ms, n = out.get(name, (0.0, 0))
        out[name] = (ms + e0.elapsed_time(e1), n + 1)
    for name in PROTOTYPES:
        fn = getattr(lib, name, None)
        if isinstance(fn, _Profiled):
            setattr(lib, name, fn.fn)
    _prof_sink = None
    return out


def last_error() -> str:
    return (load_library().svla_last_error() or b"").decode()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"libsafevla_b200 {what} failed (rc={rc}): {last_error()}")


def get_ctx(device: Optional[int] = None) -> int:
    """One svla_ctx per (process, device, stream): a context owns the scratch its kernels fold their deterministic
    reductions in (block partials, tickets, split-K slices), so launches that may run concurrently -- the towers on
    their own streams -- must not share one."""
    lib = load_library()
    if not torch.cuda.is_available():
        raise RuntimeError("safevla_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ctx = _ctxs.get(key)
    if ctx is None:
        out = c_p()
        check(lib.svla_ctx_create(int(device), C.byref(out)), "svla_ctx_create")
        ctx = _ctxs[key] = out.value
    return ctx


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")
