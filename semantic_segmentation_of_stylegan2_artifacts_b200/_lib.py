"""ctypes binding of libmsunet_sm100.so (the C ABI declared in include/msunet_b200.h).

There is deliberately no fallback: if the shared library is missing or a kernel reports an
error the call raises.  PyTorch is used only for device memory, streams and autograd plumbing.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsunet_sm100.so")

F32, BF16, F16 = 0, 1, 2
MAP_NONE, MAP_WINDOW, MAP_SHUFFLE, MAP_CONV3, MAP_MERGE, MAP_UNSHUFFLE = 0, 1, 2, 3, 4, 5

_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}


def dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}") from None


class MsuOperand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ptr2", C.c_void_p), ("ld", C.c_int64), ("ld2", C.c_int64),
                ("k_split", C.c_int32), ("orient", C.c_int32), ("map", C.c_int32), ("dtype", C.c_int32),
                ("geo", C.c_int32 * 6), ("rowscale", C.c_void_p), ("rows_per_sample", C.c_int32),
                ("_pad", C.c_int32)]


class MsuEpilogue(C.Structure):
    _fields_ = [("C", C.c_void_p), ("Cpre", C.c_void_p), ("bias", C.c_void_p), ("R", C.c_void_p),
                ("H", C.c_void_p), ("ldc", C.c_int64), ("ldr", C.c_int64), ("ldh", C.c_int64),
                ("rowscale", C.c_void_p), ("rows_per_sample", C.c_int32), ("act", C.c_int32),
                ("map", C.c_int32), ("dtype", C.c_int32), ("geo", C.c_int32 * 6), ("out_f32", C.c_int32),
                ("accumulate", C.c_int32), ("colsum", C.c_void_p),
                ("lnd_gamma", C.c_void_p), ("lnd_beta", C.c_void_p), ("lnd_w", C.c_void_p), ("lnd_logits", C.c_void_p),
                ("lnd_mean", C.c_void_p), ("lnd_rstd", C.c_void_p), ("lnd_m2", C.c_void_p)]


_lib: Optional[C.CDLL] = None

_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_SIGS = {
    "msu_adamw_chunk": [],
    "msu_adamw_step": [_P, _P, _P, _I32, _P, _P, _P],
    "msu_shadow_blocks": [C.c_int, _I64, _I64],
    "msu_refresh_shadows": [_P, _P, _P, _I32, _P],
    "msu_stage_u8": [_P, _P, _P, _P, _P, _I32, _I32, _I32, _P],
    "msu_gemm": [C.POINTER(MsuOperand), C.POINTER(MsuOperand), C.POINTER(MsuEpilogue), _I64, _I64, _I64, _P, _I64,
                 C.c_int, _P],
    "msu_colsum": [C.POINTER(MsuOperand), _I64, _I64, _P, C.c_int, _P, _I64, _P],
    "msu_ln_fwd": [C.c_int, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _P],
    "msu_ln_bwd_partial_rows": [C.c_int, _I64, _I32],
    "msu_ln_bwd": [C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _P, _P],
    "msu_ln_bwd_dual": [C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P, _P, _I32, _P, _P],
    "msu_ln_param_reduce": [_P, _I32, _I32, _P, _P, _P, C.c_int, _P],
    "msu_winattn_fwd": [C.c_int, _P, _P, _P, _I64, _I32, _P, C.c_float, _P, _P, _P],
    "msu_set_attn_backend": [C.c_int],
    "msu_set_deterministic": [C.c_int],
    "msu_winattn_bwd_grid": [C.c_int, _I64, _I32],
    "msu_winattn_bwd_direct": [C.c_int],
    "msu_winattn_bwd": [C.c_int, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P, C.c_float, _P, _P, _P],
    "msu_relbias_expand": [_P, _P, _I32, _P],
    "msu_relbias_reduce": [_P, _I32, _I32, _P, C.c_int, _P],
    "msu_prep_weight": [C.c_int, C.c_int, _P, _P, _I64, _I64, _P],
    "msu_patchify4": [C.c_int, _P, _P, _I32, _I32, _P],
    "msu_loss_fwd": [C.c_int, _P, _P, _I32, _I64, _F, _F, _F, _P, _P, _P, _P, _P],
    "msu_loss_per_sample": [C.c_int, _P, _P, _I32, _I64, _F, _F, _F, _P, _P, _P],
    "msu_loss_bwd": [C.c_int, _P, _P, _I32, _I64, _F, _F, _F, _P, _P, _P, _P, _P],
    "msu_metrics": [C.c_int, C.c_int, _P, _P, _P, _I32, _I64, _F, _P, _P, _P, _P, _P, _P],
    "msu_gather_rows": [C.POINTER(MsuOperand), _P, _I64, _I64, _P],
    "msu_cast": [C.c_int, C.c_int, _P, _P, _I64, _P],
    "msu_add": [C.c_int, _P, _P, _P, _I64, _P],
    "msu_version": [],
    "msu_struct_size": [C.c_int],
    "msu_launch_count": [],
    "msu_last_gemm_backend": [],
}
EXPORTS = tuple(_SIGS) + ("msu_last_error_string",)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the MS-UNet hot path.")
        l = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = C.c_longlong if name == "msu_launch_count" else C.c_int
        l.msu_last_error_string.restype = C.c_char_p
        l.msu_last_error_string.argtypes = []
        if l.msu_struct_size(0) != C.sizeof(MsuOperand) or l.msu_struct_size(1) != C.sizeof(MsuEpilogue):
            raise RuntimeError("libmsunet_sm100.so struct layout does not match the Python binding; rebuild it")
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().msu_last_error_string().decode(errors="replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def geo6(g: Optional[Sequence[int]]):
    arr = (C.c_int32 * 6)()
    if g is not None:
        for i, v in enumerate(g):
            arr[i] = int(v)
    return arr


def launch_count() -> int:
    return int(lib().msu_launch_count())
