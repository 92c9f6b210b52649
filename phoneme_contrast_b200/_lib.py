"""ctypes binding of libpc_b200.so (C ABI declared in include/phoneme_contrast.h).

There is no CPU fallback: if the library is missing, or a tensor is not a contiguous CUDA tensor of the
expected dtype, the call raises. PC_EINVAL becomes ValueError (the exception type the reference raises for
bad shapes, e.g. losses.py:44-45); every other failure becomes RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpc_b200.so")

PC_OK, PC_EINVAL, PC_EUNSUPPORTED, PC_ECUDA = 0, -1, -2, -3
PC_FB_MAXW = 32
FE_MFCC, FE_LOGMEL = 0, 1
CLAMP_PER_CLIP, CLAMP_NONE, CLAMP_GIVEN = 0, 1, 2
PREC_FP32, PREC_TF32X3, PREC_BF16, PREC_FP16X2 = 0, 1, 2, 3

vp, i32, i64, f32, f64, u64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_uint64, C.c_size_t


class PcViewDesc(C.Structure):
    _fields_ = [("gain", f32), ("t0", C.c_int32), ("t1", C.c_int32), ("f0", C.c_int32), ("f1", C.c_int32),
                ("noise_level", f32), ("noise_seed", C.c_uint32), ("clip", C.c_int32)]


class PcMfccConsts(C.Structure):
    _fields_ = [("window", vp), ("fb_start", vp), ("fb_len", vp), ("fb_w", vp), ("dct", vp), ("tw", vp),
                ("n_fft", C.c_int32), ("hop", C.c_int32), ("n_mels", C.c_int32), ("n_mfcc", C.c_int32), ("fb_wmax", C.c_int32),
                ("preemph", C.c_float)]


class PcConvGeom(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "H", "W", "Cin", "Ho", "Wo", "Cout", "R", "S", "stride", "pad")]


class PcInXform(C.Structure):
    _fields_ = [("scale", vp), ("shift", vp), ("drop", vp), ("relu", C.c_int32), ("presplit", C.c_int32)]


class PcBnFinalize(C.Structure):
    _fields_ = [("stats", vp), ("count", C.c_double), ("gamma", vp), ("beta", vp), ("running_mean", vp), ("running_var", vp),
                ("num_batches_tracked", vp), ("momentum", C.c_float), ("eps", C.c_float), ("scale", vp), ("shift", vp), ("mean", vp),
                ("invstd", vp)]


class PcBnBwdReduce(C.Structure):
    _fields_ = [("y", vp), ("scale", vp), ("shift", vp), ("mean", vp), ("invstd", vp), ("drop", vp), ("sums", vp), ("maxes", vp)]


class PcPackJob(C.Structure):
    _fields_ = [("w_oihw", vp), ("out", vp)] + [(n, C.c_int32) for n in ("O", "I", "R", "S", "dgrad", "prec")] + [("item_begin", C.c_int64)]


# name -> (restype, argtypes); must list every symbol include/phoneme_contrast.h declares (tests check this)
SIGNATURES = {
    "pc_last_error": (C.c_char_p, []),
    "pc_abi_version": (i32, []),
    "pc_launch_count": (C.c_ulonglong, []),
    "pc_frontend_fwd": (i32, [vp, i32, i32, i32, C.POINTER(PcMfccConsts), vp, i32, vp, i32, i32, f32, vp, vp, vp, vp]),
    "pc_reduce_max": (i32, [vp, i32, vp, vp]),
    "pc_augment_apply": (i32, [vp, vp, i32, i32, i32, vp, vp, vp]),
    "pc_compute_deltas": (i32, [vp, i32, i32, vp, vp]),
    "pc_supcon_fwd": (i32, [vp, vp, vp, i32, i32, i32, i32, f32, f32, vp, vp, vp]),
    "pc_sum_scaled": (i32, [vp, i32, f32, vp, vp]),
    "pc_dp_pack": (i32, [vp, vp, i32, i32, vp, vp]),
    "pc_dp_unpack": (i32, [vp, i32, i32, vp, vp, vp]),
    "pc_supcon_loss_from_stats": (i32, [vp, i32, f32, f32, f32, vp, vp]),
    "pc_supcon_bwd": (i32, [vp, vp, vp, i32, i32, i32, i32, f32, f32, vp, vp, vp, vp]),
    "pc_supcon_tc_supported": (i32, [i32, i32, i32, i32]),
    "pc_supcon_tc_workspace": (sz, [i32, i32, i32]),
    "pc_supcon_fwd_tc": (i32, [vp, vp, i32, i32, i32, i32, f32, f32, vp, sz, vp, vp, vp]),
    "pc_supcon_bwd_tc_workspace": (sz, [i32, i32, i32]),
    "pc_supcon_bwd_tc": (i32, [vp, vp, i32, i32, i32, i32, f32, f32, vp, vp, vp, sz, vp, vp]),
    "pc_pack_conv_weight": (i32, [vp, i32, i32, i32, i32, vp, vp, vp]),
    "pc_conv_tc_supported": (i32, [C.POINTER(PcConvGeom), i32, i32]),
    "pc_conv_tc_packed_bytes": (sz, [i32, i32, i32, i32, i32, i32]),
    "pc_pack_conv_weight_tc": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    "pc_pack_conv_weight_tc_items": (i64, [i32, i32, i32, i32, i32, i32]),
    "pc_pack_conv_weights_tc_batch": (i32, [vp, i32, i64, vp]),
    "pc_tc_gemm_workspace": (sz, [i32, i32, i32]),
    "pc_tc_gemm": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp, sz, vp]),
    "pc_conv_fwd": (i32, [vp, vp, vp, C.POINTER(PcConvGeom), C.POINTER(PcInXform), vp, vp, i32, vp]),
    "pc_conv_dgrad": (i32, [vp, vp, C.POINTER(PcConvGeom), vp, i32, i32, vp, i32, vp]),
    "pc_conv_wgrad_halo_supported": (i32, [C.POINTER(PcConvGeom)]),
    "pc_conv_wgrad_halo_workspace": (sz, [C.POINTER(PcConvGeom)]),
    "pc_conv_wgrad_halo": (i32, [vp, vp, C.POINTER(PcConvGeom), vp, vp, sz, vp, vp]),
    "pc_conv_halo_supported": (i32, [C.POINTER(PcConvGeom), i32]),
    "pc_conv_fwd_halo": (i32, [vp, vp, vp, C.POINTER(PcConvGeom), vp, vp, vp]),
    "pc_conv_dgrad_halo": (i32, [vp, vp, C.POINTER(PcConvGeom), vp, i32, vp, vp]),
    "pc_conv_dgrad_halo_bnred": (i32, [vp, vp, C.POINTER(PcConvGeom), vp, vp, C.POINTER(PcBnBwdReduce), vp]),
    "pc_conv_wgrad_workspace": (sz, [C.POINTER(PcConvGeom)]),
    "pc_conv_wgrad": (i32, [vp, vp, C.POINTER(PcConvGeom), C.POINTER(PcInXform), vp, vp, vp, sz, i32, vp, i32, vp]),
    "pc_stem_fwd_supported": (i32, [i32, i32, i32, i32]),
    "pc_stem_stats_from_gram": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, C.POINTER(PcBnFinalize), vp]),
    "pc_stem_fwd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]),
    "pc_stem_bwd_supported": (i32, [i32, i32, i32, i32]),
    "pc_stem_gram": (i32, [vp, i32, i32, i32, vp, vp, vp]),
    "pc_stem_bwd_workspace": (sz, []),
    "pc_stem_bwd": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp, vp, vp, vp, vp]),
    "pc_bn_finalize": (i32, [vp, i32, f64, vp, vp, vp, vp, vp, f32, f32, i32, vp, vp, vp, vp, vp]),
    "pc_bn_act_fwd": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp]),
    "pc_bn_act_bwd_reduce": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp]),
    "pc_bn_act_split": (i32, [vp, i64, i32, i32, vp, vp, vp, i32, vp, vp]),
    "pc_f16_overflow_query": (i32, [i32, C.POINTER(C.c_int), vp]),
    "pc_bn_act_split_fin": (i32, [vp, i64, i32, i32, C.POINTER(PcBnFinalize), vp, i32, vp, vp]),
    "pc_bn_act_fwd_fin": (i32, [vp, i32, i32, i32, i32, C.POINTER(PcBnFinalize), vp, i32, vp, vp, vp, vp]),
    "pc_bn_add_relu_fwd_fin": (i32, [vp, C.POINTER(PcBnFinalize), vp, C.POINTER(PcBnFinalize), i64, i32, vp, vp, vp]),
    "pc_bn_act_bwd_apply": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "pc_bn_add_relu_fwd": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, vp, vp, vp]),
    "pc_bn_add_relu_bwd_reduce": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp, vp, vp, vp]),
    "pc_bn_add_relu_bwd_apply": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "pc_attn_pool_fwd": (i32, [vp, i32, i32, i32, vp, vp, vp, vp, vp]),
    "pc_attn_pool_splits": (i32, [i32]),
    "pc_attn_pool_fwd_ws": (i32, [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    "pc_attn_pool_bwd": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]),
    "pc_head_workspace": (sz, [i32, i32, i32]),
    "pc_head_fwd": (i32, [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, f32, f32, i32, vp, vp, vp]),
    "pc_head_bwd": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]),
    "pc_dropout2d_mask": (i32, [vp, i32, i32, f32, u64, u64, vp, vp]),
    "pc_clip_adam_dev": (i32, [vp, vp, vp, vp, i64, vp, f32, f32, f32, f32, f32, vp, f32, vp, vp]),
    "pc_counter_add": (i32, [vp, i64, vp]),
    "pc_grad_sumsq": (i32, [vp, i64, vp, vp]),
    "pc_clip_adam": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, f32, vp, f32, i64, vp]),
    "pc_peer_flag_bytes": (i32, []),
    "pc_peer_max_ranks": (i32, []),
    "pc_peer_export": (i32, [vp, vp, C.POINTER(sz)]),
    "pc_peer_open": (i32, [vp, C.POINTER(vp)]),
    "pc_peer_alloc": (i32, [sz, C.POINTER(vp)]),
    "pc_peer_free": (i32, [vp]),
    "pc_peer_close": (i32, [vp]),
    "pc_peer_barrier": (i32, [vp, i32, i32, sz, i32, i32, vp]),
    "pc_peer_error": (i32, [vp, i32, i32, sz, i32, C.POINTER(C.c_int), vp]),
    "pc_dp_gather_peer": (i32, [vp, vp, i32, i32, vp, i32, sz, sz, i32, vp]),
    "pc_peer_bcast": (i32, [vp, sz, vp, i32, sz, vp]),
    "pc_peer_allreduce": (i32, [vp, i32, i32, sz, C.c_longlong, i32, vp]),
    "pc_peer_sum_slots": (i32, [vp, i32, i32, f64, vp, vp]),
    "pc_head_fwd_sync": (i32, [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp, f64, i32, vp]),
    "pc_head_bwd_sync": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
}

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle. Raises loudly when the CUDA library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} not found: build it with `python -m phoneme_contrast_b200.build` "
                "(there is no CPU / PyTorch fallback for this path)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def last_error() -> str:
    return lib().pc_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc == PC_OK:
        return
    msg = last_error()
    if rc == PC_EINVAL:
        raise ValueError(msg)
    if rc == PC_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


# Optional per-entry-point device timing (bench.py's live roofline measurement): CUDA events recorded on the
# launching stream around every C-ABI call while a profile is open.
_prof = None


def profile_begin() -> None:
    global _prof
    _prof = {"events": {}, "work": {}}


def profile_end() -> dict:
    """-> {entry point: {"ms": total device ms, "calls": n, "work": accumulated algorithmic work}}."""
    global _prof
    p, _prof = _prof, None
    torch.cuda.synchronize()
    out = {}
    for name, evs in p["events"].items():
        out[name] = {"ms": sum(a.elapsed_time(b) for a, b in evs), "calls": len(evs), "work": p["work"].get(name, 0.0)}
    return out


def note_work(name: str, amount: float) -> None:
    if _prof is not None:
        _prof["work"][name] = _prof["work"].get(name, 0.0) + amount


def call(name: str, *args):
    if _prof is None:
        check(getattr(lib(), name)(*args))
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(lib(), name)(*args))
    e1.record()
    _prof["events"].setdefault(name, []).append((e0, e1))


def stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t, dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor (None -> NULL). No silent copies, no CPU tensors."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError("phoneme_contrast_b200 kernels run on CUDA tensors only (no CPU fallback); got a CPU tensor")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def f16_overflow(reset: bool = True) -> bool:
    """True when an FP16X2 activation-plane writer met |a| > 65504 (or a NaN) since the last reset. Synchronises the stream."""
    flag = C.c_int(0)
    check(lib().pc_f16_overflow_query(1 if reset else 0, C.byref(flag), stream()))
    return flag.value != 0


def launch_count() -> int:
    return int(lib().pc_launch_count())
