"""Tensor-level wrappers over the C ABI (one Python function per entry point of include/phoneme_contrast.h).

torch is used here only for device memory (torch.empty) and the current stream; every computation is a
kernel of libpc_b200.so. Activations are NHWC fp32.
"""
from __future__ import annotations

import ctypes as C

import os

import torch

from . import _lib as L
from ._lib import PcConvGeom, PcInXform, call, ptr, stream

F32 = torch.float32


def conv_geom(B, H, W, Cin, Cout, k, stride, pad) -> PcConvGeom:
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    return PcConvGeom(B, H, W, Cin, Ho, Wo, Cout, k, k, stride, pad)


def _xf(scale=None, shift=None, drop=None, relu=False, presplit=False):
    if scale is None and drop is None and not relu and not presplit:
        return None
    return PcInXform(ptr(scale), ptr(shift), ptr(drop), 1 if relu else 0, 1 if presplit else 0)


def bn_act_split(y, scale=None, shift=None, drop=None, relu=False):
    """a = drop * relu(scale*y + shift) of an NHWC tensor, written once as fp16 hi | lo planes (uint8 view [2, numel*2]) for the
    FP16X2 convolutions (`presplit=True` in their xform)."""
    B, H, W, C_ = y.shape
    planes = torch.empty(2, y.numel() * 2, device=y.device, dtype=torch.uint8)
    call("pc_bn_act_split", ptr(y), B * H * W, C_, H * W, ptr(scale), ptr(shift), ptr(drop), 1 if relu else 0,
         ptr(planes, torch.uint8), stream())
    return planes


def pack_conv_weight(w: torch.Tensor, want_fwd=True, want_dgrad=True):
    O, I, R, S = w.shape
    wf = torch.empty(R * S * I, O, device=w.device, dtype=F32) if want_fwd else None
    wd = torch.empty(R * S * O, I, device=w.device, dtype=F32) if want_dgrad else None
    call("pc_pack_conv_weight", ptr(w), O, I, R, S, ptr(wf), ptr(wd), stream())
    return wf, wd


def tc_supported(g: PcConvGeom, dgrad: bool, prec: int, planes_ok: bool = False) -> bool:
    """planes_ok: the gathered operand will be supplied as pre-split fp16 planes. Layers with 32 gathered channels reach the FP16X2
    engine only that way (halo path: the TMA boxes zero-fill the missing half of the 64-channel chunk)."""
    if prec == L.PREC_FP32 or not L.lib().pc_conv_tc_supported(C.byref(g), 1 if dgrad else 0, prec):
        return False
    ca = g.Cout if dgrad else g.Cin
    return planes_ok or prec != L.PREC_FP16X2 or ca % 64 == 0


def pack_conv_weight_tc(w: torch.Tensor, dgrad: bool, prec: int) -> torch.Tensor:
    """Pre-swizzled tensor-core weight tiles (pc_pack_conv_weight_tc) for the forward (dgrad=False) or data-gradient GEMM."""
    O, I, R, S = w.shape
    nbytes = int(L.lib().pc_conv_tc_packed_bytes(O, I, R, S, 1 if dgrad else 0, prec))
    out = torch.empty(nbytes, device=w.device, dtype=torch.uint8)
    call("pc_pack_conv_weight_tc", ptr(w), O, I, R, S, 1 if dgrad else 0, prec, ptr(out, torch.uint8), stream())
    return out


class ConvWeights:
    """Per-step packed forms of one conv weight: forward operand, data-gradient operand, and the precision each runs in
    (tensor cores when the layer is eligible, exact-fp32 SIMT otherwise). With a WeightPacker the tensor-core operands of
    all layers are written by one batched launch at the start of the step instead of one launch per operand."""
    __slots__ = ("wf", "wd", "prec_f", "prec_d")

    def __init__(self, w: torch.Tensor, g: PcConvGeom, prec: int, need_dgrad: bool = True, packer=None, planes_ok: bool = False):
        """planes_ok: the caller supplies this layer's gathered operands as pre-split fp16 planes -- required for the layers that reach the
        FP16X2 engine only through the halo path (32 gathered channels: the TMA boxes zero-fill the missing half of the 64-channel chunk)."""
        import os
        if packer is not None and packer.replaying:
            self.wf, self.wd, self.prec_f, self.prec_d = packer.next(w, need_dgrad)
            return
        skip = os.environ.get("PC_TC_SKIP", "")
        tag = f"{g.Cin}-{g.Cout}-{g.R}-{g.stride}"

        def pick(dgrad):
            # FP16X2 needs 64-channel k-chunks; a layer with only 32-channel granularity runs the TF32x3 engine instead
            for cand in ((prec, L.PREC_TF32X3) if prec == L.PREC_FP16X2 else (prec,)):
                if tc_supported(g, dgrad, cand, planes_ok):
                    return cand
            return L.PREC_FP32
        self.prec_f = pick(False) if (os.environ.get("PC_TC_FWD", "1") == "1" and tag not in skip.split(",")) else L.PREC_FP32
        self.prec_d = pick(True) if os.environ.get("PC_TC_DGRAD", "1") == "1" else L.PREC_FP32
        self.wf = self.wd = None
        need_simt_f = self.prec_f == L.PREC_FP32
        need_simt_d = need_dgrad and self.prec_d == L.PREC_FP32
        if need_simt_f or need_simt_d:
            wf, wd = pack_conv_weight(w, need_simt_f, need_simt_d)
            self.wf, self.wd = wf, wd
        if not need_simt_f:
            self.wf = pack_conv_weight_tc(w, False, self.prec_f)
        if need_dgrad and not need_simt_d:
            self.wd = pack_conv_weight_tc(w, True, self.prec_d)
        if packer is not None:
            packer.record(w, need_dgrad, self, need_simt_f, need_simt_d)


class _PackerState:
    __slots__ = ("entries", "table", "n_jobs", "total")

    def __init__(self):
        self.entries, self.table, self.n_jobs, self.total = [], None, 0, 0


class WeightPacker:
    """Batches the per-step weight re-packing of a network. The first forward with a given key records, layer by layer,
    which operands are needed (and packs them one launch each, as without a packer); later forwards with the same key
    issue ONE pc_pack_conv_weights_tc_batch launch that refreshes all recorded tensor-core operands in place and hand the
    same buffers out in recording order. Exact-fp32 SIMT operands (ineligible layers) are still packed per layer.

    One recording is kept PER KEY (train / eval, every batch shape) and never freed while the packer lives: a captured CUDA
    graph bakes the job-table and operand-buffer addresses of the recording it was captured with into its launches, so an
    interleaved eval forward or odd-shaped batch must not release them (a recording invalidated by re-allocated parameters
    is retired, not dropped, for the same reason)."""

    def __init__(self):
        self.states = {}
        self._retired = []
        self.key = None
        self.cur = None
        self.replaying = False
        self.cursor = 0

    # kept for introspection / tests
    @property
    def entries(self):
        return self.cur.entries if self.cur is not None else []

    @property
    def table(self):
        return self.cur.table if self.cur is not None else None

    def begin(self, key):
        st = self.states.get(key)
        ok = st is not None and st.table is not None and all(e["w"].data_ptr() == e["ptr"] for e in st.entries)
        self.cursor = 0
        self.key = key
        if ok:
            self.cur = st
            self.replaying = True
            if st.n_jobs:
                call("pc_pack_conv_weights_tc_batch", st.table.data_ptr(), st.n_jobs, st.total, stream())
        else:
            if st is not None:
                self._retired.append(st)
            self.replaying = False
            self.cur = self.states[key] = _PackerState()

    def record(self, w, need_dgrad, cw, simt_f, simt_d):
        self.cur.entries.append(dict(w=w, ptr=w.data_ptr(), need_dgrad=need_dgrad, wf=cw.wf, wd=cw.wd, prec_f=cw.prec_f,
                                     prec_d=cw.prec_d, simt_f=simt_f, simt_d=simt_d))

    def end(self):
        if self.replaying:
            return
        st = self.cur
        jobs, total = [], 0
        for e in st.entries:
            O, I, R, S = e["w"].shape
            for dgrad, buf, prec, simt in ((0, e["wf"], e["prec_f"], e["simt_f"]), (1, e["wd"], e["prec_d"], e["simt_d"])):
                if buf is None or simt:
                    continue
                jobs.append(L.PcPackJob(e["ptr"], buf.data_ptr(), O, I, R, S, dgrad, prec, total))
                total += int(L.lib().pc_pack_conv_weight_tc_items(O, I, R, S, dgrad, prec))
        st.n_jobs, st.total = len(jobs), total
        dev = st.entries[0]["w"].device if st.entries else "cpu"
        raw = bytes(b"".join(bytes(j) for j in jobs)) or b"\0"
        st.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)

    def next(self, w, need_dgrad):
        e = self.cur.entries[self.cursor]
        self.cursor += 1
        if e["ptr"] != w.data_ptr() or e["need_dgrad"] != need_dgrad:
            raise RuntimeError("WeightPacker: layer order changed since recording")
        if e["simt_f"] or e["simt_d"]:
            wf, wd = pack_conv_weight(w, e["simt_f"], e["simt_d"])
            if e["simt_f"]:
                e["wf"] = wf
            if e["simt_d"]:
                e["wd"] = wd
        return e["wf"], e["wd"], e["prec_f"], e["prec_d"]


def tc_gemm(a: torch.Tensor, b: torch.Tensor, bias=None, prec=L.PREC_TF32X3) -> torch.Tensor:
    """C = A @ B^T (+ bias) through the tcgen05 tile engine."""
    M, K = a.shape
    N = b.shape[0]
    nbytes = int(L.lib().pc_tc_gemm_workspace(N, K, prec))
    ws = torch.empty(nbytes, device=a.device, dtype=torch.uint8)
    c = torch.empty(M, N, device=a.device, dtype=F32)
    call("pc_tc_gemm", ptr(a), ptr(b), ptr(bias), ptr(c), M, N, K, prec, ptr(ws, torch.uint8), nbytes, stream())
    return c


def conv_fwd(x, w, bias, g: PcConvGeom, xform=None, stats=None, prec=L.PREC_FP32):
    """x NHWC [B,H,W,Cin]; w = packed Wf (Cin>1) or the raw OIHW weight (Cin==1 stem)."""
    y = torch.empty(g.B, g.Ho, g.Wo, g.Cout, device=x.device, dtype=F32)
    xf = _xf(**xform) if xform else None
    L.note_work("pc_conv_fwd", 2.0 * g.B * g.Ho * g.Wo * g.Cout * g.R * g.S * g.Cin)
    call("pc_conv_fwd", ptr(x, None if (xform and xform.get("presplit")) else F32), ptr(w, None), ptr(bias), C.byref(g),
         C.byref(xf) if xf is not None else None, ptr(y),
         ptr(stats, torch.float64), prec, stream())
    return y


def conv_dgrad(dy, wd, g: PcConvGeom, out=None, accumulate=False, prec=L.PREC_FP32, dy_amax=None, dy_presplit=False):
    """dy_amax: 1-element device tensor holding max|dy| (FP16X2 operand scale; see include/phoneme_contrast.h)."""
    if out is None:
        out = torch.empty(g.B, g.H, g.W, g.Cin, device=dy.device, dtype=F32)
    L.note_work("pc_conv_dgrad", 2.0 * g.B * g.Ho * g.Wo * g.Cout * g.R * g.S * g.Cin)
    call("pc_conv_dgrad", ptr(dy, None if dy_presplit else F32), ptr(wd, None), C.byref(g), ptr(out), 1 if accumulate else 0, prec,
         ptr(dy_amax), 1 if dy_presplit else 0, stream())
    return out


def conv_dgrad_bn_reduce(dy_ps, wd, g: PcConvGeom, dy_amax, y, co: "BnCoeffs", drop=None, planes=False, zp=None):
    """conv_dgrad on the halo engine (pre-split dy planes) with the REDUCE pass of the BatchNorm backward that consumes its output
    fused into the epilogue -> (dx, (sums, maxes)) to hand to bn_act_bwd(..., reduced=...), or None when the layer is not covered
    (the caller then runs conv_dgrad + the stand-alone reduce)."""
    if not (g.R == 3 and g.stride == 1) or not L.lib().pc_conv_halo_supported(C.byref(g), 1):
        return None
    C_ = g.Cin
    out = torch.empty(g.B, g.H, g.W, C_, device=y.device, dtype=F32)
    sums = _zeros(zp, (2, C_), torch.float64, y.device)
    maxes = _zeros(zp, (2,), F32, y.device) if planes else None
    red = L.PcBnBwdReduce(ptr(y), ptr(co.scale), ptr(co.shift), ptr(co.mean), ptr(co.invstd), ptr(drop), ptr(sums, torch.float64), ptr(maxes))
    L.note_work("pc_conv_dgrad", 2.0 * g.B * g.Ho * g.Wo * g.Cout * g.R * g.S * g.Cin)
    call("pc_conv_dgrad_halo_bnred", ptr(dy_ps, None), ptr(wd, None), C.byref(g), ptr(out), ptr(dy_amax), C.byref(red), stream())
    return out, (sums, maxes)


_ws_cache: dict = {}


def _workspace(nbytes: int, device, tag: str = "conv") -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream-capture state, user); contents are dead after each call. Users that may
    run on different streams (weight gradients on the side stream, the SupCon kernels on the main one) keep separate buffers."""
    key = (device, torch.cuda.is_current_stream_capturing(), tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), device=device, dtype=torch.uint8)
        _ws_cache[key] = buf
    return buf


def conv_wgrad(x, dy, g: PcConvGeom, xform=None, dw=None, db=None, prec=L.PREC_FP32, dy_amax=None, dy_presplit=False, want_db=True):
    """want_db=False: the bias gradient comes from elsewhere (the BatchNorm backward's closed form, bn_act_bwd(db_conv=...)); the
    column-sum pass over dy is skipped."""
    if dw is None:
        dw = torch.empty(g.Cout, g.Cin, g.R, g.S, device=x.device, dtype=F32)
    if db is None and want_db:
        db = torch.empty(g.Cout, device=x.device, dtype=F32)
    if not want_db:
        db = None
    nbytes = int(L.lib().pc_conv_wgrad_workspace(C.byref(g)))
    ws = _workspace(nbytes, x.device)
    import os
    if os.environ.get("PC_TC_WGRAD", "1") != "1":
        prec = L.PREC_FP32
    xf = _xf(**xform) if xform else None
    L.note_work("pc_conv_wgrad", 2.0 * g.B * g.Ho * g.Wo * g.Cout * g.R * g.S * g.Cin)
    call("pc_conv_wgrad", ptr(x, None if (xform and xform.get("presplit")) else F32), ptr(dy, None if dy_presplit else F32), C.byref(g),
         C.byref(xf) if xf is not None else None, ptr(dw), ptr(db),
         ptr(ws, torch.uint8), ws.numel(), prec, ptr(dy_amax), 1 if dy_presplit else 0, stream())
    return dw, db


class ZeroPool:
    """One zero-filled scratch buffer per backward pass: the per-layer fp64 reduction targets and float max slots are slices of
    it (one fill kernel instead of one per tensor)."""

    def __init__(self, device, nbytes=1 << 17):
        self.buf = torch.zeros(nbytes, device=device, dtype=torch.uint8)
        self.off = 0

    def take(self, shape, dtype):
        n = 1
        for d in shape:
            n *= d
        nb = n * torch.empty((), dtype=dtype).element_size()
        start = (self.off + 15) & ~15
        if start + nb > self.buf.numel():
            return torch.zeros(shape, device=self.buf.device, dtype=dtype)
        self.off = start + nb
        return self.buf[start:start + nb].view(dtype).view(shape)


def _zeros(zp, shape, dtype, device):
    return zp.take(shape, dtype) if zp is not None else torch.zeros(shape, device=device, dtype=dtype)


class BnCoeffs:
    """scale/shift/mean/invstd of one BatchNorm for the current batch (or from running stats in eval)."""
    __slots__ = ("scale", "shift", "mean", "invstd", "C", "gamma", "stats")

    def __init__(self, C_, device):
        buf = torch.empty(4, C_, device=device, dtype=F32)
        self.scale, self.shift, self.mean, self.invstd = buf[0], buf[1], buf[2], buf[3]
        self.C = C_
        self.stats = None


def bn_finalize(stats, count, bn: torch.nn.modules.batchnorm._BatchNorm, training: bool) -> BnCoeffs:
    C_ = bn.num_features
    co = BnCoeffs(C_, bn.weight.device)
    co.gamma = bn.weight
    co.stats = stats if training else None      # fp64 [2, C] sums of y: the closed-form conv bias gradient of the backward reads them
    momentum = 0.1 if bn.momentum is None else bn.momentum
    call("pc_bn_finalize", ptr(stats, torch.float64), C_, float(count), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean),
         ptr(bn.running_var), ptr(bn.num_batches_tracked, torch.int64), momentum, bn.eps, 1 if training else 0,
         ptr(co.scale), ptr(co.shift), ptr(co.mean), ptr(co.invstd), stream())
    return co


def _bn_fin(stats, count, bn, co: BnCoeffs) -> "L.PcBnFinalize":
    momentum = 0.1 if bn.momentum is None else bn.momentum
    return L.PcBnFinalize(ptr(stats, torch.float64), float(count), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var),
                          ptr(bn.num_batches_tracked, torch.int64), momentum, bn.eps, ptr(co.scale), ptr(co.shift), ptr(co.mean), ptr(co.invstd))


def _new_coeffs(stats, bn) -> BnCoeffs:
    co = BnCoeffs(bn.num_features, bn.weight.device)
    co.gamma = bn.weight
    co.stats = stats
    return co


def bn_act_split_fin(y, stats, count, bn, drop=None, relu=True):
    """bn_finalize (train mode) + bn_act_split in ONE launch -> (planes, coefficients)."""
    B, H, W, C_ = y.shape
    co = _new_coeffs(stats, bn)
    fin = _bn_fin(stats, count, bn, co)
    planes = torch.empty(2, y.numel() * 2, device=y.device, dtype=torch.uint8)
    call("pc_bn_act_split_fin", ptr(y), B * H * W, C_, H * W, C.byref(fin), ptr(drop), 1 if relu else 0, ptr(planes, torch.uint8), stream())
    return planes, co


def bn_add_relu_fwd_fin(y2, stats2, count, bn2, ysc, stats_s=None, bn_s=None, want_planes=False):
    """bn_finalize of bn2 (and of the projection shortcut's norm) + bn_add_relu_fwd in ONE launch -> (out[, planes], co2, co_s)."""
    C_ = y2.shape[-1]
    n_pix = y2.numel() // C_
    co2 = _new_coeffs(stats2, bn2)
    fin2 = _bn_fin(stats2, count, bn2, co2)
    co_s = fin_s = None
    if bn_s is not None:
        co_s = _new_coeffs(stats_s, bn_s)
        fin_s = _bn_fin(stats_s, count, bn_s, co_s)
    out = torch.empty_like(y2)
    planes = torch.empty(2, out.numel() * 2, device=y2.device, dtype=torch.uint8) if want_planes else None
    call("pc_bn_add_relu_fwd_fin", ptr(y2), C.byref(fin2), ptr(ysc), C.byref(fin_s) if fin_s is not None else None, n_pix, C_, ptr(out),
         ptr(planes, torch.uint8), stream())
    return ((out, planes) if want_planes else out), co2, co_s


def pool_dims(H, W, pool):
    if pool == 0:
        return H, W
    if pool == 2:
        return H // 2, W // 2
    return (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1


def bn_act_fwd(y, co: BnCoeffs, pool=0, drop=None, want_planes=False):
    """want_planes: also return `out` as fp16 hi | lo planes (see bn_act_split) -> (out, argmax, planes)."""
    B, H, W, C_ = y.shape
    Ho, Wo = pool_dims(H, W, pool)
    out = torch.empty(B, Ho, Wo, C_, device=y.device, dtype=F32)
    argmax = torch.empty(B, Ho, Wo, C_, device=y.device, dtype=torch.uint8) if pool == 3 else None
    planes = torch.empty(2, out.numel() * 2, device=y.device, dtype=torch.uint8) if want_planes else None
    call("pc_bn_act_fwd", ptr(y), B, H, W, C_, ptr(co.scale), ptr(co.shift), ptr(drop), pool, ptr(out),
         ptr(argmax, torch.uint8), ptr(planes, torch.uint8), stream())
    return (out, argmax, planes) if want_planes else (out, argmax)


def bn_act_fwd_fin(y, stats, count, bn, pool=0, drop=None, want_planes=False):
    """bn_finalize (train mode) + bn_act_fwd in ONE launch -> ((out, argmax[, planes]), coefficients)."""
    B, H, W, C_ = y.shape
    Ho, Wo = pool_dims(H, W, pool)
    co = _new_coeffs(stats, bn)
    fin = _bn_fin(stats, count, bn, co)
    out = torch.empty(B, Ho, Wo, C_, device=y.device, dtype=F32)
    argmax = torch.empty(B, Ho, Wo, C_, device=y.device, dtype=torch.uint8) if pool == 3 else None
    planes = torch.empty(2, out.numel() * 2, device=y.device, dtype=torch.uint8) if want_planes else None
    call("pc_bn_act_fwd_fin", ptr(y), B, H, W, C_, C.byref(fin), ptr(drop), pool, ptr(out), ptr(argmax, torch.uint8), ptr(planes, torch.uint8),
         stream())
    return ((out, argmax, planes) if want_planes else (out, argmax)), co


def bn_act_bwd(dout, y, co: BnCoeffs, pool=0, drop=None, argmax=None, dgamma=None, dbeta=None, amax=None, planes=False, zp=None,
               db_conv=None, sync=None, reduced=None):
    """Gradient w.r.t. the pre-BatchNorm tensor y of out = drop * pool(relu(bn(y))) (train-mode statistics).
    amax: optional zero-initialised 1-element tensor that receives max|dy| (operand scale of the FP16X2 convolutions).
    planes=True: dy is returned ONLY as fp16 hi | lo planes (uint8 [2, numel*2]) scaled by the power of two derived from a bound
    of |dy| that is stored in `amax` (for convolutions called with dy_presplit=True)."""
    B, H, W, C_ = y.shape
    args = (ptr(dout), ptr(y), B, H, W, C_, ptr(co.scale), ptr(co.shift), ptr(co.mean), ptr(co.invstd), ptr(drop), pool,
            ptr(argmax, torch.uint8))
    if reduced is not None:
        sums, maxes = reduced          # produced by the data gradient that wrote dout (conv_dgrad_bn_reduce): no reduce pass here
    else:
        sums = _zeros(zp, (2, C_), torch.float64, y.device)
        maxes = _zeros(zp, (2,), F32, y.device) if planes else None
        call("pc_bn_act_bwd_reduce", *args, ptr(sums, torch.float64), ptr(maxes), stream())
    if sync is not None:
        # synchronised BatchNorm: the apply pass takes the per-rank AVERAGE of the global (sum dz, sum dz xhat); with its local pixel
        # count the projection terms are then those of the global batch, and dgamma / dbeta are this rank's share of the all-reduced sum
        sync.sync([sums], 1.0 / sync.R)
    dy = None if planes else torch.empty_like(y)
    dy_ps = torch.empty(2, y.numel() * 2, device=y.device, dtype=torch.uint8) if planes else None
    if dgamma is None:
        dgamma = torch.empty(C_, device=y.device, dtype=F32)
    if dbeta is None:
        dbeta = torch.empty(C_, device=y.device, dtype=F32)
    # db_conv: gradient of the bias of the convolution that produced y, in closed form (no pass over dy; csrc/bn_act.cu)
    st = getattr(co, "stats", None) if db_conv is not None else None
    if db_conv is not None and st is None:
        raise ValueError("bn_act_bwd: db_conv needs coefficients finalised from train-mode statistics")
    call("pc_bn_act_bwd_apply", *args, ptr(sums, torch.float64), ptr(dy), ptr(dgamma), ptr(dbeta), ptr(amax), ptr(maxes),
         ptr(dy_ps, torch.uint8), ptr(st, torch.float64), ptr(db_conv), stream())
    return (dy_ps if planes else dy), dgamma, dbeta


def stem_gram(x):
    """G [49,49] and X1 [49] (fp64) of the 7x7 patches of x [B,1,H,W]: the data-only part of the stem's weight gradient."""
    B, _, H, W = x.shape
    buf = torch.zeros(49 * 49 + 49, device=x.device, dtype=torch.float64)
    call("pc_stem_gram", ptr(x), B, H, W, ptr(buf, torch.float64), buf[49 * 49:].data_ptr(), stream())
    return buf


def stem_stats_from_gram(gram, conv, B, H, W, stats, bn=None):
    """stats [2,64] fp64 = (sum y0, sum y0^2) of y0 = conv(x) + b from the Gram matrix / tap sums of x's patches (closed form).
    bn: also finalise that BatchNorm (train mode) in the same launch and return its coefficients."""
    co = fin = None
    if bn is not None:
        co = _new_coeffs(stats, bn)
        fin = _bn_fin(stats, B * H * W, bn, co)
    call("pc_stem_stats_from_gram", ptr(gram, torch.float64), gram[49 * 49:].data_ptr(), ptr(conv.weight), ptr(conv.bias), B, H, W,
         ptr(stats, torch.float64), C.byref(fin) if fin is not None else None, stream())
    return co


def stem_fwd(x, conv, co: BnCoeffs, want_planes=False):
    """maxpool3x3s2p1(relu(bn(conv7x7(x)))) in one kernel -> (p0 [B,Hp,Wp,64], argmax uint8, planes or None)."""
    B, _, H, W = x.shape
    Hp, Wp = pool_dims(H, W, 3)
    p0 = torch.empty(B, Hp, Wp, 64, device=x.device, dtype=F32)
    argmax = torch.empty(B, Hp, Wp, 64, device=x.device, dtype=torch.uint8)
    planes = torch.empty(2, p0.numel() * 2, device=x.device, dtype=torch.uint8) if want_planes else None
    L.note_work("pc_stem_fwd", 2.0 * B * H * W * 64 * 49)
    call("pc_stem_fwd", ptr(x), ptr(conv.weight), ptr(conv.bias), ptr(co.scale), ptr(co.shift), B, H, W, ptr(p0), ptr(argmax, torch.uint8),
         ptr(planes, torch.uint8), stream())
    return p0, argmax, planes


def stem_bwd(dpool, p0, argmax, x, conv, co: BnCoeffs, gram, dw, db, dgamma, dbeta, zp=None):
    """Backward of Conv(1,64,7) -> BN -> ReLU -> MaxPool(3,2,1) from the pooled-resolution tensors (csrc/stem_bwd.cu)."""
    B, _, H, W = x.shape
    sums = _zeros(zp, (2, 64), torch.float64, x.device)
    amax = _zeros(zp, (1,), F32, x.device)
    nbytes = int(L.lib().pc_stem_bwd_workspace())
    ws = _workspace(nbytes, x.device, "stem")
    call("pc_stem_bwd", ptr(dpool), ptr(p0), ptr(argmax, torch.uint8), ptr(x), B, H, W, ptr(conv.weight), ptr(conv.bias), ptr(co.gamma),
         ptr(co.scale), ptr(co.shift), ptr(co.mean), ptr(co.invstd), ptr(gram, torch.float64), gram[49 * 49:].data_ptr(),
         ptr(sums, torch.float64), ptr(amax), ptr(ws, torch.uint8), ws.numel(), ptr(dw), ptr(db), ptr(dgamma), ptr(dbeta), stream())


def bn_add_relu_fwd(y2, co2: BnCoeffs, ysc, co_s: BnCoeffs | None, want_planes=False):
    C_ = y2.shape[-1]
    n_pix = y2.numel() // C_
    out = torch.empty_like(y2)
    planes = torch.empty(2, out.numel() * 2, device=y2.device, dtype=torch.uint8) if want_planes else None
    call("pc_bn_add_relu_fwd", ptr(y2), ptr(co2.scale), ptr(co2.shift), ptr(ysc), ptr(co_s.scale) if co_s else None,
         ptr(co_s.shift) if co_s else None, n_pix, C_, ptr(out), ptr(planes, torch.uint8), stream())
    return (out, planes) if want_planes else out


def bn_add_relu_bwd(dout, out, y2, co2: BnCoeffs, ysc, co_s: BnCoeffs | None, grads2=None, grads_s=None, amax2=None,
                    amax_s=None, planes=False, zp=None, db2=None, db_s=None, sync=None):
    """Returns dy2, d(shortcut branch input: dysc for a projection shortcut, dx for identity), (dgamma2, dbeta2), (dgamma_s, dbeta_s).
    planes=True: dy2 -- and dysc of a projection shortcut -- come back ONLY as scaled fp16 hi | lo planes (see bn_act_bwd); the
    identity-shortcut dx stays fp32 (it is accumulated into, not convolved)."""
    C_ = y2.shape[-1]
    n_pix = y2.numel() // C_
    dev = y2.device
    sums2 = _zeros(zp, (2, C_), torch.float64, dev)
    sums_s = _zeros(zp, (2, C_), torch.float64, dev) if co_s else None
    maxes = _zeros(zp, (3,), F32, dev) if planes else None
    call("pc_bn_add_relu_bwd_reduce", ptr(dout), ptr(out), ptr(y2), ptr(co2.mean), ptr(co2.invstd), ptr(ysc) if co_s else None,
         ptr(co_s.mean) if co_s else None, ptr(co_s.invstd) if co_s else None, n_pix, C_, ptr(sums2, torch.float64),
         ptr(sums_s, torch.float64), ptr(maxes), stream())
    if sync is not None:      # see bn_act_bwd
        sync.sync([sums2] + ([sums_s] if co_s else []), 1.0 / sync.R)
    ps_sc = planes and co_s is not None
    dy2 = None if planes else torch.empty_like(y2)
    dsc = None if ps_sc else torch.empty_like(y2)
    dy2_ps = torch.empty(2, y2.numel() * 2, device=dev, dtype=torch.uint8) if planes else None
    dsc_ps = torch.empty(2, y2.numel() * 2, device=dev, dtype=torch.uint8) if ps_sc else None
    g2 = grads2 or (torch.empty(C_, device=dev, dtype=F32), torch.empty(C_, device=dev, dtype=F32))
    gs = grads_s or ((torch.empty(C_, device=dev, dtype=F32), torch.empty(C_, device=dev, dtype=F32)) if co_s else (None, None))
    call("pc_bn_add_relu_bwd_apply", ptr(dout), ptr(out), ptr(y2), ptr(co2.scale), ptr(co2.mean), ptr(co2.invstd),
         ptr(sums2, torch.float64), ptr(ysc) if co_s else None, ptr(co_s.scale) if co_s else None,
         ptr(co_s.mean) if co_s else None, ptr(co_s.invstd) if co_s else None, ptr(sums_s, torch.float64), n_pix, C_,
         ptr(dy2), ptr(dsc), ptr(g2[0]), ptr(g2[1]), ptr(gs[0]), ptr(gs[1]), ptr(amax2), ptr(amax_s if (ps_sc or not planes) else None),
         ptr(maxes), ptr(dy2_ps, torch.uint8), ptr(dsc_ps, torch.uint8),
         ptr(getattr(co2, "stats", None) if db2 is not None else None, torch.float64), ptr(db2),
         ptr(getattr(co_s, "stats", None) if (db_s is not None and co_s) else None, torch.float64), ptr(db_s if co_s else None), stream())
    return (dy2_ps if planes else dy2), (dsc_ps if ps_sc else dsc), g2, gs


_attn_ws: dict = {}


def attn_pool_fwd(a, w=None, b0=None):
    B, H, W, C_ = a.shape
    gate = torch.empty(B, H * W, device=a.device, dtype=F32)
    pooled = torch.empty(B, C_, device=a.device, dtype=F32)
    S = int(L.lib().pc_attn_pool_splits(B))
    if S > 1:
        # few samples: S blocks share a sample's pixels; persistent scratch + zeroed arrival counters (the kernel leaves them zero)
        key = (a.device, B, S, C_)
        ws = _attn_ws.get(key)
        if ws is None:
            ws = _attn_ws[key] = (torch.empty(B * S * C_, device=a.device, dtype=F32), torch.zeros(B, device=a.device, dtype=torch.int32))
        call("pc_attn_pool_fwd_ws", ptr(a), B, H * W, C_, ptr(w), ptr(b0), ptr(gate), ptr(pooled), ptr(ws[0]), ptr(ws[1], torch.int32), stream())
    else:
        call("pc_attn_pool_fwd", ptr(a), B, H * W, C_, ptr(w), ptr(b0), ptr(gate), ptr(pooled), stream())
    return pooled, gate


def attn_pool_bwd(a, gate, dpooled, w=None, dw=None, db0=None):
    B, H, W, C_ = a.shape
    da = torch.empty_like(a)
    if w is not None:
        dw = torch.empty(C_, device=a.device, dtype=F32) if dw is None else dw
        db0 = torch.empty(1, device=a.device, dtype=F32) if db0 is None else db0
    call("pc_attn_pool_bwd", ptr(a), ptr(gate), ptr(dpooled), B, H * W, C_, ptr(w), ptr(da), ptr(dw), ptr(db0), stream())
    return da, dw, db0


def head_fwd(x, lin: torch.nn.Linear, bn: torch.nn.BatchNorm1d, training: bool, sync=None):
    B, K = x.shape
    N = lin.out_features
    nbytes = int(L.lib().pc_head_workspace(B, K, N))
    ws = torch.empty(nbytes // 4, device=x.device, dtype=F32)
    emb = torch.empty(B, N, device=x.device, dtype=F32)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    if sync is not None and training:
        # BatchNorm1d statistics over the GLOBAL batch: Linear + local sums | exchange | coefficients from the global sums + normalise
        sums = torch.empty(2, N, device=x.device, dtype=torch.float64)
        args = (ptr(x), B, K, N, ptr(lin.weight), ptr(lin.bias), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var),
                ptr(bn.num_batches_tracked, torch.int64), momentum, bn.eps, ptr(emb), ptr(ws), ptr(sums, torch.float64), float(B * sync.R))
        call("pc_head_fwd_sync", *args, 1, stream())
        sync.sync([sums], 1.0)
        call("pc_head_fwd_sync", *args, 2, stream())
        return emb, ws
    call("pc_head_fwd", ptr(x), B, K, N, ptr(lin.weight), ptr(lin.bias), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean),
         ptr(bn.running_var), ptr(bn.num_batches_tracked, torch.int64), momentum, bn.eps, 1 if training else 0, ptr(emb),
         ptr(ws), stream())
    return emb, ws


def head_bwd(demb, x, lin_w, bn_w, bn_b, training, ws, dW=None, dbias=None, dgamma=None, dbeta=None, sync=None):
    B, K = x.shape
    N = lin_w.shape[0]
    dev = x.device
    dx = torch.empty(B, K, device=dev, dtype=F32)
    dW = torch.empty(N, K, device=dev, dtype=F32) if dW is None else dW
    dbias = torch.empty(N, device=dev, dtype=F32) if dbias is None else dbias
    dgamma = torch.empty(N, device=dev, dtype=F32) if dgamma is None else dgamma
    dbeta = torch.empty(N, device=dev, dtype=F32) if dbeta is None else dbeta
    if sync is not None and training:
        sums = torch.empty(2, N, device=dev, dtype=torch.float64)
        args = (ptr(demb), ptr(x), B, K, N, ptr(lin_w), ptr(bn_w), ptr(bn_b), ptr(ws), ptr(dx), ptr(dW), ptr(dbias), ptr(dgamma), ptr(dbeta),
                ptr(sums, torch.float64))
        call("pc_head_bwd_sync", *args, 1, stream())
        sync.sync([sums], 1.0 / sync.R)
        call("pc_head_bwd_sync", *args, 2, stream())
        return dx, dW, dbias, dgamma, dbeta
    call("pc_head_bwd", ptr(demb), ptr(x), B, K, N, ptr(lin_w), ptr(bn_w), ptr(bn_b), 1 if training else 0, ptr(ws), ptr(dx),
         ptr(dW), ptr(dbias), ptr(dgamma), ptr(dbeta), stream())
    return dx, dW, dbias, dgamma, dbeta


def dropout2d_mask(B, C_, p, seed, offset, device, step_dev=None):
    m = torch.empty(B, C_, device=device, dtype=F32)
    call("pc_dropout2d_mask", ptr(m), B, C_, float(p), int(seed), int(offset), ptr(step_dev, torch.int64), stream())
    return m


def counter_add(counter: torch.Tensor, inc: int = 1):
    call("pc_counter_add", ptr(counter, torch.int64), int(inc), stream())


def supcon_fwd(feats, labels, mask, temperature, base_temperature, row0=0, nrows=None):
    N, D = feats.shape
    nrows = N - row0 if nrows is None else nrows
    stats = torch.empty(nrows, 4, device=feats.device, dtype=F32)
    row_loss = torch.empty(nrows, device=feats.device, dtype=F32)
    import os
    # label-form problems from N = 256 up: similarity tiles on the tensor cores (csrc/supcon_tc.cu; at N = 256 it is 0.06 ms vs
    # 0.12 ms fwd+bwd for the SIMT kernels); PC_SUPCON_TC=0 / 1 forces the choice
    tc_env = os.environ.get("PC_SUPCON_TC")
    if (mask is None and labels is not None and tc_env != "0" and (N >= 256 or tc_env == "1")
            and L.lib().pc_supcon_tc_supported(N, D, row0, nrows)):
        nbytes = int(L.lib().pc_supcon_tc_workspace(N, D, nrows))
        ws = _workspace(nbytes, feats.device, "supcon")
        call("pc_supcon_fwd_tc", ptr(feats), ptr(labels, torch.int64), N, D, row0, nrows, float(temperature), float(base_temperature),
             ptr(ws, torch.uint8), ws.numel(), ptr(stats), ptr(row_loss), stream())
        return stats, row_loss
    call("pc_supcon_fwd", ptr(feats), ptr(labels, torch.int64), ptr(mask), N, D, row0, nrows, float(temperature),
         float(base_temperature), ptr(stats), ptr(row_loss), stream())
    return stats, row_loss


def dp_pack(emb, labels):
    """[n,D] embeddings + [n] int64 labels -> [n, D+2] fp32 rows (label bits in the last two columns): one all_gather carries both."""
    n, D = emb.shape
    out = torch.empty(n, D + 2, device=emb.device, dtype=F32)
    call("pc_dp_pack", ptr(emb), ptr(labels, torch.int64), n, D, ptr(out), stream())
    return out


def dp_unpack(packed, D, F=None, y=None):
    N = packed.shape[0]
    F = torch.empty(N, D, device=packed.device, dtype=F32) if F is None else F
    y = torch.empty(N, device=packed.device, dtype=torch.int64) if y is None else y
    call("pc_dp_unpack", ptr(packed), N, D, ptr(F), ptr(y, torch.int64), stream())
    return F, y


def supcon_loss_from_stats(stats_all, temperature, base_temperature, scale, out=None):
    out = torch.empty(1, device=stats_all.device, dtype=F32) if out is None else out
    call("pc_supcon_loss_from_stats", ptr(stats_all), stats_all.shape[0], float(temperature), float(base_temperature), float(scale), ptr(out), stream())
    return out


def sum_scaled(x, scale):
    out = torch.empty((), device=x.device, dtype=F32)
    call("pc_sum_scaled", ptr(x), x.numel(), float(scale), ptr(out), stream())
    return out


def supcon_bwd(feats, labels, mask, temperature, coef, grad_scale, stats_all, row0=0, nrows=None):
    N, D = feats.shape
    nrows = N - row0 if nrows is None else nrows
    dF = torch.empty(nrows, D, device=feats.device, dtype=F32)
    import os
    tc_env = os.environ.get("PC_SUPCON_TC")
    if (mask is None and labels is not None and tc_env != "0" and (N >= 256 or tc_env == "1")
            and L.lib().pc_supcon_tc_supported(N, D, row0, nrows)):
        nbytes = int(L.lib().pc_supcon_bwd_tc_workspace(N, D, nrows))
        ws = _workspace(nbytes, feats.device, "supcon")
        call("pc_supcon_bwd_tc", ptr(feats), ptr(labels, torch.int64), N, D, row0, nrows, float(temperature), float(coef),
             ptr(grad_scale), ptr(stats_all), ptr(ws, torch.uint8), ws.numel(), ptr(dF), stream())
        return dF
    call("pc_supcon_bwd", ptr(feats), ptr(labels, torch.int64), ptr(mask), N, D, row0, nrows, float(temperature), float(coef),
         ptr(grad_scale), ptr(stats_all), ptr(dF), stream())
    return dF
