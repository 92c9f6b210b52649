"""Halo-resident convolution engine (csrc/conv_halo.cu) against an fp64 reference convolution and against the round-1
per-tap-gather kernel on the same operand bytes: forward (+ bias, BatchNorm sums) and data gradient, every cluster size,
image shapes of all four cnn_deep stages, batch sizes that leave partial / empty tiles in a cluster."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [  # B, H, W, Cin, Cout
    (3, 20, 51, 64, 64),
    (5, 10, 26, 128, 128),
    (7, 5, 13, 256, 256),
    (9, 3, 7, 512, 512),
    (2, 5, 13, 64, 192),     # Cout not a multiple of the 128 tile
    (40, 20, 51, 64, 64),    # several tiles per CTA
    (1, 3, 7, 128, 64),      # a single, mostly empty tile
]


def _ref_conv(x, w, bias):
    return torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), bias.double() if bias is not None else None, padding=1).permute(0, 2, 3, 1)


@pytest.mark.parametrize("cluster", [1, 2, 4, 0])     # 0: streaming kernel even where the weight-resident variant applies
@pytest.mark.parametrize("case", CASES)
def test_halo_forward(case, cluster, monkeypatch):
    from phoneme_contrast_b200 import _lib as L
    from phoneme_contrast_b200 import ops
    import ctypes as C
    B, H, W, Cin, Cout = case
    g = ops.conv_geom(B, H, W, Cin, Cout, 3, 1, 1)
    gen = torch.Generator(device=DEV).manual_seed(B * 1000 + Cin + Cout)
    x = torch.randn(B, H, W, Cin, device=DEV, generator=gen)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=gen) * (2.0 / (Cin * 9)) ** 0.5
    bias = torch.randn(Cout, device=DEV, generator=gen)
    cw = ops.ConvWeights(w, g, L.PREC_FP16X2)
    planes = ops.bn_act_split(x)
    monkeypatch.setenv("PC_HALO_CLUSTER", str(max(cluster, 1)))
    monkeypatch.setenv("PC_HALO_RESIDENT", "0" if cluster == 0 else "1")
    monkeypatch.setenv("PC_HALO_ALL", "1")       # also the 256-channel layers, which the default routing leaves to the per-tap kernel
    assert L.lib().pc_conv_halo_supported(C.byref(g), 0) == 1
    st = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    y = ops.conv_fwd(planes, cw.wf, bias, g, dict(presplit=True), st, cw.prec_f)
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, bias)
    scale = float(ref.abs().max())
    err = float((y.double() - ref).abs().max())
    assert err <= 5e-6 * scale, (err, scale)
    np.testing.assert_allclose(st[0].cpu().numpy(), ref.sum(dim=(0, 1, 2)).cpu().numpy(), rtol=2e-5, atol=2e-4 * scale)
    np.testing.assert_allclose(st[1].cpu().numpy(), (ref ** 2).sum(dim=(0, 1, 2)).cpu().numpy(), rtol=2e-5)
    # the round-1 kernel on the same operand bytes
    monkeypatch.setenv("PC_CONV_HALO", "0")
    assert L.lib().pc_conv_halo_supported(C.byref(g), 0) == 0
    y0 = ops.conv_fwd(planes, cw.wf, bias, g, dict(presplit=True), None, cw.prec_f)
    assert float((y - y0).abs().max()) <= 2e-6 * scale


@pytest.mark.parametrize("cluster", [1, 4])
@pytest.mark.parametrize("case", CASES[:4] + CASES[5:])
def test_halo_dgrad(case, cluster, monkeypatch):
    from phoneme_contrast_b200 import _lib as L
    from phoneme_contrast_b200 import ops
    import ctypes as C
    B, H, W, Cin, Cout = case
    g = ops.conv_geom(B, H, W, Cin, Cout, 3, 1, 1)
    gen = torch.Generator(device=DEV).manual_seed(B * 77 + Cin + 3 * Cout)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=gen) * (2.0 / (Cin * 9)) ** 0.5
    yconv = torch.randn(B, H, W, Cout, device=DEV, generator=gen)
    dout = torch.randn(B, H, W, Cout, device=DEV, generator=gen) * 1e-6
    bn = torch.nn.BatchNorm2d(Cout).to(DEV)
    st = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    st[0] = yconv.double().sum((0, 1, 2)); st[1] = (yconv.double() ** 2).sum((0, 1, 2))
    co = ops.bn_finalize(st, B * H * W, bn, True)
    cw = ops.ConvWeights(w, g, L.PREC_FP16X2)
    a0, a1 = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    dy, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a0)
    dy_ps, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a1, planes=True)
    monkeypatch.setenv("PC_HALO_CLUSTER", str(cluster))
    monkeypatch.setenv("PC_HALO_ALL", "1")
    assert L.lib().pc_conv_halo_supported(C.byref(g), 1) == 1
    dx = ops.conv_dgrad(dy_ps, cw.wd, g, prec=cw.prec_d, dy_amax=a1, dy_presplit=True)
    ref = torch.nn.functional.conv_transpose2d(dy.permute(0, 3, 1, 2).double(), w.double(), padding=1).permute(0, 2, 3, 1)
    scale = float(ref.abs().max())
    assert float((dx.double() - ref).abs().max()) <= 1e-5 * scale
    # accumulate into an existing tensor
    base = torch.randn(B, H, W, Cin, device=DEV, generator=gen) * scale
    acc = base.clone()
    ops.conv_dgrad(dy_ps, cw.wd, g, out=acc, accumulate=True, prec=cw.prec_d, dy_amax=a1, dy_presplit=True)
    assert float((acc.double() - (base.double() + ref)).abs().max()) <= 1e-5 * scale + 1e-6 * float(base.abs().max())


CASES_1X1 = [  # B, H, W, Cin, Cout, stride  (the three cnn_deep projection shortcuts + a stride-1 case + ragged sizes)
    (4, 20, 51, 64, 128, 2),
    (6, 10, 26, 128, 256, 2),
    (9, 5, 13, 256, 512, 2),
    (3, 10, 26, 64, 64, 1),
    (2, 7, 9, 64, 128, 2),
    (33, 20, 51, 64, 128, 2),
]


@pytest.mark.parametrize("case", CASES_1X1)
def test_halo_1x1_forward_and_dgrad(case, monkeypatch):
    """1x1 (shortcut) convolutions on the halo engine: strided TMA boxes sample every second pixel in the forward pass, the data
    gradient scatters its rows to every second pixel of an existing tensor. Against fp64 and against the per-tap-gather kernel."""
    from phoneme_contrast_b200 import _lib as L
    from phoneme_contrast_b200 import ops
    import ctypes as C
    B, H, W, Cin, Cout, stride = case
    g = ops.conv_geom(B, H, W, Cin, Cout, 1, stride, 0)
    gen = torch.Generator(device=DEV).manual_seed(B * 31 + Cin + Cout + stride)
    x = torch.randn(B, H, W, Cin, device=DEV, generator=gen)
    w = torch.randn(Cout, Cin, 1, 1, device=DEV, generator=gen) * (2.0 / Cin) ** 0.5
    bias = torch.randn(Cout, device=DEV, generator=gen)
    cw = ops.ConvWeights(w, g, L.PREC_FP16X2)
    planes = ops.bn_act_split(x)
    assert L.lib().pc_conv_halo_supported(C.byref(g), 0) == 1
    st = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    y = ops.conv_fwd(planes, cw.wf, bias, g, dict(presplit=True), st, cw.prec_f)
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), bias.double(), stride=stride).permute(0, 2, 3, 1)
    assert tuple(y.shape) == tuple(ref.shape)
    scale = float(ref.abs().max())
    assert float((y.double() - ref).abs().max()) <= 5e-6 * scale
    np.testing.assert_allclose(st[0].cpu().numpy(), ref.sum(dim=(0, 1, 2)).cpu().numpy(), rtol=2e-5, atol=2e-4 * scale)
    np.testing.assert_allclose(st[1].cpu().numpy(), (ref ** 2).sum(dim=(0, 1, 2)).cpu().numpy(), rtol=2e-5)
    monkeypatch.setenv("PC_HALO_1X1", "0")
    assert L.lib().pc_conv_halo_supported(C.byref(g), 0) == 0
    y0 = ops.conv_fwd(planes, cw.wf, bias, g, dict(presplit=True), None, cw.prec_f)
    assert float((y - y0).abs().max()) <= 2e-6 * scale
    monkeypatch.setenv("PC_HALO_1X1", "1")

    # data gradient: accumulate into an existing dx
    Ho, Wo = g.Ho, g.Wo
    yconv = torch.randn(B, Ho, Wo, Cout, device=DEV, generator=gen)
    dout = torch.randn(B, Ho, Wo, Cout, device=DEV, generator=gen) * 1e-6
    bn = torch.nn.BatchNorm2d(Cout).to(DEV)
    st2 = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    st2[0] = yconv.double().sum((0, 1, 2)); st2[1] = (yconv.double() ** 2).sum((0, 1, 2))
    co = ops.bn_finalize(st2, B * Ho * Wo, bn, True)
    a0, a1 = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    dy, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a0)
    dy_ps, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a1, planes=True)
    refd = torch.nn.functional.conv_transpose2d(dy.permute(0, 3, 1, 2).double(), w.double(), stride=stride,
                                                output_padding=(H - ((Ho - 1) * stride + 1), W - ((Wo - 1) * stride + 1))).permute(0, 2, 3, 1)
    dscale = float(refd.abs().max())
    base = torch.randn(B, H, W, Cin, device=DEV, generator=gen) * dscale
    acc = base.clone()
    ops.conv_dgrad(dy_ps, cw.wd, g, out=acc, accumulate=True, prec=cw.prec_d, dy_amax=a1, dy_presplit=True)
    assert float((acc.double() - (base.double() + refd)).abs().max()) <= 1e-5 * dscale + 1e-6 * float(base.abs().max())
    # without accumulation a strided 1x1 gradient must still define every pixel (the per-tap-gather kernel handles it)
    dx = ops.conv_dgrad(dy_ps, cw.wd, g, prec=cw.prec_d, dy_amax=a1, dy_presplit=True)
    assert float((dx.double() - refd).abs().max()) <= 1e-5 * dscale


WG_CASES = [  # B, H, W, Cin, Cout  (the stride-1 3x3 layers of cnn_deep + ragged batches / channel mixes)
    (6, 20, 51, 64, 64),
    (37, 20, 51, 64, 64),
    (8, 10, 26, 128, 128),
    (8, 5, 13, 256, 256),
    (8, 3, 7, 512, 512),
    (5, 5, 13, 64, 128),
    (3, 20, 51, 128, 64),
    (16, 3, 7, 128, 256),
]


@pytest.mark.parametrize("case", WG_CASES)
def test_halo_wgrad(case, monkeypatch):
    """Halo weight-gradient engine (csrc/conv_halo_wgrad.cu) against fp64 and against the per-tap-gather kernel on the same planes."""
    from phoneme_contrast_b200 import _lib as L
    from phoneme_contrast_b200 import ops
    import ctypes as C
    B, H, W, Cin, Cout = case
    g = ops.conv_geom(B, H, W, Cin, Cout, 3, 1, 1)
    gen = torch.Generator(device=DEV).manual_seed(B * 13 + Cin + 7 * Cout + H)
    x = torch.relu(torch.randn(B, H, W, Cin, device=DEV, generator=gen))
    yconv = torch.randn(B, H, W, Cout, device=DEV, generator=gen)
    dout = torch.randn(B, H, W, Cout, device=DEV, generator=gen) * 1e-6
    bn = torch.nn.BatchNorm2d(Cout).to(DEV)
    st = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    st[0] = yconv.double().sum((0, 1, 2)); st[1] = (yconv.double() ** 2).sum((0, 1, 2))
    co = ops.bn_finalize(st, B * H * W, bn, True)
    a0, a1 = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    dy, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a0)
    dy_ps, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a1, planes=True)
    x_ps = ops.bn_act_split(x)
    assert L.lib().pc_conv_wgrad_halo_supported(C.byref(g)) == 1
    dw, _ = ops.conv_wgrad(x_ps, dy_ps, g, dict(presplit=True), prec=L.PREC_FP16X2, dy_amax=a1, dy_presplit=True, want_db=False)
    torch.cuda.synchronize()
    # fp64 reference: dW[o][c][r][s] = sum dy[b,h,w,o] x[b,h+r-1,w+s-1,c]
    xp = torch.nn.functional.pad(x.double().permute(0, 3, 1, 2), (1, 1, 1, 1))
    dyd = dy.double().permute(0, 3, 1, 2)
    ref = torch.empty(Cout, Cin, 3, 3, device=DEV, dtype=torch.float64)
    for r in range(3):
        for s in range(3):
            ref[:, :, r, s] = torch.einsum("bohw,bchw->oc", dyd, xp[:, :, r:r + H, s:s + W])
    scale = float(ref.abs().max())
    err = float((dw.double() - ref).abs().max())
    assert err <= 2e-5 * scale, (err, scale)
    monkeypatch.setenv("PC_WGRAD_HALO", "0")
    assert L.lib().pc_conv_wgrad_halo_supported(C.byref(g)) == 0
    dw0, _ = ops.conv_wgrad(x_ps, dy_ps, g, dict(presplit=True), prec=L.PREC_FP16X2, dy_amax=a1, dy_presplit=True, want_db=False)
    assert float((dw - dw0).abs().max()) <= 1e-5 * scale


S2_CASES = [  # B, H, W, Cin, Cout  (cnn_deep's stride-2 3x3 layers: data gradient as four parity-class problems over dy)
    (4, 20, 51, 64, 128),
    (7, 10, 26, 128, 256),
    (33, 20, 51, 64, 128),
    (3, 9, 12, 64, 64),
]


@pytest.mark.parametrize("accumulate", [False, True])
@pytest.mark.parametrize("case", S2_CASES)
def test_halo_stride2_dgrad(case, accumulate, monkeypatch):
    from phoneme_contrast_b200 import _lib as L
    from phoneme_contrast_b200 import ops
    import ctypes as C
    B, H, W, Cin, Cout = case
    monkeypatch.setenv("PC_HALO_S2", "1")       # off by default: measured slower than the per-tap-gather kernel (csrc/conv_halo.cu)
    g = ops.conv_geom(B, H, W, Cin, Cout, 3, 2, 1)
    Ho, Wo = g.Ho, g.Wo
    gen = torch.Generator(device=DEV).manual_seed(B * 5 + Cin + 11 * Cout + W)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=gen) * (2.0 / (Cin * 9)) ** 0.5
    yconv = torch.randn(B, Ho, Wo, Cout, device=DEV, generator=gen)
    dout = torch.randn(B, Ho, Wo, Cout, device=DEV, generator=gen) * 1e-6
    bn = torch.nn.BatchNorm2d(Cout).to(DEV)
    st = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    st[0] = yconv.double().sum((0, 1, 2)); st[1] = (yconv.double() ** 2).sum((0, 1, 2))
    co = ops.bn_finalize(st, B * Ho * Wo, bn, True)
    cw = ops.ConvWeights(w, g, L.PREC_FP16X2)
    a0, a1 = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    dy, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a0)
    dy_ps, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a1, planes=True)
    assert L.lib().pc_conv_halo_supported(C.byref(g), 1) == 1
    ref = torch.nn.functional.conv_transpose2d(dy.permute(0, 3, 1, 2).double(), w.double(), stride=2, padding=1,
                                               output_padding=(H - ((Ho - 1) * 2 - 2 + 3), W - ((Wo - 1) * 2 - 2 + 3))).permute(0, 2, 3, 1)
    assert tuple(ref.shape) == (B, H, W, Cin)
    scale = float(ref.abs().max())
    if accumulate:
        base = torch.randn(B, H, W, Cin, device=DEV, generator=gen) * scale
        dx = base.clone()
        ops.conv_dgrad(dy_ps, cw.wd, g, out=dx, accumulate=True, prec=cw.prec_d, dy_amax=a1, dy_presplit=True)
        assert float((dx.double() - (base.double() + ref)).abs().max()) <= 1e-5 * scale + 1e-6 * float(base.abs().max())
    else:
        dx = ops.conv_dgrad(dy_ps, cw.wd, g, prec=cw.prec_d, dy_amax=a1, dy_presplit=True)
        assert float((dx.double() - ref).abs().max()) <= 1e-5 * scale
        monkeypatch.setenv("PC_HALO_S2", "0")
        assert L.lib().pc_conv_halo_supported(C.byref(g), 1) == 0
        dx0 = ops.conv_dgrad(dy_ps, cw.wd, g, prec=cw.prec_d, dy_amax=a1, dy_presplit=True)
        assert float((dx - dx0).abs().max()) <= 2e-6 * scale


@pytest.mark.parametrize("drop_on", [False, True])
@pytest.mark.parametrize("case", CASES[:4] + CASES[5:])
def test_halo_dgrad_fused_bn_reduce(case, drop_on, monkeypatch):
    """The data gradient whose epilogue also runs the REDUCE pass of the BatchNorm backward below it (Params::red): dx bit-identical
    to the plain data gradient, (sum dz, sum dz xhat) and (max |dz|, max |xhat|) equal to pc_bn_act_bwd_reduce on that dx."""
    from phoneme_contrast_b200 import _lib as L
    from phoneme_contrast_b200 import ops
    from phoneme_contrast_b200._lib import call, ptr, stream
    B, H, W, Cin, Cout = case
    g = ops.conv_geom(B, H, W, Cin, Cout, 3, 1, 1)
    gen = torch.Generator(device=DEV).manual_seed(B * 31 + Cin + 5 * Cout)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=gen) * (2.0 / (Cin * 9)) ** 0.5
    yconv = torch.randn(B, H, W, Cout, device=DEV, generator=gen)
    dout = torch.randn(B, H, W, Cout, device=DEV, generator=gen) * 1e-6
    bn = torch.nn.BatchNorm2d(Cout).to(DEV)
    st = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    st[0] = yconv.double().sum((0, 1, 2)); st[1] = (yconv.double() ** 2).sum((0, 1, 2))
    co = ops.bn_finalize(st, B * H * W, bn, True)
    cw = ops.ConvWeights(w, g, L.PREC_FP16X2)
    a1 = torch.zeros(1, device=DEV)
    dy_ps, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a1, planes=True)
    monkeypatch.setenv("PC_HALO_ALL", "1")
    # the layer BELOW (whose BatchNorm backward consumes dx): Cin channels, its own pre-BN tensor, coefficients and dropout multipliers
    ylow = torch.randn(B, H, W, Cin, device=DEV, generator=gen) * 1.5 + 0.2
    bnl = torch.nn.BatchNorm2d(Cin).to(DEV)
    with torch.no_grad():
        bnl.weight.copy_(torch.rand(Cin, device=DEV, generator=gen) + 0.5); bnl.bias.copy_(torch.randn(Cin, device=DEV, generator=gen) * 0.3)
    stl = torch.zeros(2, Cin, device=DEV, dtype=torch.float64)
    stl[0] = ylow.double().sum((0, 1, 2)); stl[1] = (ylow.double() ** 2).sum((0, 1, 2))
    col = ops.bn_finalize(stl, B * H * W, bnl, True)
    drop = ((torch.rand(B, Cin, device=DEV, generator=gen) > 0.2).float() / 0.8) if drop_on else None
    fused = ops.conv_dgrad_bn_reduce(dy_ps, cw.wd, g, a1, ylow, col, drop, planes=True)
    assert fused is not None
    dx_f, (sums_f, maxes_f) = fused
    dx = ops.conv_dgrad(dy_ps, cw.wd, g, prec=cw.prec_d, dy_amax=a1, dy_presplit=True)
    assert torch.equal(dx, dx_f)
    sums = torch.zeros(2, Cin, device=DEV, dtype=torch.float64)
    maxes = torch.zeros(2, device=DEV)
    call("pc_bn_act_bwd_reduce", ptr(dx), ptr(ylow), B, H, W, Cin, ptr(col.scale), ptr(col.shift), ptr(col.mean), ptr(col.invstd), ptr(drop), 0,
         None, ptr(sums, torch.float64), ptr(maxes), stream())
    torch.cuda.synchronize()
    assert torch.equal(maxes, maxes_f), (maxes, maxes_f)
    ref_scale = float(sums.abs().max())
    assert float((sums - sums_f).abs().max()) <= 2e-5 * ref_scale, (float((sums - sums_f).abs().max()), ref_scale)
    # and the whole BatchNorm backward from the fused reduce equals the two-pass one
    m1, m2 = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    dg1, db1 = torch.empty(Cin, device=DEV), torch.empty(Cin, device=DEV)
    dg2, db2 = torch.empty(Cin, device=DEV), torch.empty(Cin, device=DEV)
    dy_a, _, _ = ops.bn_act_bwd(dx, ylow, col, 0, drop, None, dg1, db1, m1)
    dy_b, _, _ = ops.bn_act_bwd(dx, ylow, col, 0, drop, None, dg2, db2, m2, reduced=(sums_f, None))
    sc = float(dy_a.abs().max())
    assert float((dy_a - dy_b).abs().max()) <= 2e-5 * sc
    torch.testing.assert_close(dg1, dg2, rtol=1e-4, atol=1e-5 * float(dg1.abs().max()))
    torch.testing.assert_close(db1, db2, rtol=1e-4, atol=1e-5 * float(db1.abs().max()))


CASES_C32 = [  # B, H, W, Cin, Cout: the 32-channel layers of cnn_small (reference src/models/phoneme_cnn.py:36-60) and ragged variants
    (2, 40, 101, 32, 32),
    (3, 20, 50, 32, 64),
    (5, 10, 25, 32, 32),
    (1, 7, 9, 32, 64),
]


@pytest.mark.parametrize("case", CASES_C32)
def test_halo_32_channel_layers(case, monkeypatch):
    """32 gathered channels = one 64-channel chunk whose upper half the TMA boxes zero-fill (box wider than the tensor), 32 produced
    channels = the weight-resident kernel with a 32-column tile: forward (+ bias, BatchNorm sums) and data gradient against fp64."""
    from phoneme_contrast_b200 import _lib as L
    from phoneme_contrast_b200 import ops
    import ctypes as C
    B, H, W, Cin, Cout = case
    g = ops.conv_geom(B, H, W, Cin, Cout, 3, 1, 1)
    gen = torch.Generator(device=DEV).manual_seed(B * 13 + Cin + 7 * Cout)
    x = torch.randn(B, H, W, Cin, device=DEV, generator=gen)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=gen) * (2.0 / (Cin * 9)) ** 0.5
    bias = torch.randn(Cout, device=DEV, generator=gen)
    assert L.lib().pc_conv_halo_supported(C.byref(g), 0) == 1 and L.lib().pc_conv_halo_supported(C.byref(g), 1) == 1
    cw = ops.ConvWeights(w, g, L.PREC_FP16X2, planes_ok=True)
    assert cw.prec_f == L.PREC_FP16X2 and cw.prec_d == L.PREC_FP16X2
    assert ops.ConvWeights(w, g, L.PREC_FP16X2).prec_f == L.PREC_TF32X3        # without planes the layer stays on the TF32x3 engine
    planes = ops.bn_act_split(x)
    st = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    y = ops.conv_fwd(planes, cw.wf, bias, g, dict(presplit=True), st, cw.prec_f)
    ref = _ref_conv(x, w, bias)
    scale = float(ref.abs().max())
    assert float((y.double() - ref).abs().max()) <= 5e-6 * scale
    np.testing.assert_allclose(st[0].cpu().numpy(), ref.sum(dim=(0, 1, 2)).cpu().numpy(), rtol=2e-5, atol=2e-4 * scale)
    np.testing.assert_allclose(st[1].cpu().numpy(), (ref ** 2).sum(dim=(0, 1, 2)).cpu().numpy(), rtol=2e-5)
    # data gradient
    yconv = torch.randn(B, H, W, Cout, device=DEV, generator=gen)
    dout = torch.randn(B, H, W, Cout, device=DEV, generator=gen) * 1e-6
    bn = torch.nn.BatchNorm2d(Cout).to(DEV)
    s2 = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    s2[0] = yconv.double().sum((0, 1, 2)); s2[1] = (yconv.double() ** 2).sum((0, 1, 2))
    co = ops.bn_finalize(s2, B * H * W, bn, True)
    a0, a1 = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    dy, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a0)
    dy_ps, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a1, planes=True)
    dx = ops.conv_dgrad(dy_ps, cw.wd, g, prec=cw.prec_d, dy_amax=a1, dy_presplit=True)
    refd = torch.nn.functional.conv_transpose2d(dy.permute(0, 3, 1, 2).double(), w.double(), padding=1).permute(0, 2, 3, 1)
    sd = float(refd.abs().max())
    assert float((dx.double() - refd).abs().max()) <= 1e-5 * sd
    base = torch.randn(B, H, W, Cin, device=DEV, generator=gen) * sd
    acc = base.clone()
    ops.conv_dgrad(dy_ps, cw.wd, g, out=acc, accumulate=True, prec=cw.prec_d, dy_amax=a1, dy_presplit=True)
    assert float((acc.double() - (base.double() + refd)).abs().max()) <= 1e-5 * sd + 1e-6 * float(base.abs().max())


WG_CASES_C32 = [(4, 40, 101, 32, 32), (64, 40, 101, 32, 32), (6, 20, 50, 32, 64), (3, 9, 25, 32, 32), (2, 7, 9, 64, 32)]


@pytest.mark.parametrize("case", WG_CASES_C32)
def test_halo_wgrad_32_channel_layers(case):
    """Halo weight gradient with 32 input and / or output channels (cnn_small): the missing half of each 64-channel box is zero-filled
    by the TMA unit, the reduce writes only the channels that exist. Against fp64."""
    from phoneme_contrast_b200 import _lib as L
    from phoneme_contrast_b200 import ops
    import ctypes as C
    B, H, W, Cin, Cout = case
    g = ops.conv_geom(B, H, W, Cin, Cout, 3, 1, 1)
    gen = torch.Generator(device=DEV).manual_seed(B * 17 + Cin + 3 * Cout + H)
    x = torch.relu(torch.randn(B, H, W, Cin, device=DEV, generator=gen))
    yconv = torch.randn(B, H, W, Cout, device=DEV, generator=gen)
    dout = torch.randn(B, H, W, Cout, device=DEV, generator=gen) * 1e-6
    bn = torch.nn.BatchNorm2d(Cout).to(DEV)
    st = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    st[0] = yconv.double().sum((0, 1, 2)); st[1] = (yconv.double() ** 2).sum((0, 1, 2))
    co = ops.bn_finalize(st, B * H * W, bn, True)
    a0, a1 = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    dy, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a0)
    dy_ps, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a1, planes=True)
    x_ps = ops.bn_act_split(x)
    assert L.lib().pc_conv_wgrad_halo_supported(C.byref(g)) == 1
    dw = torch.full((Cout, Cin, 3, 3), float("nan"), device=DEV)
    ops.conv_wgrad(x_ps, dy_ps, g, dict(presplit=True), dw=dw, prec=L.PREC_FP16X2, dy_amax=a1, dy_presplit=True, want_db=False)
    torch.cuda.synchronize()
    xp = torch.nn.functional.pad(x.double().permute(0, 3, 1, 2), (1, 1, 1, 1))
    dyd = dy.double().permute(0, 3, 1, 2)
    ref = torch.empty(Cout, Cin, 3, 3, device=DEV, dtype=torch.float64)
    for r in range(3):
        for s in range(3):
            ref[:, :, r, s] = torch.einsum("bohw,bchw->oc", dyd, xp[:, :, r:r + H, s:s + W])
    scale = float(ref.abs().max())
    err = float((dw.double() - ref).abs().max())
    assert err <= 4e-5 * scale, (err, scale)      # (a reduction over 258 k positions per weight at the cnn_small shape: 2.03e-5 measured)
