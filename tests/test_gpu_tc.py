"""Unit checks of the tcgen05 tensor-core tile engine (csrc/conv_tc.cu) against fp64 references:
plain GEMM first (descriptor / swizzle / TMEM plumbing), then convolution forward and data-gradient
against the exact-fp32 SIMT kernels and torch's fp64 conv."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"

TOL = {1: 2e-5, 2: 1e-2, 3: 2e-5}     # PC_PREC_TF32X3 (fp32-level), PC_PREC_BF16, PC_PREC_FP16X2 (fp32-level); relative to max |ref|


@pytest.mark.parametrize("prec", [1, 2, 3])
@pytest.mark.parametrize("M,N,K", [(128, 32, 64), (300, 64, 128), (1000, 128, 576), (4096, 256, 1152), (77, 16, 64), (260, 512, 256)])
def test_tc_gemm(prec, M, N, K):
    from phoneme_contrast_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N)
    a = torch.randn(M, K, device=DEV, generator=g)
    b = torch.randn(N, K, device=DEV, generator=g)
    bias = torch.randn(N, device=DEV, generator=g)
    c = ops.tc_gemm(a, b, bias, prec)
    ref = (a.double() @ b.double().T + bias.double())
    err = float((c.double() - ref).abs().max() / ref.abs().max())
    assert err < TOL[prec], err


CONV_CASES = [
    # B, H, W, Cin, Cout, k, stride, pad
    (4, 20, 51, 64, 64, 3, 1, 1),
    (3, 20, 51, 64, 128, 3, 2, 1),
    (3, 20, 51, 64, 128, 1, 2, 0),
    (2, 10, 26, 128, 256, 3, 2, 1),
    (5, 3, 7, 512, 512, 3, 1, 1),
    (2, 40, 101, 32, 32, 3, 1, 1),
    (2, 20, 50, 32, 64, 3, 1, 1),
]


@pytest.mark.parametrize("prec", [1, 2, 3])
@pytest.mark.parametrize("case", CONV_CASES)
def test_tc_conv_fwd_dgrad(prec, case):
    from phoneme_contrast_b200 import _lib as L
    from phoneme_contrast_b200 import ops
    B, H, W, Cin, Cout, k, stride, pad = case
    g = ops.conv_geom(B, H, W, Cin, Cout, k, stride, pad)
    if not ops.tc_supported(g, False, prec):
        pytest.skip("layer not eligible for this precision")
    gen = torch.Generator(device=DEV).manual_seed(Cin + Cout + k)
    x = torch.randn(B, H, W, Cin, device=DEV, generator=gen)
    w = torch.randn(Cout, Cin, k, k, device=DEV, generator=gen) * (2.0 / (Cin * k * k)) ** 0.5
    bias = torch.randn(Cout, device=DEV, generator=gen)
    scale = 1.0 + 0.1 * torch.randn(Cin, device=DEV, generator=gen)
    shift = 0.1 * torch.randn(Cin, device=DEV, generator=gen)
    drop = ((torch.rand(B, Cin, device=DEV, generator=gen) > 0.2).float() / 0.8).contiguous()
    xf = dict(scale=scale, shift=shift, relu=True, drop=drop)
    cw = ops.ConvWeights(w, g, prec)
    assert cw.prec_f == prec
    stats = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    y = ops.conv_fwd(x, cw.wf, bias, g, xf, stats, cw.prec_f)
    a = torch.relu(x.double() * scale.double() + shift.double()) * drop.double()[:, None, None, :]
    ref = torch.nn.functional.conv2d(a.permute(0, 3, 1, 2), w.double(), bias.double(), stride=stride, padding=pad).permute(0, 2, 3, 1)
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    assert err < TOL[prec], ("fwd", err)
    # BatchNorm statistics from the epilogue
    np.testing.assert_allclose(stats[0].cpu().numpy(), ref.sum(dim=(0, 1, 2)).cpu().numpy(), rtol=5 * TOL[prec], atol=5 * TOL[prec] * float(ref.abs().sum(dim=(0, 1, 2)).max()))
    np.testing.assert_allclose(stats[1].cpu().numpy(), (ref * ref).sum(dim=(0, 1, 2)).cpu().numpy(), rtol=5 * TOL[prec])
    # data gradient (+ accumulate)
    if cw.prec_d == prec:
        # FP16X2: gradients as small as real ones (1e-7) must survive fp16's range through the amax operand scale
        dy = torch.randn(B, g.Ho, g.Wo, Cout, device=DEV, generator=gen) * (3e-7 if prec == 3 else 1.0)
        amax = dy.abs().max().reshape(1) if prec == 3 else None
        dx = ops.conv_dgrad(dy, cw.wd, g, prec=cw.prec_d, dy_amax=amax)
        xr = torch.zeros(B, Cin, H, W, device=DEV, dtype=torch.float64, requires_grad=True)
        torch.nn.functional.conv2d(xr, w.double(), None, stride=stride, padding=pad).backward(dy.permute(0, 3, 1, 2).double())
        dref = xr.grad.permute(0, 2, 3, 1)
        err = float((dx.double() - dref).abs().max() / dref.abs().max())
        assert err < TOL[prec], ("dgrad", err)
        dx2 = ops.conv_dgrad(dy, cw.wd, g, out=dx.clone(), accumulate=True, prec=cw.prec_d, dy_amax=amax)
        assert float((dx2.double() - 2 * dref).abs().max() / dref.abs().max()) < 2 * TOL[prec]


@pytest.mark.parametrize("case", CONV_CASES + [(16, 40, 101, 32, 32, 3, 1, 1), (9, 20, 51, 64, 64, 3, 1, 1), (4, 10, 50, 128, 128, 3, 1, 1),
                                               (4, 20, 100, 64, 64, 3, 1, 1), (4, 5, 25, 256, 256, 3, 1, 1), (4, 20, 100, 64, 128, 3, 2, 1)])
@pytest.mark.parametrize("prec", [1, 3])
def test_tc_conv_wgrad(case, prec):
    """Weight / bias gradient on the tensor cores (MN-major TF32x3 or FP16x2 tiles, pixel-split partials) vs torch fp64."""
    from phoneme_contrast_b200 import ops
    B, H, W, Cin, Cout, k, stride, pad = case
    g = ops.conv_geom(B, H, W, Cin, Cout, k, stride, pad)
    gen = torch.Generator(device=DEV).manual_seed(Cin * 3 + Cout + k)
    x = torch.randn(B, H, W, Cin, device=DEV, generator=gen)
    # FP16X2: realistic tiny gradients; the amax operand scale must keep them inside fp16's range
    dy = torch.randn(B, g.Ho, g.Wo, Cout, device=DEV, generator=gen) * (3e-7 if prec == 3 else 1.0)
    amax = dy.abs().max().reshape(1) if prec == 3 else None
    scale = 1.0 + 0.1 * torch.randn(Cin, device=DEV, generator=gen)
    shift = 0.1 * torch.randn(Cin, device=DEV, generator=gen)
    drop = ((torch.rand(B, Cin, device=DEV, generator=gen) > 0.2).float() / 0.8).contiguous()
    xf = dict(scale=scale, shift=shift, relu=True, drop=drop)
    dw, db = ops.conv_wgrad(x, dy, g, xf, prec=prec, dy_amax=amax)
    a = (torch.relu(x.double() * scale.double() + shift.double()) * drop.double()[:, None, None, :]).permute(0, 3, 1, 2)
    wr = torch.zeros(Cout, Cin, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    br = torch.zeros(Cout, device=DEV, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.conv2d(a, wr, br, stride=stride, padding=pad).backward(dy.permute(0, 3, 1, 2).double())
    err = float((dw.double() - wr.grad).abs().max() / wr.grad.abs().max())
    assert err < TOL[prec], ("dw", err)
    assert float((db.double() - br.grad).abs().max() / br.grad.abs().max()) < 1e-5


@pytest.mark.parametrize("B,H,W,Cout,k", [(3, 40, 101, 64, 7), (2, 17, 33, 32, 7), (2, 40, 50, 64, 3), (1, 9, 12, 64, 7)])
def test_tc_stem_wgrad(B, H, W, Cout, k):
    """Single-channel stem weight gradient on the tensor cores (taps as M rows, FP16X2) vs torch fp64."""
    from phoneme_contrast_b200 import ops
    g = ops.conv_geom(B, H, W, 1, Cout, k, 1, k // 2)
    gen = torch.Generator(device=DEV).manual_seed(B * 13 + Cout + k)
    x = torch.randn(B, H, W, 1, device=DEV, generator=gen) * 3.0
    dy = torch.randn(B, H, W, Cout, device=DEV, generator=gen) * 2e-7
    amax = dy.abs().max().reshape(1)
    dw, db = ops.conv_wgrad(x, dy, g, None, prec=3, dy_amax=amax)
    wr = torch.zeros(Cout, 1, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    br = torch.zeros(Cout, device=DEV, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), wr, br, padding=k // 2).backward(dy.permute(0, 3, 1, 2).double())
    assert float((dw.double() - wr.grad).abs().max() / wr.grad.abs().max()) < TOL[3]
    assert float((db.double() - br.grad).abs().max() / br.grad.abs().max()) < 1e-5
    dw0, _ = ops.conv_wgrad(x, dy, g, None, prec=0)       # exact-fp32 SIMT stem kernel
    assert float((dw - dw0).abs().max() / dw0.abs().max()) < TOL[3]


@pytest.mark.parametrize("precision", ["tf32x3", "fp16x2"])
@pytest.mark.parametrize("arch,B", [("phoneme_cnn", 16), ("phoneme_cnn_deep", 8)])
def test_nets_split_precision_match_oracle(arch, B, precision, monkeypatch):
    """Whole networks with the tensor-core convolutions in a split-operand mode (TF32x3 / FP16x2) keep the fp32 parity bar (1e-4)."""
    monkeypatch.setenv("PC_PRECISION", precision)
    from tests.test_gpu_parity import _check_grads, _oracle_net, _run_net
    from oracle import nets_oracle
    cfg = {"dropout_rate": 0.0}
    sd = nets_oracle.synthetic_state_dict(arch, cfg, seed=3)
    rs = np.random.RandomState(8)
    x = rs.standard_normal((B, 1, 40, 101)).astype(np.float32)
    y = (np.arange(B) // 2).astype(np.int64)
    _, emb, loss, grads = _run_net(arch, cfg, sd, x, y)
    emb_ref, loss_ref, grads_ref, _ = _oracle_net(arch, cfg, sd, x, y)
    assert np.abs(emb - emb_ref).max() <= 1e-4 * np.abs(emb_ref).max()
    assert abs(loss - loss_ref) <= 1e-4 * abs(loss_ref)
    _check_grads(grads, grads_ref)


def test_bf16_network_mode_withdrawn(monkeypatch):
    """Round 1 offered a single-product bf16 mode; it measures 1.5e-2 .. 2.9e-2 on the embeddings (profiles/tools/bf16_diag.py),
    outside north_star's 1e-2 bar, so selecting it for a network is an error rather than a silently inaccurate run."""
    from phoneme_contrast_b200.models import model_registry
    with pytest.raises(ValueError, match="withdrawn"):
        model_registry.create("phoneme_cnn_deep", {"precision": "bf16"})
    monkeypatch.setenv("PC_PRECISION", "bf16")
    with pytest.raises(ValueError, match="withdrawn"):
        model_registry.create("phoneme_cnn", {})


@pytest.mark.parametrize("n,d,row0,nrows", [(1024, 128, 0, 1024), (1100, 128, 0, 1100), (2000, 64, 0, 2000), (2048, 128, 512, 300),
                                            (4096, 128, 1024, 1024), (130, 128, 8, 100)])
def test_supcon_tc_forward_matches_simt_and_oracle(n, d, row0, nrows, monkeypatch):
    """tcgen05 SupCon forward (similarity tiles in TMEM, FP16x2 split) vs the exact-fp32 SIMT kernel and the numpy oracle:
    row max / denominators within 1e-5, positive counts bit-exact, loss within 1e-5."""
    from oracle import supcon_oracle
    from phoneme_contrast_b200 import ops
    rs = np.random.RandomState(n + d)
    f = rs.standard_normal((n, d)).astype(np.float32)
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    y = rs.randint(0, 38, n)
    ft, yt = torch.from_numpy(f).to(DEV), torch.from_numpy(y).to(DEV)
    monkeypatch.setenv("PC_SUPCON_TC", "0")
    st0, rl0 = ops.supcon_fwd(ft, yt, None, 0.15, 0.07, row0, nrows)
    monkeypatch.setenv("PC_SUPCON_TC", "1")
    from phoneme_contrast_b200 import _lib as L
    before = L.lib().pc_launch_count()
    st1, rl1 = ops.supcon_fwd(ft, yt, None, 0.15, 0.07, row0, nrows)
    assert L.lib().pc_launch_count() - before == 3          # pack + tensor-core kernel + merge: the TC path really ran
    s0, s1 = st0.cpu().numpy(), st1.cpu().numpy()
    np.testing.assert_allclose(s1[:, 0], s0[:, 0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(s1[:, 1], s0[:, 1], rtol=1e-5)
    assert np.array_equal(s1[:, 2], s0[:, 2])
    np.testing.assert_allclose(s1[:, 3], s0[:, 3], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(rl1.cpu().numpy(), rl0.cpu().numpy(), rtol=2e-5, atol=1e-6)
    if row0 == 0 and nrows == n and n <= 2048:
        want = supcon_oracle.loss(f, y, temperature=0.15)
        assert abs(float(rl1.double().mean()) - want) <= 1e-5 * abs(want)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("n,d,row0,nrows", [(1024, 128, 0, 1024), (1100, 128, 0, 1100), (2000, 64, 0, 2000), (2048, 128, 512, 300),
                                            (4096, 128, 1024, 1024), (130, 128, 8, 100)])
def test_supcon_tc_backward_matches_simt_and_oracle(n, d, row0, nrows, monkeypatch):
    """tcgen05 SupCon backward (S recomputed in TMEM, dL/dS as the fp16 hi/lo operand of a second MMA accumulating dF in TMEM)
    vs the exact-fp32 SIMT kernel (1e-4 of max|dF|) and, for the small full problems, the numpy oracle."""
    from oracle import supcon_oracle
    from phoneme_contrast_b200 import _lib as L
    from phoneme_contrast_b200 import ops
    rs = np.random.RandomState(n + d + 1)
    f = rs.standard_normal((n, d)).astype(np.float32)
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    y = rs.randint(0, 38, n)
    ft, yt = torch.from_numpy(f).to(DEV), torch.from_numpy(y).to(DEV)
    monkeypatch.setenv("PC_SUPCON_TC", "0")
    stats_all, _ = ops.supcon_fwd(ft, yt, None, 0.15, 0.07)
    coef = (0.15 / 0.07) / n
    gs = torch.full((1,), 0.75, device=DEV)
    g0 = ops.supcon_bwd(ft, yt, None, 0.15, coef, gs, stats_all, row0, nrows)
    monkeypatch.setenv("PC_SUPCON_TC", "1")
    before = L.lib().pc_launch_count()
    g1 = ops.supcon_bwd(ft, yt, None, 0.15, coef, gs, stats_all, row0, nrows)
    torch.cuda.synchronize()
    assert L.lib().pc_launch_count() - before == 3          # pack + tensor-core kernel + split reduce
    err = float((g1 - g0).abs().max() / g0.abs().max())
    assert err <= 1e-4, err
    if row0 == 0 and nrows == n and n <= 2048:
        gref = 0.75 * supcon_oracle.grad(f, y, temperature=0.15)
        assert np.abs(g1.cpu().numpy() - gref).max() <= 1e-4 * np.abs(gref).max()


@pytest.mark.parametrize("case", [(4, 20, 51, 64, 64, 3, 1, 1), (3, 20, 51, 64, 128, 3, 2, 1), (2, 10, 26, 128, 256, 3, 2, 1), (5, 3, 7, 512, 512, 3, 1, 1),
                                  (3, 20, 51, 64, 128, 1, 2, 0)])
def test_tc_conv_fwd_presplit_input(case, monkeypatch):
    """A convolution reading the activation as pre-split fp16 planes (pc_bn_act_split) gives bit-identical outputs to the one
    that applies BatchNorm + ReLU + dropout + split inside its gather: the operand bytes are the same. (Per-tap-gather kernel on
    both sides: the halo engine sums the same products in a different order and is compared separately in test_gpu_halo.py.)"""
    from phoneme_contrast_b200 import ops
    monkeypatch.setenv("PC_CONV_HALO", "0")
    B, H, W, Cin, Cout, k, stride, pad = case
    g = ops.conv_geom(B, H, W, Cin, Cout, k, stride, pad)
    gen = torch.Generator(device=DEV).manual_seed(Cin + 2 * Cout + k)
    x = torch.randn(B, H, W, Cin, device=DEV, generator=gen)
    w = torch.randn(Cout, Cin, k, k, device=DEV, generator=gen) * (2.0 / (Cin * k * k)) ** 0.5
    bias = torch.randn(Cout, device=DEV, generator=gen)
    scale = 1.0 + 0.1 * torch.randn(Cin, device=DEV, generator=gen)
    shift = 0.1 * torch.randn(Cin, device=DEV, generator=gen)
    drop = ((torch.rand(B, Cin, device=DEV, generator=gen) > 0.2).float() / 0.8).contiguous()
    cw = ops.ConvWeights(w, g, 3)
    y0 = ops.conv_fwd(x, cw.wf, bias, g, dict(scale=scale, shift=shift, relu=True, drop=drop), None, cw.prec_f)
    planes = ops.bn_act_split(x, scale, shift, drop, relu=True)
    st = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    y1 = ops.conv_fwd(planes, cw.wf, bias, g, dict(presplit=True), st, cw.prec_f)
    assert torch.equal(y0, y1)
    np.testing.assert_allclose(st[0].cpu().numpy(), y0.double().sum(dim=(0, 1, 2)).cpu().numpy(), rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("case", [(4, 20, 51, 64, 64, 3, 1, 1), (3, 20, 51, 64, 128, 3, 2, 1), (5, 3, 7, 512, 512, 3, 1, 1)])
def test_gradient_planes_match_fp32_gradients(case):
    """BatchNorm backward writing dy only as scaled fp16 hi | lo planes (scale from a bound of |dy|) + dgrad / wgrad gathering
    those bytes == the fp32-dy path (same FP16X2 engine) to 2e-5 of the result's max."""
    from phoneme_contrast_b200 import ops
    B, H, W, Cin, Cout, k, stride, pad = case
    g = ops.conv_geom(B, H, W, Cin, Cout, k, stride, pad)
    gen = torch.Generator(device=DEV).manual_seed(Cin + 3 * Cout + k)
    x = torch.randn(B, H, W, Cin, device=DEV, generator=gen)
    w = torch.randn(Cout, Cin, k, k, device=DEV, generator=gen) * (2.0 / (Cin * k * k)) ** 0.5
    yconv = torch.randn(B, g.Ho, g.Wo, Cout, device=DEV, generator=gen)          # pre-BN conv output
    dout = torch.randn(B, g.Ho, g.Wo, Cout, device=DEV, generator=gen) * 1e-6    # realistic gradient magnitude
    bn = torch.nn.BatchNorm2d(Cout).to(DEV)
    st = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    st[0] = yconv.double().sum((0, 1, 2)); st[1] = (yconv.double() ** 2).sum((0, 1, 2))
    co = ops.bn_finalize(st, B * g.Ho * g.Wo, bn, True)
    cw = ops.ConvWeights(w, g, 3)
    a0, a1 = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    dy, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a0)
    dy_ps, _, _ = ops.bn_act_bwd(dout, yconv, co, 0, None, None, amax=a1, planes=True)
    assert float(a1) >= float(a0) > 0 and float(a1) <= 64 * float(a0)            # the bound is above, and near, the true max
    dx0 = ops.conv_dgrad(dy, cw.wd, g, prec=cw.prec_d, dy_amax=a0)
    dx1 = ops.conv_dgrad(dy_ps, cw.wd, g, prec=cw.prec_d, dy_amax=a1, dy_presplit=True)
    assert float((dx0 - dx1).abs().max()) <= 2e-5 * float(dx0.abs().max())
    planes = ops.bn_act_split(x)
    dw0, db0 = ops.conv_wgrad(x, dy, g, None, prec=3, dy_amax=a0)
    dw1, db1 = ops.conv_wgrad(planes, dy_ps, g, dict(presplit=True), prec=3, dy_amax=a1, dy_presplit=True)
    assert float((dw0 - dw1).abs().max()) <= 2e-5 * float(dw0.abs().max())
    assert float((db0 - db1).abs().max()) <= 1e-5 * float(dy.abs().sum(dim=(0, 1, 2)).max())
