"""Pin the CPU oracle against golden vectors produced by the real reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import augment_oracle, mfcc_oracle, nets_oracle, optim_oracle, supcon_oracle
from tests.golden.synth import synth_waves
from tests.helpers import analytically_zero_grad, grad_atol


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


# ----------------------------------------------------------------------------- MFCC
def test_mfcc_constants(golden):
    g = golden.mfcc
    np.testing.assert_allclose(mfcc_oracle.hann_periodic(400), g["window"], atol=5e-7)  # torch builds it in fp32
    np.testing.assert_allclose(mfcc_oracle.mel_filterbank(), g["fb"], atol=1e-5)  # 1-ulp powf difference, see oracle
    np.testing.assert_allclose(mfcc_oracle.dct_matrix(), g["dct"], atol=3e-6)  # torch evaluates cos() on fp32 arguments


@pytest.mark.parametrize("tag,s", [("s16000", 16000), ("s4000", 4000), ("s1234", 1234)])
def test_mfcc_vs_reference(golden, tag, s):
    g = golden.mfcc
    w = synth_waves(4, s, seed=int(g[f"{tag}_seed"]))
    per_clip = mfcc_oracle.mfcc(w, clamp_scope="clip")
    batched = mfcc_oracle.mfcc(w, clamp_scope="call")
    # MFCC tolerance (north_star): 1e-4 relative. The reference itself is fp32; our fp64 oracle
    # agrees with it to ~4e-6 rel-L2, well inside the budget.
    assert rel_l2(per_clip, g[f"{tag}_per_clip"]) < 2e-5
    assert rel_l2(batched, g[f"{tag}_batched"]) < 2e-5
    scale = np.abs(g[f"{tag}_per_clip"]).max()
    assert np.abs(per_clip - g[f"{tag}_per_clip"]).max() < 1e-4 * scale * 10


def test_mfcc_delta_mel_gain(golden):
    g = golden.mfcc
    w = synth_waves(2, 4000, seed=int(g["delta_in_seed"]))
    dd = mfcc_oracle.mfcc(w[0:1], add_delta=True, add_delta_delta=True)
    assert dd.shape == g["delta_dd"].shape == (1, 1, 120, 26)
    assert rel_l2(dd, g["delta_dd"]) < 2e-5
    assert rel_l2(mfcc_oracle.log_mel(w), g["mel_s4000"]) < 2e-5
    gain = np.float32(float(g["gain_value"]))
    assert rel_l2(mfcc_oracle.mfcc(w[1:2] * gain), g["gain_per_clip"]) < 2e-5


def test_mfcc_torch_cpu_variant(golden):
    g = golden.mfcc
    w = synth_waves(4, 4000, seed=7)
    assert rel_l2(mfcc_oracle.mfcc_torch_cpu(w, per_clip=True).numpy(), g["s4000_per_clip"]) < 1e-6
    assert rel_l2(mfcc_oracle.mfcc_torch_cpu(w, per_clip=False).numpy(), g["s4000_batched"]) < 1e-6


# ----------------------------------------------------------------------------- augmentation
def test_augment_decisions_bit_exact(golden):
    g = golden.augment
    rec = g["rec"]
    x = g["x"][0, 0]
    outs = g["outs"]
    k = 0
    for idx in range(rec.shape[0]):
        for v in range(rec.shape[1]):
            d = augment_oracle.view_descriptor(idx, v, 40, 101, noise_shape=(1, 1, 40, 101))
            gain, t0, t1, f0, f1, n_app, level = rec[idx, v, :7]
            assert np.float32(d["gain"]) == np.float32(gain)
            ta, ts, te = d["t"]
            assert ((ts, te) if ta and te > ts else (0, 0)) == (int(t0), int(t1))
            fa, fs, fe = d["f"]
            assert ((fs, fe) if fa and fe > fs else (0, 0)) == (int(f0), int(f1))
            assert d["noise"][0] == bool(n_app)
            assert d["noise"][1] == level
            if idx < 8:
                got = augment_oracle.apply_view(x, d)
                np.testing.assert_allclose(got, outs[k], rtol=0, atol=1e-6)
                k += 1


def test_augment_kats():
    # seed 1 -> time mask [21,43); idx 2 view 1 -> freq rows [4,11); gains
    assert augment_oracle.axis_mask_decision(1, 0.5, 30, 101) == (True, 21, 43)
    assert augment_oracle.axis_mask_decision(2 * 20000 + 1 + 1000, 0.5, 10, 40) == (True, 4, 11)
    assert augment_oracle.gain_decision(0) == (False, 1.0)
    a, gval = augment_oracle.gain_decision(1)
    assert a and abs(gval - 1.138973) < 1e-6


# ----------------------------------------------------------------------------- SupCon
def test_supcon_kats(golden):
    g = golden.supcon
    f8 = np.tile(np.eye(4), (2, 1))
    y8 = np.array([0, 0, 1, 1, 2, 2, 3, 3])
    for T, want in ((0.5, 18.53170204), (0.15, 14.30201626), (0.07, 14.28571987)):
        got = supcon_oracle.loss(f8, y8, temperature=T)
        assert abs(got - want) < 2e-5
        assert abs(got - float(g[f"kat_eye_T{T}"])) < 2e-5
    rs = torch.Generator().manual_seed(1234)
    f = torch.nn.functional.normalize(torch.randn(64, 128, generator=rs), dim=1).numpy()
    y = np.arange(64) // 8
    assert abs(supcon_oracle.loss(f, y, temperature=0.15) - 9.33036900) < 2e-5
    gr = supcon_oracle.grad(f, y, temperature=0.15)
    assert abs(np.linalg.norm(gr) - 1.30263817) < 2e-6
    np.testing.assert_allclose(gr[0, :3], [0.01415017, -0.02836241, 0.02471864], atol=2e-7)


@pytest.mark.parametrize("tag", ["n64_d128", "n37_d64", "n256_d128", "n130_d128_singletons", "n96_d256"])
def test_supcon_vs_reference(golden, tag):
    g = golden.supcon
    f, y = g[f"{tag}_f"], g[f"{tag}_y"]
    loss = supcon_oracle.loss(f, y, temperature=0.15)
    assert abs(loss - float(g[f"{tag}_loss"])) <= 1e-5 * max(1.0, abs(loss))
    gr = supcon_oracle.grad(f, y, temperature=0.15)
    assert np.abs(gr - g[f"{tag}_grad"]).max() < 1e-6
    # row-block (data-parallel) gradient == slice of the full gradient
    n = f.shape[0]
    half = slice(n // 2, n)
    np.testing.assert_allclose(supcon_oracle.grad(f, y, temperature=0.15, rows=half), gr[half])


def test_supcon_unnorm_sum_and_mask(golden):
    g = golden.supcon
    kw = dict(temperature=0.3, base_temperature=0.2, reduction="sum")
    assert abs(supcon_oracle.loss(g["unnorm_f"], g["unnorm_y"], **kw) - float(g["unnorm_loss"])) < 1e-4 * abs(float(g["unnorm_loss"]))
    assert rel_l2(supcon_oracle.grad(g["unnorm_f"], g["unnorm_y"], **kw), g["unnorm_grad"]) < 1e-5
    assert abs(supcon_oracle.loss(g["mask_f"], None, mask=g["mask_m"], temperature=0.15) - float(g["mask_loss"])) < 1e-5 * abs(float(g["mask_loss"]))
    assert rel_l2(supcon_oracle.grad(g["mask_f"], None, mask=g["mask_m"], temperature=0.15), g["mask_grad"]) < 1e-5


def test_supcon_n1_raises():
    with pytest.raises(ValueError):
        supcon_oracle.loss(np.ones((1, 8)), np.array([0]))


def test_supcon_torch_variant_matches(golden):
    g = golden.supcon
    f, y = torch.from_numpy(g["n64_d128_f"]), torch.from_numpy(g["n64_d128_y"])
    assert abs(float(supcon_oracle.loss_torch_cpu(f, y)) - float(g["n64_d128_loss"])) < 1e-5


# ----------------------------------------------------------------------------- nets
NET_CASES = [
    ("small", "phoneme_cnn", {"embedding_dim": 128, "use_attention": True, "dropout_rate": 0.0}),
    ("small_noattn_e64", "phoneme_cnn", {"embedding_dim": 64, "use_attention": False, "dropout_rate": 0.0}),
    ("deep_mini", "phoneme_cnn_deep", {"embedding_dim": 128, "use_attention": True, "dropout_rate": 0.0,
                                       "hidden_dims": [16, 32, 64, 128]}),
    ("deep_mini_odd", "phoneme_cnn_deep", {"embedding_dim": 32, "use_attention": True, "dropout_rate": 0.0,
                                           "hidden_dims": [16, 16, 32, 32]}),
]


@pytest.mark.parametrize("tag,arch,cfg", NET_CASES)
def test_nets_vs_reference(golden, tag, arch, cfg):
    g = golden.nets
    sd = nets_oracle.synthetic_state_dict(arch, cfg, seed=21)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k}
    live = dict(sd)
    live.update(params)
    x = torch.from_numpy(g[f"{tag}_x"])
    y = torch.from_numpy(g[f"{tag}_y"])
    emb = nets_oracle.forward(arch, live, x, training=True, use_attention=cfg["use_attention"])
    np.testing.assert_allclose(emb.detach().numpy(), g[f"{tag}_emb_train"], atol=2e-6)
    loss = supcon_oracle.loss_torch_cpu(emb, y, temperature=0.15)
    assert abs(float(loss.detach()) - float(g[f"{tag}_loss"])) < 1e-5 * abs(float(g[f"{tag}_loss"]))
    loss.backward()
    for name in g[f"{tag}_param_names"]:
        name = str(name)
        gr = params[name].grad.numpy()
        want = float(g[f"{tag}_gnorm_{name}"])
        if not analytically_zero_grad(name):
            assert abs(np.linalg.norm(gr.astype(np.float64)) - want) <= 1e-4 * want, name
        key = f"{tag}_grad_{name}"
        if key in g.files:
            np.testing.assert_allclose(gr, g[key], rtol=0, atol=grad_atol(name, g[key]), err_msg=name)
    for k in g.files:
        if k.startswith(f"{tag}_after_"):
            np.testing.assert_allclose(live[k[len(tag) + 7:]].numpy(), g[k], atol=1e-6, err_msg=k)
    emb_eval = nets_oracle.forward(arch, live, x, training=False, use_attention=cfg["use_attention"])
    np.testing.assert_allclose(emb_eval.detach().numpy(), g[f"{tag}_emb_eval"], atol=2e-6)


def test_param_inventory(golden):
    g = golden.nets
    for arch in ("phoneme_cnn", "phoneme_cnn_deep"):
        shapes = nets_oracle.param_shapes(arch, {})
        assert [n for n, _, _ in shapes] == [str(k) for k in g[f"{arch}_keys"]]
        n_params = sum(int(np.prod(s)) for n, s, k in shapes if k in ("conv_w", "lin_w", "bn_w", "bias"))
        assert n_params == int(g[f"{arch}_n_params"])


# ----------------------------------------------------------------------------- optimiser
def test_clip_adam_vs_torch(golden):
    g = golden.optim
    n = int(g["n"])
    params = [g[f"p0_{i}"].copy() for i in range(n)]
    m = [np.zeros_like(p) for p in params]
    v = [np.zeros_like(p) for p in params]
    for step in range(3):
        grads = [g[f"g{step}_{i}"] for i in range(n)]
        total = optim_oracle.adam_step(params, grads, m, v, step + 1, lr=3e-4, weight_decay=1e-4, max_norm=1.0)
        assert abs(total - float(g[f"norm{step}"])) < 1e-5 * total
        for i in range(n):
            np.testing.assert_allclose(params[i], g[f"p{step + 1}_{i}"], rtol=0, atol=2e-7)
