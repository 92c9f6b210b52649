"""Round-2 GPU parity tests (pytest -m gpu): the BENCHMARKED configurations against the oracle (cnn_deep at 256 views and
cnn_small at 64 views, Dropout2d masks injected, eagerly and through ContrastiveTrainer.step with cuda_graph=True), the
reference goldens of tests/golden/r2.npz (NTXent with labels, use_residual=False blocks, near-silent clip with gain,
pre-emphasis, dataset __getitem__ / _pad_or_trim, a reference-written checkpoint), and the regression tests for the round-1
advisor findings (weight-packer lifetime under graphs, stem weight gradient at tiny |dy|, fp16 range flag).
Tolerances: embeddings / loss 1e-4 relative, gradients rel-L2 <= 3e-3 per tensor (tests/test_gpu_parity.py:GRAD_RTOL)."""
import logging
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import mfcc_oracle, nets_oracle, supcon_oracle
from tests.helpers import analytically_zero_grad, rel_l2
from tests.test_gpu_parity import GRAD_RTOL, _check_grads, cu

pytestmark = pytest.mark.gpu
DEV = "cuda"
HERE = os.path.dirname(os.path.abspath(__file__))

BENCH_SHAPES = {
    "phoneme_cnn_deep": dict(views=256, chans=[64, 128, 256, 512], p=0.2),     # BASELINE configs[2]
    "phoneme_cnn": dict(views=64, chans=[32, 64, 128], p=0.1),                 # BASELINE configs[0] (8 classes x 4 x 2 views)
}
_oracle_cache = {}
# Gradient tolerance AT THE BENCHMARKED BATCH SIZE. Measured on B200 (profiles/tools/stem_grad_diag.py, profiles/r2_notes.md): against
# the CPU oracle the exact-fp32 SIMT path of this library shows up to 2.6e-3 per-tensor rel-L2 at 256 views, the default fp16x2 path
# 2.8e-3 .. 7.3e-3 from run to run (identical inputs): the spread is not arithmetic error -- embeddings agree to 1.5e-5 and the loss to
# 4e-7 -- but ReLU / max-pool gates that flip when a pre-activation lies within rounding distance of a tie (66 M stem activations per
# step; the BatchNorm statistics are accumulated with atomics, so the last bit differs between runs). Each flip moves one element of dz
# by a full dout element. The per-tensor bar at this size is therefore 1e-2 rel-L2 (3e-3 stays the bar of the small-batch parity
# tests, where flips are rare), plus the same bound on the whole flattened gradient.
BENCH_GRAD_RTOL = 1e-2


def _check_grads_bench(grads, ref):
    num = den = 0.0
    for name, gr in ref.items():
        got = grads[name]
        if analytically_zero_grad(name):
            assert np.abs(got).max() < 2e-4, name
            continue
        d = np.linalg.norm((got - gr).astype(np.float64))
        n = np.linalg.norm(gr.astype(np.float64))
        assert d <= BENCH_GRAD_RTOL * max(n, 1e-12), (name, "rel-L2", d / max(n, 1e-12))
        num += d * d
        den += n * n
    assert num ** 0.5 <= BENCH_GRAD_RTOL * den ** 0.5, ("flat gradient rel-L2", (num / den) ** 0.5)


def _bench_case(arch):
    """Inputs of bench.py's workload (train_inputs: randn views, labels repeat_interleave(arange(V/2)//4, 2)), reference-default
    dropout rates with explicit masks, synthetic parameters; oracle embeddings / loss / gradients computed once per arch."""
    if arch in _oracle_cache:
        return _oracle_cache[arch]
    spec = BENCH_SHAPES[arch]
    V = spec["views"]
    cfg = {"dropout_rate": spec["p"]}
    sd = nets_oracle.synthetic_state_dict(arch, cfg, seed=31)
    g = torch.Generator().manual_seed(1000)
    x = torch.randn(V, 1, 40, 101, generator=g)
    y = torch.repeat_interleave(torch.arange(V // 2) // 4, 2).to(torch.int64)
    rs = np.random.RandomState(5)
    masks = [((rs.uniform(size=(V, c)) >= spec["p"]) / (1.0 - spec["p"])).astype(np.float32) for c in spec["chans"]]
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k}
    live = {k: v.clone() for k, v in sd.items()}
    live.update(params)
    torch.set_num_threads(os.cpu_count() or 1)
    emb = nets_oracle.forward(arch, live, x, training=True, drop=[torch.from_numpy(m) for m in masks])
    loss = supcon_oracle.loss_torch_cpu(emb, y, temperature=0.15)
    loss.backward()
    out = dict(cfg=cfg, sd=sd, x=x, y=y, masks=masks, emb=emb.detach().numpy(), loss=float(loss.detach()),
               grads={k: p.grad.numpy() for k, p in params.items()})
    _oracle_cache[arch] = out
    return out


def _trainer(arch, case, graph):
    from phoneme_contrast_b200.models import model_registry
    from phoneme_contrast_b200.training import ContrastiveTrainer, FusedClipAdam, get_loss_fn
    m = model_registry.create(arch, dict(case["cfg"])).to(DEV)
    m.load_state_dict({k: v.clone() for k, v in case["sd"].items()})
    m._inject_drop = [cu(k) for k in case["masks"]]
    m.train()
    opt = FusedClipAdam(m.parameters(), lr=3e-4, weight_decay=1e-4)
    tr = ContrastiveTrainer(m, [], None, get_loss_fn("supervised_contrastive", temperature=0.15), opt, None, torch.device(DEV),
                            {"gradient_clip_val": 1.0, "cuda_graph": graph, "progress": False}, tempfile.mkdtemp(), logging.getLogger("t"))
    return m, tr


@pytest.mark.parametrize("arch", ["phoneme_cnn_deep", "phoneme_cnn"])
def test_bench_shape_eager_vs_oracle(arch):
    """The benchmarked step's forward + loss + backward, default precision (fp16x2), at the benchmarked batch size."""
    from phoneme_contrast_b200.training import get_loss_fn
    case = _bench_case(arch)
    m, _ = _trainer(arch, case, False)
    emb = m(case["x"].to(DEV))
    loss = get_loss_fn("supervised_contrastive", temperature=0.15)(emb, case["y"].to(DEV))
    loss.backward()
    e = emb.detach().cpu().numpy()
    assert np.abs(e - case["emb"]).max() <= 1e-4 * np.abs(case["emb"]).max(), np.abs(e - case["emb"]).max()
    assert abs(float(loss) - case["loss"]) <= 1e-4 * abs(case["loss"]), (float(loss), case["loss"])
    _check_grads_bench({n: p.grad.detach().cpu().numpy() for n, p in m.named_parameters()}, case["grads"])


@pytest.mark.parametrize("arch", ["phoneme_cnn_deep", "phoneme_cnn"])
def test_bench_shape_graph_replay_vs_oracle(arch):
    """Same step through ContrastiveTrainer.step with cuda_graph=True (what bench.py times): the replayed graph's loss and the
    gradients it leaves in .grad (computed at the pre-update parameters) match the oracle; a second replay on the same batch
    gives a different (post-Adam) loss, i.e. the replay really trains."""
    case = _bench_case(arch)
    m, tr = _trainer(arch, case, True)
    x, y = case["x"].to(DEV), case["y"].to(DEV)
    loss = float(tr.step(x, y))
    assert tr._graphed, "graph capture fell back to eager launches"
    assert abs(loss - case["loss"]) <= 1e-4 * abs(case["loss"]), (loss, case["loss"])
    _check_grads_bench({n: p.grad.detach().cpu().numpy() for n, p in m.named_parameters()}, case["grads"])
    assert int(tr.optimizer._step_dev.item()) == 1 and tr.optimizer._step == 1
    loss2 = float(tr.step(x, y))
    assert np.isfinite(loss2) and loss2 != loss


def test_graphed_step_survives_interleaved_eval_and_odd_batches():
    """an eval forward / odd-shaped eager batch between two graph replays must not free the weight-packer
    job table and operand buffers the captured launches point at. Default precision, so the packer has jobs."""
    case = _bench_case("phoneme_cnn")
    m, tr = _trainer("phoneme_cnn", case, True)
    m2, tr2 = _trainer("phoneme_cnn", case, True)
    x, y = case["x"].to(DEV), case["y"].to(DEV)
    l_a, l_b = float(tr.step(x, y)), float(tr2.step(x, y))
    assert abs(l_a - l_b) <= 1e-5 * abs(l_a)
    # trainer 1: eval forward + odd-shaped train step + allocator churn between the replays; trainer 2: nothing
    m.eval()
    with torch.no_grad():
        m(x[:10])
    m.train()
    junk = [torch.full((1 << 18,), float("nan"), device=DEV) for _ in range(8)]      # reuse whatever the packer might have freed
    snap = [p.detach().clone() for p in m.parameters()]
    opt_state = (tr.optimizer.flat_m.clone(), tr.optimizer.flat_v.clone(), tr.optimizer._step_dev.clone(), tr.optimizer._step)
    tr.step(x[:32], y[:32])                                                           # eager fallback (different shape)
    with torch.no_grad():                                                             # undo its update so both trainers stay comparable
        for p, s in zip(m.parameters(), snap):
            p.copy_(s)
        tr.optimizer.flat_m.copy_(opt_state[0]); tr.optimizer.flat_v.copy_(opt_state[1])
        tr.optimizer._step_dev.copy_(opt_state[2]); tr.optimizer._step = opt_state[3]
        for b, b2 in zip(m.buffers(), m2.buffers()):
            b.copy_(b2)
    del junk
    l_a2, l_b2 = float(tr.step(x, y)), float(tr2.step(x, y))
    assert np.isfinite(l_a2) and abs(l_a2 - l_b2) <= 1e-4 * abs(l_b2), (l_a2, l_b2)
    for p, q in zip(m.parameters(), m2.parameters()):
        assert torch.isfinite(p).all()
        assert float((p - q).abs().max()) <= 4 * 3e-4 + 1e-6


# ============================================================================================= reference goldens (r2.npz)
def test_ntxent_with_labels_vs_reference_golden(golden):
    """NTXentLoss(labels) (reference losses.py:114-151) == the SupCon kernel with base_temperature = temperature."""
    from phoneme_contrast_b200.training import get_loss_fn
    g = golden.r2
    for tag, T, red in (("t07_mean", 0.07, "mean"), ("t15_mean", 0.15, "mean"), ("t15_sum", 0.15, "sum")):
        ft = cu(g["ntx_f"]).requires_grad_(True)
        loss = get_loss_fn("ntxent", temperature=T, reduction=red)(ft, cu(g["ntx_y"], torch.int64))
        loss.backward()
        want = float(g[f"ntx_{tag}_loss"])
        assert abs(float(loss) - want) <= 1e-4 * abs(want), tag
        gref = g[f"ntx_{tag}_grad"]
        assert np.abs(ft.grad.cpu().numpy() - gref).max() <= 1e-4 * np.abs(gref).max() + 1e-8, tag
    with pytest.raises(NotImplementedError):
        get_loss_fn("ntxent")(cu(g["ntx_f"]))
    with pytest.raises(ValueError):
        get_loss_fn("ntxent")(cu(g["ntx_f"][:63]))


def test_plain_blocks_use_residual_false_vs_reference_golden(golden):
    """PhonemeNetDeep(use_residual=False) (reference phoneme_cnn.py:229-245): same state_dict keys, embeddings, loss, gradients."""
    from phoneme_contrast_b200.models import model_registry
    from phoneme_contrast_b200.training import get_loss_fn
    g = golden.r2
    cfg = {"embedding_dim": 32, "use_attention": True, "dropout_rate": 0.0, "hidden_dims": [16, 32, 32, 64], "use_residual": False}
    sd = nets_oracle.synthetic_state_dict("phoneme_cnn_deep", cfg, seed=23)
    m = model_registry.create("phoneme_cnn_deep", dict(cfg)).to(DEV)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m.train()
    emb = m(cu(g["plain_x"]))
    loss = get_loss_fn("supervised_contrastive", temperature=0.15)(emb, cu(g["plain_y"], torch.int64))
    loss.backward()
    ref = g["plain_emb_train"]
    assert np.abs(emb.detach().cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
    assert abs(float(loss) - float(g["plain_loss"])) <= 1e-4 * abs(float(g["plain_loss"]))
    grads = {n: p.grad.cpu().numpy() for n, p in m.named_parameters()}
    zero_bias = lambda n: n.endswith(".bias") and (n.startswith("init_conv.0") or n.split(".")[-2] in ("0", "3") and n.startswith(("conv_blocks", "projection")))
    for name in g["plain_param_names"]:
        name = str(name)
        if analytically_zero_grad(name) or zero_bias(name):
            continue
        want = float(g[f"plain_gnorm_{name}"])
        assert abs(np.linalg.norm(grads[name].astype(np.float64)) - want) <= GRAD_RTOL * want, name
        key = f"plain_grad_{name}"
        if key in g.files:
            assert rel_l2(grads[name], g[key]) <= GRAD_RTOL, name
    m.eval()
    with torch.no_grad():
        assert np.abs(m(cu(g["plain_x"])).cpu().numpy() - g["plain_emb_eval"]).max() <= 1e-4


def test_quiet_clip_with_gain_vs_reference_golden(golden):
    """A near-silent zero-tailed clip whose mel power falls under amin = 1e-10: the reference clamps AFTER the waveform gain
    (10 log10(max(g^2 mel, 1e-10))); the fused dB-offset epilogue must agree on both sides of g = 1."""
    from phoneme_contrast_b200.datasets import MFCCExtractor, pack_view_descs
    from phoneme_contrast_b200.datasets.transforms import blank_view_descs
    g = golden.r2
    w = cu(g["quiet_wave"])
    ext = MFCCExtractor()
    for tag in ("up", "down"):
        recs = blank_view_descs(1)
        recs["gain"] = np.float32(float(g[f"quiet_gain_{tag}"]))
        out = ext.forward_views(w, pack_view_descs(recs, DEV), 1).cpu().numpy()
        ref = g[f"quiet_mfcc_{tag}"]
        assert np.abs(out - ref).max() <= 1e-4 * np.abs(ref).max(), (tag, np.abs(out - ref).max())


def test_preemphasis_vs_reference_golden(golden):
    """Optional pre-emphasis (off by default = the reference): torchaudio.functional.preemphasis -> reference MFCCExtractor."""
    from phoneme_contrast_b200.datasets import MFCCExtractor
    g = golden.r2
    out = MFCCExtractor(preemphasis=float(g["pre_coeff"]))(cu(g["pre_wave"]), clamp_scope="clip").cpu().numpy()
    ref = g["pre_mfcc"]
    assert rel_l2(out, ref) <= 1e-4 and np.abs(out - ref).max() <= 1e-4 * np.abs(ref).max() * 5
    want = mfcc_oracle.mfcc(g["pre_wave"], clamp_scope="clip", preemph=float(g["pre_coeff"]))
    assert rel_l2(out, want) <= 1e-4
    base = MFCCExtractor()(cu(g["pre_wave"]), clamp_scope="clip").cpu().numpy()
    assert rel_l2(base, ref) > 1e-2          # the coefficient really changes the features


def test_reference_written_checkpoint_round_trip(golden, tmp_path):
    """A checkpoint_*.pt written by the REFERENCE trainer (trainer.py:231-245, file committed by make_golden_r2.py) loads into
    the drop-in (model + Adam state + scheduler + counters), reproduces the reference's eval embeddings, and a checkpoint
    written by the drop-in has the same key structure (so scripts/evaluate.py:276-278 can read it)."""
    from phoneme_contrast_b200.models import model_registry
    from phoneme_contrast_b200.training import ContrastiveTrainer, FusedClipAdam, get_loss_fn
    g = golden.r2
    ref_path = os.path.join(HERE, "golden", "ref_checkpoint_tiny.pt")
    ref_ck = torch.load(ref_path, map_location="cpu", weights_only=False)
    cfg = ref_ck["config"]["model"]
    for fused in (True, False):
        m = model_registry.create("phoneme_cnn_deep", dict(cfg)).to(DEV)
        opt = FusedClipAdam(m.parameters(), lr=3e-4, weight_decay=1e-4) if fused else torch.optim.Adam(m.parameters(), lr=3e-4, weight_decay=1e-4)
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10)
        tr = ContrastiveTrainer(m, [], None, get_loss_fn("supervised_contrastive", temperature=0.15), opt, sched, torch.device(DEV),
                                {"model": cfg, "gradient_clip_val": 1.0, "progress": False}, tmp_path / f"o{int(fused)}", logging.getLogger("t"))
        tr.load_checkpoint(ref_path)
        assert (tr.current_epoch, tr.global_step, tr.best_val_loss) == (4, 123, 0.5)
        m.eval()
        with torch.no_grad():
            emb = m(cu(g["ckpt_x"])).cpu().numpy()
        assert np.abs(emb - g["ckpt_emb_eval"]).max() <= 1e-4
        osd = opt.state_dict()
        ref_state = ref_ck["optimizer_state_dict"]["state"]
        assert len(osd["state"]) == len(ref_state)
        for k, st in ref_state.items():
            np.testing.assert_allclose(osd["state"][k]["exp_avg"].cpu().numpy(), st["exp_avg"].numpy(), rtol=0, atol=0)
            assert float(osd["state"][k]["step"]) == float(st["step"]) == 2.0
        tr._save_checkpoint("mine")
        mine = torch.load(tr.checkpoint_dir / "checkpoint_mine.pt", map_location="cpu", weights_only=False)
        assert set(mine) == set(ref_ck)
        assert list(mine["model_state_dict"].keys()) == list(ref_ck["model_state_dict"].keys())
        assert set(mine["optimizer_state_dict"]) == set(ref_ck["optimizer_state_dict"])
        for k, v in ref_ck["model_state_dict"].items():
            assert mine["model_state_dict"][k].shape == v.shape and mine["model_state_dict"][k].dtype == v.dtype, k


# ============================================================================================= advisor regressions
def test_stem_weight_gradient_with_tiny_dy():
    """the stem weight gradient runs on the FP16X2 kernel; with |dy| ~ 1e-8 (late training: the loss is a
    mean over N) unscaled fp16 operands would be subnormal. With the max|dy| operand scale the result keeps 1e-4."""
    from phoneme_contrast_b200 import ops
    B, H, W, Cout, k = 6, 40, 101, 64, 7
    g = ops.conv_geom(B, H, W, 1, Cout, k, 1, 3)
    gen = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(B, H, W, 1, device=DEV, generator=gen)
    dy = torch.randn(B, H, W, Cout, device=DEV, generator=gen) * 1e-8
    amax = dy.abs().max().reshape(1).contiguous()
    dw, db = ops.conv_wgrad(x, dy, g, None, prec=3, dy_amax=amax)
    wr = torch.zeros(Cout, 1, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    br = torch.zeros(Cout, device=DEV, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), wr, br, padding=3).backward(dy.permute(0, 3, 1, 2).double())
    assert float((dw.double() - wr.grad).abs().max() / wr.grad.abs().max()) < 1e-4
    assert float((db.double() - br.grad).abs().max() / br.grad.abs().max()) < 1e-4
    dw_bad, _ = ops.conv_wgrad(x, dy, g, None, prec=3, dy_amax=None)                   # what round 1 did: fp16 subnormal operands
    err_scaled = float((dw.double() - wr.grad).abs().max() / wr.grad.abs().max())
    err_unscaled = float((dw_bad.double() - wr.grad).abs().max() / wr.grad.abs().max())
    assert err_unscaled > 10 * err_scaled, (err_scaled, err_unscaled)


def test_f16_overflow_flag_and_trainer_fallback():
    """Activations beyond fp16's range raise the device flag in every plane writer; the trainer's per-epoch check switches the
    model to tf32x3 (or raises with config f16_overflow='raise')."""
    from phoneme_contrast_b200 import _lib as L
    from phoneme_contrast_b200 import ops
    L.f16_overflow(reset=True)
    y = torch.randn(2, 8, 8, 64, device=DEV)
    ops.bn_act_split(y)
    assert not L.f16_overflow(reset=True)
    y[1, 3, 3, 7] = 7.0e4
    ops.bn_act_split(y)
    assert L.f16_overflow(reset=False) and L.f16_overflow(reset=True) and not L.f16_overflow(reset=True)
    case = _bench_case("phoneme_cnn")
    m, tr = _trainer("phoneme_cnn", case, False)
    y[0, 0, 0, 0] = float("nan")
    ops.bn_act_split(y)
    tr._check_f16_range()
    assert m._prec == L.PREC_TF32X3
    ops.bn_act_split(y)
    tr.config["f16_overflow"] = "raise"
    with pytest.raises(FloatingPointError):
        tr._check_f16_range()
    assert not L.f16_overflow(reset=True)


# ============================================================================================= dataset path
def _golden_dataset(g, device_frontend, noise_source="torch_cpu"):
    import random

    from phoneme_contrast_b200.datasets import MFCCExtractor, PhonemeContrastiveDataset, build_augmentation_pipeline
    aug = {"time_mask": {"enabled": True, "max_width": 30, "prob": 0.5}, "freq_mask": {"enabled": True, "max_width": 10, "prob": 0.5},
           "noise": {"enabled": True, "min_snr": 0.001, "max_snr": 0.005, "prob": 0.5}}
    raw = [torch.from_numpy(g[f"ds_raw_{i}"]) for i in range(4)]
    ds = PhonemeContrastiveDataset(list(range(4)), [3, 1, 4, 1], [{} for _ in raw], MFCCExtractor(), build_augmentation_pipeline(aug, noise_source),
                                   {"target_sr": 16000, "max_length_ms": 500, "contrastive": {"views_per_sample": 2}}, mode="train",
                                   device=DEV, device_frontend=device_frontend, waveforms=raw)
    for i, seed in enumerate(g["ds_pad_seeds"]):
        random.seed(int(seed))
        ds._load_waveform(i)                           # pad / crop decisions drawn as the reference drew them
    return ds


def test_dataset_getitem_vs_reference_golden(golden):
    """PhonemeContrastiveDataset.__getitem__ (reference dataset.py:65-111) on injected waveforms: cached fixed-length clips are
    bit-identical, both views of every item match the reference's features (noise drawn from the torch CPU generator)."""
    g = golden.r2
    ds = _golden_dataset(g, device_frontend=False)
    assert len(ds) == 4 and ds.max_samples == 8000 and ds.n_views == 2 and ds.use_cache
    for i in range(4):
        assert np.array_equal(ds.waveform_cache[i].numpy(), g[f"ds_fixed_{i}"])
        item = ds[i]
        assert set(item) == {"views", "label", "metadata", "index"} and item["label"] == int(g[f"ds_label_{i}"]) and item["index"] == i
        ref = g[f"ds_views_{i}"]
        got = item["views"].numpy()
        assert got.shape == ref.shape == (2, 1, 40, 51)
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max() * 5, (i, np.abs(got - ref).max())
        for v in range(2):      # mask cells: exact zeros only where the reference has them, unless noise was added on top
            if np.any(ref[v] == 0.0):
                assert np.array_equal(got[v] == 0.0, ref[v] == 0.0)


def test_device_resident_batch_views_vs_reference_golden(golden):
    """device_frontend=True: waveforms cached on the GPU, one fused launch per batch, clips addressed through the descriptors
    (no gather) == the reference's per-item loop; and the loader feeds ContrastiveTrainer._prepare_batch unchanged."""
    from phoneme_contrast_b200.datasets import DeviceFrontendLoader, build_view_descriptors
    g = golden.r2
    ds = _golden_dataset(g, device_frontend=True)
    assert set(ds[2]) == {"label", "metadata", "index"}
    order = [2, 0, 3, 1]
    _, noise = build_view_descriptors(order, 2, 40, 51, ds.augmentation_pipeline, want_noise=True)
    views, labels = ds.batch_views(order, noise=noise.to(DEV))
    assert views.shape == (4, 2, 1, 40, 51) and labels.tolist() == [4, 3, 1, 1]
    got = views.cpu().numpy()
    for k, i in enumerate(order):
        ref = g[f"ds_views_{i}"]
        assert np.abs(got[k] - ref).max() <= 1e-4 * np.abs(ref).max() * 5, (i, np.abs(got[k] - ref).max())
    # production flavour: device Philox noise; loader + trainer plumbing
    ds2 = _golden_dataset(g, device_frontend=True, noise_source="device")
    loader = DeviceFrontendLoader(ds2, [[0, 1], [2, 3]])
    assert len(loader) == 2
    case = _bench_case("phoneme_cnn")
    m, tr = _trainer("phoneme_cnn", case, False)
    m._inject_drop = None
    for batch in loader:
        v, y = tr._prepare_batch(batch)
        assert v.shape == (4, 1, 40, 51) and v.is_cuda and y.shape == (4,)
        assert np.isfinite(float(tr.step(v, y)))


# ============================================================================================= pooled-resolution stem backward
def test_stem_gram_matches_im2col():
    """G = X^T X and X1 = X^T 1 of the 7x7 'same' patches (csrc/stem_bwd.cu: lag correlations + exact border handling) against
    an explicit im2col in fp64."""
    from phoneme_contrast_b200 import ops
    for (B, H, W) in ((3, 40, 101), (2, 9, 11), (5, 40, 51)):
        x = torch.randn(B, 1, H, W, device=DEV, generator=torch.Generator(device=DEV).manual_seed(B + H))
        buf = ops.stem_gram(x).cpu().numpy()
        G, X1 = buf[:49 * 49].reshape(49, 49), buf[49 * 49:]
        cols = torch.nn.functional.unfold(x.double().cpu(), kernel_size=7, padding=3)      # [B, 49, H*W]
        X = cols.permute(0, 2, 1).reshape(-1, 49).numpy()
        Gref, X1ref = X.T @ X, X.sum(0)
        assert np.abs(G - Gref).max() <= 2e-6 * np.abs(Gref).max(), np.abs(G - Gref).max() / np.abs(Gref).max()
        assert np.abs(X1 - X1ref).max() <= 1e-5 * max(np.abs(X1ref).max(), 1.0)


@pytest.mark.parametrize("B,shape", [(8, (40, 101)), (6, (40, 64)), (64, (40, 101))])
def test_stem_backward_pooled_vs_full_resolution_and_oracle(B, shape, monkeypatch):
    """init_conv gradients (conv weight, BatchNorm weight / bias) from the pooled-resolution closed form == the round-1 path that
    streams the pre-BatchNorm tensor three times, and both match the oracle."""
    from tests.test_gpu_parity import _oracle_net, _run_net
    arch, cfg = "phoneme_cnn_deep", {"dropout_rate": 0.0}
    sd = nets_oracle.synthetic_state_dict(arch, cfg, seed=13)
    rs = np.random.RandomState(B)
    x = rs.standard_normal((B, 1) + shape).astype(np.float32)
    y = (np.arange(B) // 2).astype(np.int64)
    monkeypatch.setenv("PC_STEM_BWD", "0")
    _, emb0, loss0, g0 = _run_net(arch, cfg, sd, x, y)
    monkeypatch.setenv("PC_STEM_BWD", "1")
    _, emb1, loss1, g1 = _run_net(arch, cfg, sd, x, y)
    assert abs(loss0 - loss1) <= 1e-5 * abs(loss0)      # two runs differ in the last bits of the atomically accumulated BatchNorm statistics
    _, _, gref, _ = _oracle_net(arch, cfg, sd, x, y)
    for name in ("init_conv.0.weight", "init_conv.1.weight", "init_conv.1.bias"):
        a, b, r = g1[name].astype(np.float64), g0[name].astype(np.float64), gref[name].astype(np.float64)
        assert np.linalg.norm(a - b) <= GRAD_RTOL * np.linalg.norm(b), (name, "vs full-resolution path", np.linalg.norm(a - b) / np.linalg.norm(b))
        assert np.linalg.norm(a - r) <= GRAD_RTOL * np.linalg.norm(r), (name, "vs oracle", np.linalg.norm(a - r) / np.linalg.norm(r))
    assert np.abs(g1["init_conv.0.bias"]).max() == 0.0          # analytically zero, written as such


@pytest.mark.parametrize("B,shape,training", [(4, (40, 101), True), (3, (40, 64), True), (2, (9, 17), True), (5, (40, 101), False), (40, (40, 101), True)])
def test_stem_forward_one_pass_vs_two_pass(B, shape, training, monkeypatch):
    """Fused stem (statistics from the Gram matrix, conv + BatchNorm + ReLU + max-pool in one tcgen05 kernel, no y0) == the
    two-pass path (SIMT conv + statistics, then normalise + pool): pooled output to 1e-5 of its max, argmax positions identical
    except at ties within rounding, planes consistent, running statistics equal."""
    from phoneme_contrast_b200 import _lib as L
    from phoneme_contrast_b200 import ops
    H, W = shape
    gen = torch.Generator(device=DEV).manual_seed(B + W)
    x = torch.randn(B, 1, H, W, device=DEV, generator=gen)
    outs = []
    for fused in (False, True):
        torch.manual_seed(3)
        conv = torch.nn.Conv2d(1, 64, 7, padding=3).to(DEV)
        bn = torch.nn.BatchNorm2d(64).to(DEV)
        with torch.no_grad():
            bn.weight.copy_(1.0 + 0.1 * torch.randn(64, device=DEV, generator=gen)); bn.bias.copy_(0.1 * torch.randn(64, device=DEV, generator=gen))
            bn.running_mean.copy_(0.05 * torch.randn(64, device=DEV, generator=gen)); bn.running_var.copy_(1.0 + 0.1 * torch.rand(64, device=DEV, generator=gen))
        gen = torch.Generator(device=DEV).manual_seed(B + W)           # same draws for both variants
        x = torch.randn(B, 1, H, W, device=DEV, generator=gen)
        st = torch.zeros(2, 64, device=DEV, dtype=torch.float64)
        if fused:
            if training:
                gram = ops.stem_gram(x)
                ops.stem_stats_from_gram(gram, conv, B, H, W, st)
            co = ops.bn_finalize(st if training else None, B * H * W, bn, training)
            p0, am, planes = ops.stem_fwd(x, conv, co, want_planes=True)
        else:
            g = ops.conv_geom(B, H, W, 1, 64, 7, 1, 3)
            y0 = ops.conv_fwd(x, conv.weight, conv.bias, g, None, st if training else None, L.PREC_FP32)
            co = ops.bn_finalize(st if training else None, B * H * W, bn, training)
            p0, am, planes = ops.bn_act_fwd(y0, co, 3, None, want_planes=True)
        outs.append((p0, am, planes, st.clone(), bn.running_mean.clone(), bn.running_var.clone()))
    (p_a, am_a, pl_a, st_a, rm_a, rv_a), (p_b, am_b, pl_b, st_b, rm_b, rv_b) = outs
    scale = float(p_a.abs().max())
    assert float((p_a - p_b).abs().max()) <= 1e-5 * scale, float((p_a - p_b).abs().max()) / scale
    if training:
        np.testing.assert_allclose(st_b.cpu().numpy(), st_a.cpu().numpy(), rtol=2e-6, atol=1e-3)
    np.testing.assert_allclose(rm_b.cpu().numpy(), rm_a.cpu().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(rv_b.cpu().numpy(), rv_a.cpu().numpy(), rtol=1e-5, atol=1e-7)
    differ = (am_a != am_b)
    assert float(differ.float().mean()) <= 1e-4                      # ties / near-ties inside a window only
    assert float((p_a - p_b).abs()[differ].max()) <= 1e-5 * scale if differ.any() else True
    halves = pl_b.view(torch.float16).reshape(-1)
    hi = halves[: p_b.numel()].float().view_as(p_b)
    lo = halves[p_b.numel():].float().view_as(p_b)
    assert float((hi + lo / 2048.0 - p_b).abs().max()) <= 1e-6 * scale


@pytest.mark.parametrize("shape", [(64, 10, 25, 128), (16, 3, 7, 512), (8, 5, 13, 256), (4, 3, 5, 96)])
@pytest.mark.parametrize("attention", [True, False])
def test_attn_pool_single_pass_vs_torch(shape, attention):
    """SpatialAttention gate + global average pool (reference phoneme_cnn.py:129-143, :117-119): the single-pass warp-per-pixel kernels
    (C = 128 / 256 / 512) and the general kernels (C = 96) against torch autograd on the GPU."""
    from phoneme_contrast_b200 import ops
    B, H, W, C_ = shape
    g = torch.Generator().manual_seed(B + C_)
    a = torch.randn(B, H, W, C_, generator=g).cuda()
    w = (torch.randn(C_, generator=g) * 0.2).cuda() if attention else None
    b0 = torch.randn(1, generator=g).cuda() if attention else None
    dp = torch.randn(B, C_, generator=g).cuda()
    pooled, gate = ops.attn_pool_fwd(a, w, b0)
    da, dw, db0 = ops.attn_pool_bwd(a, gate, dp, w)
    ar = a.clone().requires_grad_(True)
    if attention:
        wr, br = w.clone().requires_grad_(True), b0.clone().requires_grad_(True)
        gr = torch.sigmoid((ar * wr).sum(-1) + br)
        want = (ar * gr.unsqueeze(-1)).mean((1, 2))
    else:
        want = ar.mean((1, 2))
    want.backward(dp)
    tol = dict(rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(pooled, want.detach(), **tol)
    torch.testing.assert_close(da, ar.grad, **tol)
    if attention:
        torch.testing.assert_close(gate, gr.detach().reshape(B, H * W), **tol)
        torch.testing.assert_close(dw, wr.grad, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(db0, br.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("graph", [False, True])
def test_run_batches_delivers_every_loss_in_order(graph):
    """ContrastiveTrainer.run_batches (the epoch loop of reference trainer.py:138-160): host batches are prefetched one ahead, the loss is
    accumulated on the device, and the on_loss callback receives EVERY step's loss, in order, one step behind -- the same values a loop
    that reads loss.item() after each step (as the reference does) sees."""
    from phoneme_contrast_b200.models import model_registry
    from phoneme_contrast_b200.training import ContrastiveTrainer, FusedClipAdam, get_loss_fn
    cfg = {"dropout_rate": 0.0, "precision": "fp32"}
    sd = nets_oracle.synthetic_state_dict("phoneme_cnn", cfg, seed=9)
    rs = np.random.RandomState(12)
    batches = [{"views": torch.from_numpy(rs.standard_normal((16, 2, 1, 40, 64)).astype(np.float32)).pin_memory(),
                "label": torch.from_numpy((np.arange(16) // 2).astype(np.int64)).pin_memory()} for _ in range(5)]
    got = {}
    for mode in ("item", "pipelined"):
        m = model_registry.create("phoneme_cnn", cfg).to(DEV)
        m.load_state_dict(sd)
        opt = FusedClipAdam(m.parameters(), lr=1e-3, weight_decay=1e-4)
        tr = ContrastiveTrainer(m, [], None, get_loss_fn("supervised_contrastive", temperature=0.15), opt, None, torch.device(DEV),
                                {"gradient_clip_val": 1.0, "cuda_graph": graph, "progress": False}, tempfile.mkdtemp(), logging.getLogger("t"))
        m.train()
        if mode == "item":
            got[mode] = [float(tr.step(*tr._prepare_batch(b)).item()) for b in batches]
        else:
            seen = []
            total, n = tr.run_batches(iter(batches), on_loss=lambda i, v: seen.append((i, v)))
            assert n == len(batches) and [i for i, _ in seen] == list(range(len(batches))) and tr.global_step == len(batches)
            got[mode] = [v for _, v in seen]
            assert abs(float(total) - sum(got[mode])) <= 1e-4 * abs(sum(got[mode]))
    np.testing.assert_allclose(got["pipelined"], got["item"], rtol=2e-5)
