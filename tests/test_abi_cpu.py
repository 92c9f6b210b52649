"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/phoneme_contrast.h declares; the host-side mirrors keep the reference's interface and error behaviour;
descriptor tables are bit-exact against the reference goldens. No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from phoneme_contrast_b200 import build
    return build.build()


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "phoneme_contrast.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/phoneme_contrast.h but not exported"
    lib.pc_abi_version.restype = ctypes.c_int
    assert lib.pc_abi_version() == 1


def test_ctypes_signatures_cover_header(built_lib):
    from phoneme_contrast_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_symbols()
    assert _lib.lib() is not None
    assert ctypes.sizeof(_lib.PcViewDesc) == 32
    assert ctypes.sizeof(_lib.PcConvGeom) == 44


def test_missing_library_fails_loudly(monkeypatch):
    from phoneme_contrast_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpc_b200.so")
    with pytest.raises(_lib.NativeLibraryMissing):
        _lib.lib()


def test_no_cpu_fallback():
    from phoneme_contrast_b200.datasets import MFCCExtractor, TimeMask
    from phoneme_contrast_b200.models import model_registry
    from phoneme_contrast_b200.training import SupervisedContrastiveLoss
    with pytest.raises(RuntimeError):
        model_registry.create("phoneme_cnn", {})(torch.randn(2, 1, 40, 50))
    with pytest.raises(RuntimeError):
        MFCCExtractor()(torch.randn(1, 4000))
    with pytest.raises(RuntimeError):
        SupervisedContrastiveLoss()(torch.randn(4, 8), torch.tensor([0, 0, 1, 1]))
    with pytest.raises(RuntimeError):
        TimeMask(prob=1.0)(torch.randn(1, 1, 40, 100), seed=42)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "phoneme_contrast_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("oracle/mfcc_oracle.py for the line-by-line", ""), f"{f} references oracle/"


# ----------------------------------------------------------------------------- registry / factories (reference tests/test_models.py:8-30)
def test_registry_interface():
    from phoneme_contrast_b200.models import BaseModel, model_registry
    assert "phoneme_cnn" in model_registry.list() and "phoneme_cnn_deep" in model_registry.list()
    assert issubclass(model_registry.get("phoneme_cnn"), BaseModel)
    m = model_registry.create("phoneme_cnn", {"embedding_dim": 64})
    assert m.get_embedding_dim() == 64 and len(m.conv_blocks) == 3
    with pytest.raises(ValueError):
        model_registry.create("nope", {})
    with pytest.raises(ValueError):
        model_registry.register("phoneme_cnn")(type(m))


def test_state_dict_keys_and_param_counts(golden):
    from phoneme_contrast_b200.models import model_registry
    g = golden.nets
    for arch in ("phoneme_cnn", "phoneme_cnn_deep"):
        m = model_registry.create(arch, {})
        assert list(m.state_dict().keys()) == [str(k) for k in g[f"{arch}_keys"]]
        assert sum(p.numel() for p in m.parameters()) == int(g[f"{arch}_n_params"])


def test_loss_and_extractor_factories():
    from phoneme_contrast_b200.datasets import MFCCExtractor, MelSpectrogramExtractor, build_feature_extractor
    from phoneme_contrast_b200.training import SupervisedContrastiveLoss, get_loss_fn
    lf = get_loss_fn("supervised_contrastive", temperature=0.15)
    assert isinstance(lf, SupervisedContrastiveLoss) and lf.base_temperature == 0.07
    with pytest.raises(ValueError):
        get_loss_fn("nope")
    assert isinstance(build_feature_extractor({}), MFCCExtractor)
    assert isinstance(build_feature_extractor({"type": "mel"}), MelSpectrogramExtractor)
    with pytest.raises(ValueError):
        build_feature_extractor({"type": "nope"})
    with pytest.raises(ValueError):
        MFCCExtractor(n_mfcc=100, n_mels=80)


def test_frontend_constants_bit_identical_to_reference(golden):
    from phoneme_contrast_b200.datasets.features import _FrontEndConsts
    g = golden.mfcc
    c = _FrontEndConsts(16000, 400, 160, 80, 0.0, 8000.0, 40)
    assert np.array_equal(c.window.numpy(), g["window"])
    assert np.array_equal(c.fb.numpy(), g["fb"])
    assert np.array_equal(c.dct.numpy(), g["dct"])
    # banded form reproduces the dense filterbank exactly
    dense = np.zeros_like(g["fb"])
    for m in range(80):
        s, n = int(c.start[m]), int(c.length[m])
        dense[s:s + n, m] = c.w[m, :n].numpy()
    assert np.array_equal(dense, g["fb"])


# ----------------------------------------------------------------------------- augmentation decisions
AUG_CFG = {"time_mask": {"enabled": True, "max_width": 30, "prob": 0.5},
           "freq_mask": {"enabled": True, "max_width": 10, "prob": 0.5},
           "noise": {"enabled": True, "min_snr": 0.001, "max_snr": 0.005, "prob": 0.3}}


def test_view_descriptors_bit_exact_vs_reference(golden):
    from phoneme_contrast_b200.datasets import build_augmentation_pipeline, build_view_descriptors
    rec = golden.augment["rec"]
    pipe = build_augmentation_pipeline(AUG_CFG)
    assert [type(t).__name__ for t in pipe.transforms] == ["TimeMask", "FrequencyMask", "GaussianNoise"]
    recs, _ = build_view_descriptors(range(rec.shape[0]), 2, 40, 101, pipe)
    r = recs.reshape(rec.shape[0], 2)
    for i in range(rec.shape[0]):
        for v in range(2):
            a, b = r[i, v], rec[i, v]
            t = (int(a["t0"]), int(a["t1"])) if a["t1"] > a["t0"] else (0, 0)
            f = (int(a["f0"]), int(a["f1"])) if a["f1"] > a["f0"] else (0, 0)
            assert a["gain"] == np.float32(b[0])
            assert t == (int(b[1]), int(b[2])) and f == (int(b[3]), int(b[4]))
            assert a["noise_level"] == np.float32(b[6])
            assert a["clip"] == i


def test_view_descriptors_match_oracle_on_other_seeds():
    from oracle import augment_oracle
    from phoneme_contrast_b200.datasets import build_augmentation_pipeline, build_view_descriptors
    pipe = build_augmentation_pipeline(AUG_CFG)
    idx = [1000, 54321, 7, 99999]
    recs, _ = build_view_descriptors(idx, 2, 40, 201, pipe)
    k = 0
    for i in idx:
        for v in range(2):
            d = augment_oracle.view_descriptor(i, v, 40, 201)
            a = recs[k]
            k += 1
            assert a["gain"] == np.float32(d["gain"])
            assert (int(a["t0"]), int(a["t1"])) == ((d["t"][1], d["t"][2]) if d["t"][0] else (0, 0))
            assert (int(a["f0"]), int(a["f1"])) == ((d["f"][1], d["f"][2]) if d["f"][0] else (0, 0))
            assert a["noise_level"] == np.float32(d["noise"][1])


def test_pipeline_factory(golden):
    from phoneme_contrast_b200.datasets import FrequencyMask, TimeMask, build_augmentation_pipeline
    cfg = {"time_mask": {"enabled": True, "max_width": 20}, "freq_mask": {"enabled": True, "max_width": 5},
           "noise": {"enabled": False}}
    pipe = build_augmentation_pipeline(cfg)      # reference tests/test_transforms.py:90-103
    assert len(pipe.transforms) == 2
    assert isinstance(pipe.transforms[0], TimeMask) and isinstance(pipe.transforms[1], FrequencyMask)


def test_zero_pool_slices_are_zero_aligned_and_disjoint():
    """ops.ZeroPool hands out dtype views of one zeroed buffer (host logic, runs on CPU tensors too)."""
    import torch
    from phoneme_contrast_b200 import ops
    zp = ops.ZeroPool("cpu", nbytes=4096)
    a = zp.take((2, 64), torch.float64)
    b = zp.take((3,), torch.float32)
    c = zp.take((2, 32), torch.float64)
    assert a.dtype == torch.float64 and a.shape == (2, 64) and b.shape == (3,) and c.shape == (2, 32)
    assert float(a.abs().sum()) == 0.0 and float(b.abs().sum()) == 0.0
    a.fill_(1.0); b.fill_(2.0); c.fill_(3.0)
    assert float(a.sum()) == 128.0 and float(b.sum()) == 6.0 and float(c.sum()) == 192.0      # no overlap
    for t in (a, b, c):
        assert t.data_ptr() % 16 == 0
    big = zp.take((4096,), torch.float64)            # does not fit any more: falls back to a fresh zero tensor
    assert big.shape == (4096,) and float(big.abs().sum()) == 0.0


def test_fastdiv_formula_matches_integer_division():
    """Mirror of common.cuh:FastDiv (multiply-shift division used by the weight-gradient producers)."""
    import random

    def make(d):
        l = 0
        while (1 << l) < d:
            l += 1
        return (((1 << 32) * ((1 << l) - d)) // d + 1) & 0xFFFFFFFF, min(l, 1), max(l - 1, 0)

    rnd = random.Random(0)
    for d in [1, 2, 3, 5, 7, 13, 20, 26, 40, 51, 101, 1020, 4040, 65535, 1000003, 2**31 - 1]:
        mul, s1, s2 = make(d)
        for n in [0, 1, d - 1, d, d + 1, 2**31 - 1, 2**32 - 1] + [rnd.randrange(2**32) for _ in range(2000)]:
            t = (mul * n) >> 32
            q = ((t + ((n - t) >> s1)) & 0xFFFFFFFF) >> s2
            assert q == n // d, (d, n)


def test_dataset_pad_or_trim_matches_reference_golden(golden):
    """Host logic of the dataset mirror (reference dataset.py:174-203): train-mode random crop / left pad drawn from Python's
    global RNG with the same calls, centred crop / pad in val mode; bit-identical to the reference's outputs."""
    import random

    import torch

    from phoneme_contrast_b200.datasets.dataset import PhonemeContrastiveDataset
    g = golden.r2
    raw = [torch.from_numpy(g[f"ds_raw_{i}"]) for i in range(4)]
    cfg = {"target_sr": 16000, "max_length_ms": 500, "contrastive": {"views_per_sample": 2}}
    ds = PhonemeContrastiveDataset(list(range(4)), [3, 1, 4, 1], [{}] * 4, None, None, cfg, mode="train", device="cpu", waveforms=raw)
    assert ds.max_samples == 8000 and ds.n_views == 2
    for i, seed in enumerate(g["ds_pad_seeds"]):
        random.seed(int(seed))
        assert np.array_equal(ds._load_waveform(i).numpy(), g[f"ds_fixed_{i}"])
    val = PhonemeContrastiveDataset(list(range(4)), [3, 1, 4, 1], [{}] * 4, None, None, cfg, mode="val", device="cpu", waveforms=raw)
    assert val.n_views == 1
    for i in range(4):
        assert np.array_equal(val._load_waveform(i).numpy(), g[f"ds_val_fixed_{i}"])


def test_exchange_packing_round_trip_cpu():
    """Host restatement of csrc/dp.cu's packing (tests/exchange_torch.py, used by the gloo tests): int64 labels survive the
    trip through two fp32 columns bit for bit, including negative and > 2^31 values; loss_from_stats equals the mean row loss."""
    import torch

    from oracle import supcon_oracle
    from tests.exchange_torch import TorchExchangeMixin as X
    emb = torch.randn(7, 12)
    lab = torch.tensor([0, 1, -5, 2 ** 31 + 3, 2 ** 40 + 17, -2 ** 62, 37], dtype=torch.int64)
    F, y = X.unpack(X.pack(emb, lab), 12)
    assert torch.equal(F, emb) and torch.equal(y, lab)
    rs = np.random.RandomState(3)
    f = rs.standard_normal((24, 16))
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    yy = rs.randint(0, 4, 24)
    st = supcon_oracle.row_stats(f, yy, temperature=0.15)
    stats = torch.from_numpy(np.stack([st["m"], st["den"], st["npos"], st["spos"]], 1))
    got = float(X.loss_from_stats(stats, 0.15, 0.07))
    want = supcon_oracle.loss(f, yy, temperature=0.15)
    assert abs(got - want) <= 1e-9 * abs(want)
