"""Oracle: view augmentation decisions (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates, call for call, the host RNG traffic of the reference so that every decision and
index is bit-exact:

  src/datasets/dataset.py:147-172       _augment_waveform  (gain, seed idx*10000+view)
  src/datasets/transforms.py:38-46      TimeMask.__call__      (seed s + 0)
  src/datasets/transforms.py:62-70      FrequencyMask.__call__ (seed s + 1000)
  src/datasets/transforms.py:87-97      GaussianNoise.__call__ (seed s + 2000)
  src/datasets/transforms.py:139-144    Compose (seed + i*1000), s = idx*20000+view (dataset.py:94)
  torchaudio/functional/functional.py:885-958  mask_along_axis (two torch.rand(1) draws, fp32)

The straightforward per-call restatement here is deliberately independent of the vectorised
descriptor builder in phoneme_contrast_b200/datasets/transforms.py.
"""
from __future__ import annotations

import random

import numpy as np
import torch


def gain_decision(seed: int, prob: float = 0.5, lo: float = 0.8, hi: float = 1.2):
    """dataset.py:160-167 -> (applied, gain)."""
    seed = int(seed)
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    if random.random() < prob:
        return True, random.uniform(lo, hi)
    return False, 1.0


def axis_mask_decision(seed: int | None, prob: float, max_width: int, axis_size: int):
    """transforms.py:38-46 / :62-70 + functional.py:932-947 -> (applied, start, end).

    All arithmetic after torch.rand(1) is float32, as in torchaudio; .long() truncates.
    """
    if seed is not None:
        seed = int(seed)
        random.seed(seed)
        torch.manual_seed(seed)
    if not (random.random() < prob):
        return False, 0, 0
    if max_width < 1:
        return True, 0, 0
    value = torch.rand(1) * max_width
    min_value = torch.rand(1) * (axis_size - value)
    start = int(min_value.long())
    end = int(min_value.long() + value.long())
    return True, start, end


def noise_decision(seed: int | None, prob: float, min_snr: float, max_snr: float, shape=None):
    """transforms.py:87-97 -> (applied, level, noise or None). noise = torch.randn(shape) drawn from the
    CPU generator right after torch.manual_seed(seed) (what randn_like(x) consumes)."""
    if seed is not None:
        seed = int(seed)
        random.seed(seed)
        torch.manual_seed(seed)
    if not (random.random() < prob):
        return False, 0.0, None
    level = random.uniform(min_snr, max_snr)
    noise = torch.randn(shape) if shape is not None else None
    return True, level, noise


DEFAULT_AUG = {
    "time_mask": {"enabled": True, "max_width": 30, "prob": 0.5},
    "freq_mask": {"enabled": True, "max_width": 10, "prob": 0.5},
    "noise": {"enabled": True, "min_snr": 0.001, "max_snr": 0.005, "prob": 0.3},
}


def view_descriptor(idx: int, view: int, n_freq: int, n_time: int, cfg: dict | None = None,
                    waveform_gain: bool = True, noise_shape=None) -> dict:
    """Everything random about view `view` of dataset item `idx` (dataset.py:79-98)."""
    cfg = DEFAULT_AUG if cfg is None else cfg
    d = {"gain": 1.0, "gain_applied": False, "t": (False, 0, 0), "f": (False, 0, 0),
         "noise": (False, 0.0), "noise_tensor": None}
    if waveform_gain:
        d["gain_applied"], d["gain"] = gain_decision(idx * 10000 + view)
    seed = idx * 20000 + view
    i = 0
    if cfg.get("time_mask", {}).get("enabled", False):
        p = cfg["time_mask"]
        d["t"] = axis_mask_decision(seed + i * 1000, p.get("prob", 0.5), p.get("max_width", 30), n_time)
        i += 1
    if cfg.get("freq_mask", {}).get("enabled", False):
        p = cfg["freq_mask"]
        d["f"] = axis_mask_decision(seed + i * 1000, p.get("prob", 0.5), p.get("max_width", 10), n_freq)
        i += 1
    if cfg.get("noise", {}).get("enabled", False):
        p = cfg["noise"]
        a, lvl, nz = noise_decision(seed + i * 1000, p.get("prob", 0.3), p.get("min_snr", 0.001),
                                    p.get("max_snr", 0.005), noise_shape)
        d["noise"] = (a, lvl)
        d["noise_tensor"] = nz
    return d


def apply_view(feats: np.ndarray, d: dict) -> np.ndarray:
    """Apply masks then noise to feats [F,T] in the reference's order (Compose: time, freq, noise).
    Masks zero-fill [start,end) (functional.py:941-952); noise is x + noise*level (transforms.py:95-96)."""
    out = np.array(feats, dtype=np.float64, copy=True)
    a, s, e = d["t"]
    if a:
        out[:, s:e] = 0.0
    a, s, e = d["f"]
    if a:
        out[s:e, :] = 0.0
    a, lvl = d["noise"]
    if a and d.get("noise_tensor") is not None:
        out = out + np.asarray(d["noise_tensor"], dtype=np.float64).reshape(out.shape) * np.float32(lvl).astype(np.float64)
    return out
