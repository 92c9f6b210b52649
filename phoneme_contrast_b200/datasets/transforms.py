"""Drop-in for src/datasets/transforms.py: BaseTransform / TimeMask / FrequencyMask / GaussianNoise /
TimeStretch / Compose / build_augmentation_pipeline with the reference's constructor arguments, call
signature `t(x, seed=None)` and -- bit for bit -- its random decisions.

Every decision (apply or skip, mask start/end, noise level, waveform gain) is drawn on the host with exactly
the RNG calls the reference makes (Python `random` + `torch.manual_seed`/`torch.rand(1)`, transforms.py:38-46,
62-70, 87-97; torchaudio mask_along_axis functional.py:932-947; dataset.py:160-167), so indices are
bit-exact and the global RNG state evolves identically. The tensor work (zero-fill, noise add) is a CUDA
kernel (pc_augment_apply), or is fused into the MFCC kernel's epilogue via a PcViewDesc table.
"""
from __future__ import annotations

import random
from typing import List, Optional, Sequence

import numpy as np
import torch

from .._lib import call, ptr, stream

VIEW_DESC_DTYPE = np.dtype([("gain", "<f4"), ("t0", "<i4"), ("t1", "<i4"), ("f0", "<i4"), ("f1", "<i4"),
                            ("noise_level", "<f4"), ("noise_seed", "<u4"), ("clip", "<i4")])   # == struct PcViewDesc
assert VIEW_DESC_DTYPE.itemsize == 32


def _seed_host(seed) -> None:
    seed = int(seed)
    random.seed(seed)
    torch.manual_seed(seed)


def _mask_interval(max_width: int, axis_size: int):
    """torchaudio mask_along_axis with p=1.0: two fp32 uniform draws from the torch CPU generator."""
    if max_width < 1:
        return 0, 0
    value = torch.rand(1) * max_width
    min_value = torch.rand(1) * (axis_size - value)
    start = int(min_value.long())
    end = int(min_value.long() + value.long())
    if end - start >= max_width:
        raise ValueError("Number of columns to be masked should be less than mask_param")   # functional.py:949-950
    return start, end


def _apply(x: torch.Tensor, rec: np.ndarray, noise: Optional[torch.Tensor]) -> torch.Tensor:
    """Apply one descriptor to every [F,T] plane of x (torchaudio applies one mask to all leading dims)."""
    if x.dim() < 2:
        raise ValueError(f"Spectrogram must have at least two dimensions (time and frequency) ({x.dim()} given).")
    if not x.is_cuda:
        raise RuntimeError("phoneme_contrast_b200 transforms run on CUDA tensors only (no CPU fallback)")
    xc = x.to(torch.float32).contiguous()
    F_, T = xc.shape[-2], xc.shape[-1]
    n = xc.numel() // (F_ * T)
    recs = np.repeat(rec, n)
    views = torch.from_numpy(recs.view(np.uint8)).to(x.device)
    out = torch.empty_like(xc)
    call("pc_augment_apply", ptr(xc), ptr(views, torch.uint8), n, F_, T, ptr(noise), ptr(out), stream())
    return out


def _blank(n=1) -> np.ndarray:
    rec = np.zeros(n, dtype=VIEW_DESC_DTYPE)
    rec["gain"] = 1.0
    return rec


blank_view_descs = _blank      # n identity descriptors (gain 1, no masks, no noise): fill in fields, then pack_view_descs


class BaseTransform:
    """Base class for augmentations (transforms.py:9-22)."""

    def __call__(self, x: torch.Tensor, seed: Optional[int] = None) -> torch.Tensor:
        raise NotImplementedError

    def decide(self, rec: np.ndarray, shape) -> Optional[torch.Tensor]:
        """Draw this transform's random decision (consuming host RNG exactly like the reference, assuming the
        caller already seeded) and record it in the one-element descriptor `rec`. Returns explicit noise or None."""
        raise NotImplementedError


class TimeMask(BaseTransform):
    """Zero a random run of time steps (transforms.py:25-46)."""

    def __init__(self, max_width: int = 30, prob: float = 0.5):
        self.max_width = max_width
        self.prob = prob

    def decide(self, rec, shape):
        if random.random() < self.prob:
            rec["t0"], rec["t1"] = _mask_interval(self.max_width, shape[-1])
        return None

    def __call__(self, x, seed=None):
        if seed is not None:
            _seed_host(seed)
        rec = _blank()
        self.decide(rec, x.shape)
        if rec["t1"][0] <= rec["t0"][0]:
            return x
        return _apply(x, rec, None)


class FrequencyMask(BaseTransform):
    """Zero a random run of frequency bins (transforms.py:49-70)."""

    def __init__(self, max_width: int = 10, prob: float = 0.5):
        self.max_width = max_width
        self.prob = prob

    def decide(self, rec, shape):
        if random.random() < self.prob:
            rec["f0"], rec["f1"] = _mask_interval(self.max_width, shape[-2])
        return None

    def __call__(self, x, seed=None):
        if seed is not None:
            _seed_host(seed)
        rec = _blank()
        self.decide(rec, x.shape)
        if rec["f1"][0] <= rec["f0"][0]:
            return x
        return _apply(x, rec, None)


_noise_calls = [0]


def _unseeded_noise_seed() -> int:
    _noise_calls[0] += 1
    return (torch.initial_seed() * 0x9E3779B1 + _noise_calls[0] * 0x85EBCA6B) & 0xFFFFFFFF


class GaussianNoise(BaseTransform):
    """x + N(0,1) * level, level ~ U(min_snr, max_snr) (transforms.py:73-97).

    noise_source "device": Philox normals generated inside the kernel (fast path; distributional parity).
    noise_source "torch_cpu": the N(0,1) draws come from the torch CPU generator exactly as the reference's
    `torch.randn_like(x)` on a CPU tensor does, then are shipped to the device (bit parity; used by the tests)."""

    def __init__(self, min_snr: float = 0.001, max_snr: float = 0.005, prob: float = 0.3, noise_source: str = "device"):
        if noise_source not in ("device", "torch_cpu"):
            raise ValueError("noise_source must be 'device' or 'torch_cpu'")
        self.min_snr = min_snr
        self.max_snr = max_snr
        self.prob = prob
        self.noise_source = noise_source

    def decide(self, rec, shape, seed=None):
        if random.random() < self.prob:
            level = random.uniform(self.min_snr, self.max_snr)
            rec["noise_level"] = np.float32(level)
            # seed=None: the device noise seed comes from a private counter mixed with torch's initial seed -- it must not
            # consume Python's `random` stream, which the reference never touches here
            rec["noise_seed"] = np.uint32((int(seed) if seed is not None else _unseeded_noise_seed()) & 0xFFFFFFFF)
            if self.noise_source == "torch_cpu":
                return torch.randn(tuple(shape))
        return None

    def __call__(self, x, seed=None):
        if seed is not None:
            _seed_host(seed)
        rec = _blank()
        noise = self.decide(rec, x.shape, seed)
        if rec["noise_level"][0] == 0.0:
            return x
        if noise is not None:
            noise = noise.to(x.device).contiguous()
        return _apply(x, rec, noise)


class TimeStretch(BaseTransform):
    """Kept for interface parity: the reference draws its random numbers and returns x unchanged
    (transforms.py:115-126), and build_augmentation_pipeline never instantiates it."""

    def __init__(self, min_rate: float = 0.9, max_rate: float = 1.1, prob: float = 0.5):
        self.min_rate = min_rate
        self.max_rate = max_rate
        self.prob = prob

    def decide(self, rec, shape):
        if random.random() < self.prob:
            random.uniform(self.min_rate, self.max_rate)
        return None

    def __call__(self, x, seed=None):
        if seed is not None:
            _seed_host(seed)
        self.decide(_blank(), x.shape)
        return x


class Compose:
    """Sequential pipeline; transform i is seeded with seed + i*1000 (transforms.py:129-144)."""

    def __init__(self, transforms: list):
        self.transforms = transforms

    def _fusable(self) -> bool:
        """One pass applies masks, then noise: only an order with every mask ahead of the noise can be fused."""
        seen_noise = False
        for t in self.transforms:
            if isinstance(t, GaussianNoise):
                seen_noise = True
            elif seen_noise and isinstance(t, (TimeMask, FrequencyMask)):
                return False
            elif not isinstance(t, (TimeMask, FrequencyMask, TimeStretch)):
                return False
        return True

    def __call__(self, x: torch.Tensor, seed: Optional[int] = None) -> torch.Tensor:
        if not self._fusable():       # any other composition: the reference's literal loop (transforms.py:139-144)
            for i, t in enumerate(self.transforms):
                x = t(x, seed=None if seed is None else seed + i * 1000)
            return x
        rec, noise = self.describe(seed, x.shape)
        if rec["t1"][0] <= rec["t0"][0] and rec["f1"][0] <= rec["f0"][0] and rec["noise_level"][0] == 0.0:
            return x
        if noise is not None:
            noise = noise.to(x.device).contiguous()
        return _apply(x, rec, noise)              # masks then noise, one pass (order fixed by the factory below)

    def describe(self, seed: Optional[int], shape):
        """All decisions of the pipeline for one call, as a one-element PcViewDesc record (+ explicit noise)."""
        rec, noise = _blank(), None
        seen_noise = False
        for i, t in enumerate(self.transforms):
            ts = None if seed is None else seed + i * 1000
            if ts is not None:
                _seed_host(ts)
            if isinstance(t, GaussianNoise):
                noise = t.decide(rec, shape, ts)
                seen_noise = True
            else:
                if seen_noise and isinstance(t, (TimeMask, FrequencyMask)):
                    raise NotImplementedError("fused pipeline applies masks before noise (the order "
                                              "build_augmentation_pipeline produces); call the transforms one by one")
                t.decide(rec, shape)
        return rec, noise


def build_augmentation_pipeline(config: dict, noise_source: str = "device") -> Compose:
    """Factory of transforms.py:147-182: fixed order time_mask, freq_mask, noise; `time_stretch` is never built."""
    transforms: List[BaseTransform] = []
    if config.get("time_mask", {}).get("enabled", False):
        p = config["time_mask"]
        transforms.append(TimeMask(max_width=p.get("max_width", 30), prob=p.get("prob", 0.5)))
    if config.get("freq_mask", {}).get("enabled", False):
        p = config["freq_mask"]
        transforms.append(FrequencyMask(max_width=p.get("max_width", 10), prob=p.get("prob", 0.5)))
    if config.get("noise", {}).get("enabled", False):
        p = config["noise"]
        transforms.append(GaussianNoise(min_snr=p.get("min_snr", 0.001), max_snr=p.get("max_snr", 0.005),
                                        prob=p.get("prob", 0.3), noise_source=noise_source))
    return Compose(transforms)


# ------------------------------------------------------------------------------------------------ batched descriptors
def waveform_gain(seed: int, prob: float = 0.5, lo: float = 0.8, hi: float = 1.2) -> float:
    """dataset.py:147-172 (_augment_waveform): seeds random / numpy / torch, then one decision + one uniform."""
    seed = int(seed)
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    if random.random() < prob:
        return random.uniform(lo, hi)
    return 1.0


def build_view_descriptors(indices: Sequence[int], n_views: int, n_freq: int, n_time: int, pipeline: Optional[Compose],
                           with_gain: bool = True, want_noise: bool = False):
    """Descriptor table for dataset items `indices` x `n_views` views, in the order the trainer flattens them
    (item-major: s0v0, s0v1, s1v0, ... -- trainer.py:192-197). Seeds follow dataset.py:85-94:
    gain seed = idx*10000 + view, pipeline seed = idx*20000 + view. The table depends only on (idx, view), never on
    the epoch, so callers cache it. Returns (records[np, VIEW_DESC_DTYPE], noise or None)."""
    recs = _blank(len(indices) * n_views)
    noises = []
    k = 0
    for ci, idx in enumerate(indices):
        idx = int(idx)
        for v in range(n_views):
            if pipeline is not None:
                r, nz = pipeline.describe(idx * 20000 + v, (1, 1, n_freq, n_time))
                recs[k] = r[0]
                if want_noise:
                    noises.append(nz.reshape(-1) if nz is not None else torch.zeros(n_freq * n_time))
            if with_gain:
                recs[k]["gain"] = np.float32(waveform_gain(idx * 10000 + v))
            recs[k]["clip"] = ci
            k += 1
    noise = torch.stack(noises) if (want_noise and noises) else None
    return recs, noise


def pack_view_descs(recs: np.ndarray, device) -> torch.Tensor:
    """numpy records -> uint8 CUDA tensor laid out as PcViewDesc[]."""
    host = torch.from_numpy(np.ascontiguousarray(recs).view(np.uint8).copy())
    if torch.device(device).type == "cuda":
        # pinned staging + asynchronous copy: a pageable copy would block the host until everything queued on the stream before
        # it (the previous training step) has finished
        host = host.pin_memory()
        return host.to(device, non_blocking=True)
    return host.to(device)
