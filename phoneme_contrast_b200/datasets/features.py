"""Drop-in for src/datasets/features.py: FeatureExtractor / MFCCExtractor / MelSpectrogramExtractor /
build_feature_extractor with the reference's constructor arguments, input-rank handling (features.py:70-74)
and output layout [B, 1, F, T]. The arithmetic of torchaudio.transforms.MFCC runs in one fused sm_100a
kernel (csrc/mfcc.cu); torch is used on the host only to build the constant tables once.

Clamp semantics. AmplitudeToDB(top_db=80) clamps to (max - 80) with the max taken over the whole call for a
[B,80,T] input (torchaudio functional.py:396-399). `forward(batch)` reproduces exactly that (clamp_scope
"call"). The training path calls the extractor one clip at a time (dataset.py:90), i.e. per-clip clamping;
`forward_views` is the batched equivalent of that loop (clamp_scope "clip") with the view augmentations fused.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from .._lib import PcMfccConsts, call, ptr, stream


# ------------------------------------------------------------------------------------------------ constants
def _melscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """Triangular HTK mel filterbank, norm=None, built with the same fp32 torch ops as torchaudio's
    melscale_fbanks (functional.py:518-580) so the table is bit-identical to the reference's buffer."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def _create_dct(n_mfcc: int, n_mels: int) -> torch.Tensor:
    """DCT-II, norm='ortho' -> [n_mels, n_mfcc] (torchaudio functional.py:636-665, same fp32 ops)."""
    n = torch.arange(float(n_mels))
    k = torch.arange(float(n_mfcc)).unsqueeze(1)
    dct = torch.cos(math.pi / float(n_mels) * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / float(n_mels))
    return dct.t().contiguous()


def _twiddles() -> torch.Tensor:
    """[W200 (200 complex) | W400 (201 complex)] interleaved re/im, computed in fp64."""
    k = np.arange(200, dtype=np.float64)
    w200 = np.exp(-2j * np.pi * k / 200.0)
    k = np.arange(201, dtype=np.float64)
    w400 = np.exp(-2j * np.pi * k / 400.0)
    out = np.concatenate([np.stack([w200.real, w200.imag], 1).ravel(), np.stack([w400.real, w400.imag], 1).ravel()])
    return torch.from_numpy(out.astype(np.float32))


def _band_form(fb: torch.Tensor):
    """fb [n_freqs, n_mels] -> (start[n_mels], len[n_mels], w[n_mels, PC_FB_MAXW]) covering each filter's support."""
    n_freqs, n_mels = fb.shape
    start = torch.zeros(n_mels, dtype=torch.int32)
    length = torch.zeros(n_mels, dtype=torch.int32)
    w = torch.zeros(n_mels, L.PC_FB_MAXW, dtype=torch.float32)
    for m in range(n_mels):
        nz = torch.nonzero(fb[:, m]).flatten()
        if nz.numel() == 0:
            continue
        lo, hi = int(nz[0]), int(nz[-1]) + 1
        if hi - lo > L.PC_FB_MAXW:
            raise NotImplementedError(f"mel filter {m} spans {hi - lo} FFT bins (> {L.PC_FB_MAXW}); lower n_fft/n_mels ratio not built")
        start[m], length[m] = lo, hi - lo
        w[m, :hi - lo] = fb[lo:hi, m]
    return start, length, w


class _FrontEndConsts:
    def __init__(self, sample_rate, n_fft, hop_length, n_mels, f_min, f_max, n_mfcc, preemphasis=0.0):
        self.n_fft, self.hop, self.n_mels, self.n_mfcc = n_fft, hop_length, n_mels, n_mfcc
        self.preemphasis = float(preemphasis)
        self.window = torch.hann_window(n_fft)                                     # periodic, fp32 (torchaudio default)
        self.fb = _melscale_fbanks(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate)
        self.dct = _create_dct(n_mfcc, n_mels) if n_mfcc else None
        self.start, self.length, self.w = _band_form(self.fb)
        self.tw = _twiddles()
        self._dev = {}

    def on(self, device) -> "tuple[PcMfccConsts, list]":
        key = str(device)
        if key not in self._dev:
            keep = [t.to(device) if t is not None else None for t in (self.window, self.start, self.length, self.w, self.dct, self.tw)]
            st = PcMfccConsts(ptr(keep[0]), ptr(keep[1], torch.int32), ptr(keep[2], torch.int32), ptr(keep[3]),
                              ptr(keep[4]) if keep[4] is not None else None, ptr(keep[5]), self.n_fft, self.hop, self.n_mels,
                              self.n_mfcc or 0, int(self.length.max()), self.preemphasis)
            self._dev[key] = (st, keep)
        return self._dev[key]


def _as_batch(waveform: torch.Tensor) -> torch.Tensor:
    """features.py:70-74: [S] -> [1,S]; [B,1,S] -> [B,S]."""
    if waveform.dim() == 1:
        waveform = waveform.unsqueeze(0)
    elif waveform.dim() == 3:
        waveform = waveform.squeeze(1)
    if waveform.dim() != 2:
        raise ValueError(f"expected waveform [batch, samples], got shape {tuple(waveform.shape)}")
    if not waveform.is_cuda:
        raise RuntimeError("phoneme_contrast_b200 feature extractors run on CUDA tensors only (no CPU fallback)")
    return waveform.to(torch.float32).contiguous()


def _run_frontend(consts: _FrontEndConsts, wave, kind, clamp_mode, top_db, views=None, n_views=None, noise=None,
                  clamp_ref=None, want_max=False, n_clips=None):
    B, S = wave.shape
    if n_clips is not None:      # `wave` is a larger resident cache; the views' clip indices select the rows this call reads
        B = int(n_clips)
    st, _keep = consts.on(wave.device)
    T = 1 + S // consts.hop
    n_out = consts.n_mfcc if kind == L.FE_MFCC else consts.n_mels
    nv = B if views is None else n_views
    out = torch.empty(nv, 1, n_out, T, device=wave.device, dtype=torch.float32)
    vmax = torch.empty(nv, device=wave.device, dtype=torch.float32) if want_max else None
    call("pc_frontend_fwd", ptr(wave), B, S, wave.stride(0), C.byref(st), ptr(views, torch.uint8) if views is not None else None,
         nv, ptr(noise), kind, clamp_mode, float(top_db), ptr(clamp_ref), ptr(vmax), ptr(out), stream())
    return out, vmax


def _deltas(x: torch.Tensor) -> torch.Tensor:
    B, F_, T = x.shape
    out = torch.empty_like(x)
    call("pc_compute_deltas", ptr(x), B * F_, T, ptr(out), stream())
    return out


# ------------------------------------------------------------------------------------------------ public API
class FeatureExtractor(nn.Module):
    """Base class for feature extractors (features.py:9-19)."""

    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class MFCCExtractor(FeatureExtractor):
    """MFCC with optional delta / delta-delta (constructor of features.py:25-59)."""

    top_db = 80.0

    def __init__(self, sample_rate: int = 16000, n_mfcc: int = 40, n_fft: int = 400, hop_length: int = 160,
                 n_mels: int = 80, f_min: float = 0.0, f_max: Optional[float] = None, add_delta: bool = False,
                 add_delta_delta: bool = False, preemphasis: float = 0.0):
        # preemphasis (extension, default off = the reference's arithmetic): y[n] = x[n] - a*x[n-1] before the STFT, the
        # stage BASELINE.json's north_star names and the manuscript quotes at 0.97; the reference's code has none
        super().__init__()
        if n_mfcc > n_mels:
            raise ValueError("Cannot select more MFCC coefficients than # mel bins")      # torchaudio MFCC.__init__
        self.sample_rate = sample_rate
        self.n_mfcc = n_mfcc
        self.add_delta = add_delta
        self.add_delta_delta = add_delta_delta
        self._consts = _FrontEndConsts(sample_rate, n_fft, hop_length, n_mels, f_min, f_max or sample_rate / 2, n_mfcc, preemphasis)

    def _mfcc(self, wave, clamp_scope):
        if clamp_scope == "clip" or wave.shape[0] == 1:
            out, _ = _run_frontend(self._consts, wave, L.FE_MFCC, L.CLAMP_PER_CLIP, self.top_db)
            return out
        # whole-call clamp: pass 1 finds every clip's max dB, pass 2 clamps against the global max
        _, vmax = _run_frontend(self._consts, wave, L.FE_MFCC, L.CLAMP_NONE, self.top_db, want_max=True)
        gmax = torch.empty(1, device=wave.device, dtype=torch.float32)
        call("pc_reduce_max", ptr(vmax), vmax.numel(), ptr(gmax), stream())
        out, _ = _run_frontend(self._consts, wave, L.FE_MFCC, L.CLAMP_GIVEN, self.top_db, clamp_ref=gmax)
        return out

    def forward(self, waveform: torch.Tensor, clamp_scope: str = "call") -> torch.Tensor:
        """waveform [B,S] | [S] | [B,1,S] (CUDA) -> [B, 1, n_mfcc*(1+delta+delta2), T]."""
        if clamp_scope not in ("call", "clip"):
            raise ValueError("clamp_scope must be 'call' or 'clip'")
        wave = _as_batch(waveform)
        mfcc = self._mfcc(wave, clamp_scope)                                             # [B,1,F,T]
        if not (self.add_delta or self.add_delta_delta):
            return mfcc
        base = mfcc[:, 0]
        feats = [base]
        if self.add_delta:
            feats.append(_deltas(base))
        if self.add_delta_delta:
            feats.append(_deltas(feats[1] if self.add_delta else _deltas(base)))          # features.py:88-95
        return torch.cat(feats, dim=1).unsqueeze(1)

    def forward_views(self, waveform: torch.Tensor, views: torch.Tensor, n_views: int, noise: Optional[torch.Tensor] = None,
                      n_clips: Optional[int] = None):
        """Batched form of the dataset loop (dataset.py:79-98): for clip i, views[i*V:(i+1)*V] (a uint8 tensor
        holding PcViewDesc records, see datasets.transforms.pack_view_descs) give gain / masks / noise of each
        view. Per-clip clamp. Returns [B*V, 1, n_mfcc, T]. noise: optional explicit N(0,1) draws [B*V, F*T].
        n_clips: when given, `waveform` is a device-resident cache with more rows than this batch and the descriptors' clip
        field names the row each group of V views reads (datasets.dataset.PhonemeContrastiveDataset.batch_views)."""
        if self.add_delta or self.add_delta_delta:
            raise NotImplementedError("forward_views covers the plain-MFCC training configuration")
        wave = _as_batch(waveform)
        out, _ = _run_frontend(self._consts, wave, L.FE_MFCC, L.CLAMP_PER_CLIP, self.top_db, views=views, n_views=n_views, noise=noise,
                               n_clips=n_clips)
        return out


class MelSpectrogramExtractor(FeatureExtractor):
    """Log-mel spectrogram (features.py:109-153); AmplitudeToDB() there has top_db=None, i.e. no clamp."""

    def __init__(self, sample_rate: int = 16000, n_fft: int = 400, hop_length: int = 160, n_mels: int = 80,
                 f_min: float = 0.0, f_max: Optional[float] = None, preemphasis: float = 0.0):
        super().__init__()
        self.sample_rate = sample_rate
        self._consts = _FrontEndConsts(sample_rate, n_fft, hop_length, n_mels, f_min, f_max or sample_rate / 2, 0, preemphasis)

    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        wave = _as_batch(waveform)
        out, _ = _run_frontend(self._consts, wave, L.FE_LOGMEL, L.CLAMP_NONE, 0.0)
        return out


def build_feature_extractor(config: Dict) -> FeatureExtractor:
    """Build feature extractor from config (features.py:156-168)."""
    extractor_type = config.get("type", "mfcc")
    if extractor_type == "mfcc":
        return MFCCExtractor(**config.get("mfcc_params", {}))
    if extractor_type == "mel":
        return MelSpectrogramExtractor(**config.get("mel_params", {}))
    raise ValueError(f"Unknown feature extractor type: {extractor_type}")
