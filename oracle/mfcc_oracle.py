"""Oracle: MFCC / log-mel front end (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates ``torchaudio.transforms.MFCC`` exactly as ``MFCCExtractor`` configures it
(reference src/datasets/features.py:25-55): Spectrogram(n_fft 400, win 400, hop 160,
periodic Hann, power 2, center, reflect, onesided) -> MelScale(80 mels, 0..8000 Hz,
htk, norm None) -> AmplitudeToDB("power", top_db 80) -> DCT-II ortho (40 coeffs).

torchaudio lines followed (site-packages/torchaudio, v2.11.0; reference pins 2.7.0):
  functional/functional.py:123-144   spectrogram -> torch.stft + |.|^power
  functional/functional.py:390-405   amplitude_to_DB (+ top_db clamp)
  functional/functional.py:518-580   melscale_fbanks
  functional/functional.py:636-665   create_dct
  transforms/_transforms.py:701-718  MFCC.forward
  functional/functional.py:compute_deltas (5-tap regression, replicate pad)
"""
from __future__ import annotations

import math

import numpy as np


# --------------------------------------------------------------------------- constants
def hann_periodic(n: int) -> np.ndarray:
    """torch.hann_window(n, periodic=True): 0.5 - 0.5 cos(2 pi k / n)."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def hz_to_mel_htk(f):
    return 2595.0 * np.log10(1.0 + f / 700.0)


def mel_to_hz_htk(m):
    return 700.0 * (10.0 ** (m / 2595.0) - 1.0)


def _linspace_f32(start, end, steps):
    """torch.linspace in float32 (ATen RangeFactories: start + step*i below the midpoint, end - step*(n-1-i) above)."""
    start, end = np.float32(start), np.float32(end)
    step = np.float32((end - start) / np.float32(steps - 1))
    i = np.arange(steps)
    lo = (start + step * i.astype(np.float32)).astype(np.float32)
    hi = (end - step * (steps - 1 - i).astype(np.float32)).astype(np.float32)
    return np.where(i < steps // 2, lo, hi).astype(np.float32)


def mel_filterbank(n_freqs: int = 201, f_min: float = 0.0, f_max: float = 8000.0,
                   n_mels: int = 80, sample_rate: int = 16000) -> np.ndarray:
    """melscale_fbanks(..., norm=None, mel_scale='htk') -> fb[n_freqs, n_mels].

    torchaudio builds this in float32 (torch.linspace); we mirror the fp32 steps so the
    filter weights agree with the reference's registered buffer to ~5e-6 absolute (the residual is a 1-ulp
    difference between torch's SLEEF powf and libm's in f_pts near 8 kHz; its effect on the MFCC is < 1e-6 relative).
    """
    all_freqs = _linspace_f32(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = _linspace_f32(m_min, m_max, n_mels + 2)
    f_pts = (np.float32(700.0) * (np.float32(10.0) ** (m_pts / np.float32(2595.0)) - np.float32(1.0))).astype(np.float32)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(np.float32(0.0), np.minimum(down, up))
    return fb.astype(np.float32)


def dct_matrix(n_mfcc: int = 40, n_mels: int = 80) -> np.ndarray:
    """create_dct(n_mfcc, n_mels, norm='ortho') -> [n_mels, n_mfcc]."""
    n = np.arange(n_mels, dtype=np.float64)
    k = np.arange(n_mfcc, dtype=np.float64)[:, None]
    dct = np.cos(math.pi / n_mels * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / n_mels)
    return dct.T.copy()


# --------------------------------------------------------------------------- pipeline
def preemphasis(wave: np.ndarray, coeff: float) -> np.ndarray:
    """torchaudio.functional.preemphasis (functional.py: `waveform[..., 1:] -= coeff * waveform[..., :-1]`): the optional
    stage north_star names; the reference's MFCCExtractor has none (coeff 0 = identity)."""
    wave = np.array(wave, dtype=np.float64, copy=True)
    if coeff:
        wave[..., 1:] -= coeff * np.asarray(wave[..., :-1]).copy()
    return wave


def power_spectrogram(wave: np.ndarray, n_fft: int = 400, hop: int = 160, preemph: float = 0.0) -> np.ndarray:
    """wave [B,S] -> |STFT|^2 [B, n_fft//2+1, T], center=True, reflect pad, periodic Hann."""
    wave = preemphasis(wave, preemph)
    pad = n_fft // 2
    padded = np.pad(wave, ((0, 0), (pad, pad)), mode="reflect")
    n_frames = 1 + (padded.shape[1] - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = padded[:, idx] * hann_periodic(n_fft)[None, None, :]          # [B,T,n_fft]
    spec = np.fft.rfft(frames, n=n_fft, axis=-1)                            # [B,T,201]
    power = spec.real ** 2 + spec.imag ** 2
    return np.transpose(power, (0, 2, 1))                                   # [B,201,T]


def mel_db(wave: np.ndarray, *, top_db: float | None = 80.0, clamp_scope: str = "clip",
           n_fft: int = 400, hop: int = 160, n_mels: int = 80, sample_rate: int = 16000,
           f_min: float = 0.0, f_max: float | None = None, preemph: float = 0.0) -> np.ndarray:
    """wave [B,S] -> mel power in dB [B, n_mels, T].

    clamp_scope 'clip': the amax of the top_db clamp is taken per clip (what the dataset does:
    one clip per call, dataset.py:90). 'call': over the whole 3-D tensor, which is what
    amplitude_to_DB does for a batched [B,80,T] input (functional.py:396-399).
    """
    f_max = f_max or sample_rate / 2
    power = power_spectrogram(wave, n_fft, hop, preemph)
    fb = mel_filterbank(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate).astype(np.float64)
    mel = np.einsum("bft,fm->bmt", power, fb)
    db = 10.0 * np.log10(np.maximum(mel, 1e-10))
    if top_db is not None:
        if clamp_scope == "clip":
            floor = db.max(axis=(1, 2), keepdims=True) - top_db
        elif clamp_scope == "call":
            floor = db.max() - top_db
        else:
            raise ValueError(clamp_scope)
        db = np.maximum(db, floor)
    return db


def compute_deltas(x: np.ndarray, win_length: int = 5) -> np.ndarray:
    """torchaudio.functional.compute_deltas: d_t = sum_n n (c_{t+n} - c_{t-n}) / (2 sum n^2), replicate pad."""
    n = (win_length - 1) // 2
    denom = n * (n + 1) * (2 * n + 1) / 3
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(n, n)], mode="edge")
    out = np.zeros_like(x, dtype=np.float64)
    T = x.shape[-1]
    for k in range(-n, n + 1):
        out += k * xp[..., n + k:n + k + T]
    return out / denom


def mfcc(wave: np.ndarray, *, n_mfcc: int = 40, clamp_scope: str = "clip", add_delta: bool = False,
         add_delta_delta: bool = False, **kw) -> np.ndarray:
    """MFCCExtractor.forward (features.py:61-103): wave [B,S] | [S] | [B,1,S] -> [B,1,n_mfcc*k,T] (fp64)."""
    wave = np.asarray(wave)
    if wave.ndim == 1:
        wave = wave[None]
    elif wave.ndim == 3:
        wave = wave[:, 0]
    db = mel_db(wave, clamp_scope=clamp_scope, **kw)
    n_mels = db.shape[1]
    out = np.einsum("bmt,mc->bct", db, dct_matrix(n_mfcc, n_mels))
    feats = [out]
    if add_delta:
        feats.append(compute_deltas(out))
    if add_delta_delta:
        feats.append(compute_deltas(feats[1] if add_delta else compute_deltas(out)))
    return np.concatenate(feats, axis=1)[:, None]


def log_mel(wave: np.ndarray, **kw) -> np.ndarray:
    """MelSpectrogramExtractor.forward (features.py:134-153): AmplitudeToDB() has top_db=None."""
    wave = np.asarray(wave)
    if wave.ndim == 1:
        wave = wave[None]
    elif wave.ndim == 3:
        wave = wave[:, 0]
    return mel_db(wave, top_db=None, **kw)[:, None]


# --------------------------------------------------------------------------- CPU-timing variant
def mfcc_torch_cpu(wave, *, per_clip: bool = True):
    """Same pipeline through the library calls torchaudio itself makes (torch.stft, matmul, log10),
    fp32 on the host CPU. This is what bench.py times as the CPU baseline (kind "port"): the op
    sequence is the reference's, so the timing is representative of MFCCExtractor on CPU.
    per_clip=True loops one clip per call like dataset.py:90; False is one batched call."""
    import torch

    wave = torch.as_tensor(wave, dtype=torch.float32)
    window = torch.hann_window(400)
    fb = torch.from_numpy(mel_filterbank())
    dct = torch.from_numpy(dct_matrix().astype(np.float32))

    def one(w):
        spec = torch.stft(w, 400, 160, 400, window, center=True, pad_mode="reflect", normalized=False,
                          onesided=True, return_complex=True)
        power = spec.abs().pow(2.0)
        mel = torch.matmul(power.transpose(-1, -2), fb).transpose(-1, -2)
        db = 10.0 * torch.log10(torch.clamp(mel, min=1e-10))
        db = torch.max(db, db.amax() - 80.0)
        return torch.matmul(db.transpose(-1, -2), dct).transpose(-1, -2).unsqueeze(1)

    if per_clip:
        return torch.cat([one(wave[i:i + 1]) for i in range(wave.shape[0])], 0)
    return one(wave)
