"""Seeded synthetic clips shared by the golden generator and the tests (numpy RandomState: stable across versions)."""
import numpy as np


def synth_waves(n, s, seed):
    rs = np.random.RandomState(seed)
    w = (0.1 * rs.standard_normal((n, s))).astype(np.float32)
    for i in range(1, n, 2):
        cut = int(s * (0.25 + 0.5 * rs.uniform()))
        w[i, cut:] = 0.0
    t = np.arange(s) / 16000.0
    w[0] = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.01 * np.sin(2 * np.pi * 3000 * t)).astype(np.float32)
    return w
