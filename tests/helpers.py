"""Shared test helpers."""
import re

import numpy as np

# Biases of a conv / linear that feeds a train-mode BatchNorm have an analytically ZERO gradient
# (the batch mean absorbs them); what the reference reports there is fp32 round-off of order 1e-5
# that differs from run to run with the thread count. They are compared with an absolute floor.
_ZERO_GRAD = re.compile(r"^(conv_blocks\.\d+\.(0|3|conv1|conv2|shortcut\.0)|init_conv\.0|projection\.0)\.bias$")


def analytically_zero_grad(name: str) -> bool:
    return bool(_ZERO_GRAD.match(name))


def grad_atol(name: str, ref: np.ndarray, rel: float = 1e-4) -> float:
    if analytically_zero_grad(name):
        return 1e-4
    return rel * max(float(np.abs(ref).max()), 1e-6)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
