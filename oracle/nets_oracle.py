"""Oracle: PhonemeNet / PhonemeNetDeep forward as plain torch functional ops on the CPU
(TEST INFRASTRUCTURE -- see oracle/__init__.py). Backward comes from torch autograd.

Follows src/models/phoneme_cnn.py: PhonemeNet :10-126, SpatialAttention :129-143,
ResidualBlock :146-184, PhonemeNetDeep :187-304. Parameters are taken from a state_dict with the
reference's key names, so the same dict drives oracle, reference and product.

Dropout2d is modelled by explicit per-(sample, channel) multiplier tensors (`drop[i]`, already scaled
by 1/(1-p)); None means dropout off -- the reference's masks come from the torch CPU bernoulli stream
and are not reproducible on a device (RNG parity).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _bn(x, sd, prefix, training, momentum=0.1, eps=1e-5, update=True):
    """nn.BatchNorm{1,2}d forward; in training mode also updates running stats in `sd` in place."""
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training and not update:
        rm, rv = rm.clone(), rv.clone()
    out = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training, momentum, eps)
    if training and update and (prefix + ".num_batches_tracked") in sd:
        sd[prefix + ".num_batches_tracked"] += 1
    return out


def _drop(x, m):
    return x if m is None else x * m[:, :, None, None]


def phoneme_net_forward(sd, x, training=True, use_attention=True, drop=None, update_stats=True):
    """PhonemeNet.forward (phoneme_cnn.py:98-126). drop: list of 3 [B,C] multipliers or None."""
    drop = drop or [None, None, None]
    for b in range(3):
        p = f"conv_blocks.{b}"
        x = F.conv2d(x, sd[f"{p}.0.weight"], sd[f"{p}.0.bias"], padding=1)          # :36,46,56
        x = F.relu(_bn(x, sd, f"{p}.1", training, update=update_stats))
        x = F.conv2d(x, sd[f"{p}.3.weight"], sd[f"{p}.3.bias"], padding=1)          # :39,49,59
        x = F.relu(_bn(x, sd, f"{p}.4", training, update=update_stats))
        if b < 2:
            x = F.max_pool2d(x, 2, 2)                                               # :42,52
        x = _drop(x, drop[b])                                                        # :43,53,62
    if use_attention:
        attn = torch.sigmoid(F.conv2d(x, sd["attention.conv.weight"], sd["attention.conv.bias"]))
        x = x * attn                                                                 # :134-143
    x = x.mean(dim=(2, 3))                                                           # :117-118
    x = F.linear(x, sd["projection.0.weight"], sd["projection.0.bias"])              # :75-77
    x = _bn(x, sd, "projection.1", training, update=update_stats)
    return F.normalize(x, p=2, dim=1)                                                # :124


def _resblock(sd, p, x, stride, training, dropm, update_stats):
    """ResidualBlock.forward (phoneme_cnn.py:173-184)."""
    out = F.conv2d(x, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], stride=stride, padding=1)
    out = F.relu(_bn(out, sd, f"{p}.bn1", training, update=update_stats))
    out = _drop(out, dropm)
    out = F.conv2d(out, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
    out = _bn(out, sd, f"{p}.bn2", training, update=update_stats)
    if f"{p}.shortcut.0.weight" in sd:
        sc = F.conv2d(x, sd[f"{p}.shortcut.0.weight"], sd[f"{p}.shortcut.0.bias"], stride=stride)
        sc = _bn(sc, sd, f"{p}.shortcut.1", training, update=update_stats)
    else:
        sc = x
    return F.relu(out + sc)


def _plainblock(sd, p, x, stride, training, dropm, update_stats):
    """The use_residual=False block (phoneme_cnn.py:231-245): conv(stride) BN ReLU conv BN ReLU Dropout2d."""
    x = F.conv2d(x, sd[f"{p}.0.weight"], sd[f"{p}.0.bias"], stride=stride, padding=1)
    x = F.relu(_bn(x, sd, f"{p}.1", training, update=update_stats))
    x = F.conv2d(x, sd[f"{p}.3.weight"], sd[f"{p}.3.bias"], padding=1)
    x = F.relu(_bn(x, sd, f"{p}.4", training, update=update_stats))
    return _drop(x, dropm)


def phoneme_net_deep_forward(sd, x, training=True, use_attention=True, drop=None, n_blocks=4, update_stats=True):
    """PhonemeNetDeep.forward (phoneme_cnn.py:274-304); residual blocks, or the plain blocks of use_residual=False
    (recognised by their state_dict keys). drop: list of n_blocks [B,C] or None."""
    drop = drop or [None] * n_blocks
    block = _resblock if "conv_blocks.0.conv1.weight" in sd else _plainblock
    x = F.conv2d(x, sd["init_conv.0.weight"], sd["init_conv.0.bias"], stride=1, padding=3)   # :212
    x = F.relu(_bn(x, sd, "init_conv.1", training, update=update_stats))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)                                   # :215
    for i in range(n_blocks):
        x = block(sd, f"conv_blocks.{i}", x, 1 if i == 0 else 2, training, drop[i], update_stats)      # :225-245
    if use_attention:
        attn = torch.sigmoid(F.conv2d(x, sd["attention.conv.weight"], sd["attention.conv.bias"]))
        x = x * attn
    x = x.mean(dim=(2, 3))
    x = F.linear(x, sd["projection.0.weight"], sd["projection.0.bias"])
    x = _bn(x, sd, "projection.1", training, update=update_stats)
    return F.normalize(x, p=2, dim=1)


def forward(arch: str, sd, x, **kw):
    if arch == "phoneme_cnn":
        return phoneme_net_forward(sd, x, **kw)
    if arch == "phoneme_cnn_deep":
        return phoneme_net_deep_forward(sd, x, **kw)
    raise ValueError(arch)


# --------------------------------------------------------------------------- parameter construction
def param_shapes(arch: str, cfg: dict | None = None):
    """Ordered (name, shape, kind) list equal to the reference module's state_dict."""
    cfg = cfg or {}
    emb = cfg.get("embedding_dim", 128)
    cin = cfg.get("in_channels", 1)
    att = cfg.get("use_attention", True)
    out = []

    def conv(p, co, ci, k):
        out.append((p + ".weight", (co, ci, k, k), "conv_w"))
        out.append((p + ".bias", (co,), "bias"))

    def bn(p, c):
        out.append((p + ".weight", (c,), "bn_w"))
        out.append((p + ".bias", (c,), "bias"))
        out.append((p + ".running_mean", (c,), "rm"))
        out.append((p + ".running_var", (c,), "rv"))
        out.append((p + ".num_batches_tracked", (), "nbt"))

    if arch == "phoneme_cnn":
        chans = [(cin, 32), (32, 64), (64, 128)]
        for b, (ci, co) in enumerate(chans):
            conv(f"conv_blocks.{b}.0", co, ci, 3); bn(f"conv_blocks.{b}.1", co)
            conv(f"conv_blocks.{b}.3", co, co, 3); bn(f"conv_blocks.{b}.4", co)
        last = 128
    elif arch == "phoneme_cnn_deep":
        hd = cfg.get("hidden_dims", [64, 128, 256, 512])
        conv("init_conv.0", hd[0], cin, 7); bn("init_conv.1", hd[0])
        ci = hd[0]
        for i, co in enumerate(hd):
            p = f"conv_blocks.{i}"
            if not cfg.get("use_residual", True):
                conv(p + ".0", co, ci, 3); bn(p + ".1", co)
                conv(p + ".3", co, co, 3); bn(p + ".4", co)
                ci = co
                continue
            conv(p + ".conv1", co, ci, 3); bn(p + ".bn1", co)
            conv(p + ".conv2", co, co, 3); bn(p + ".bn2", co)
            if i != 0 or ci != co:
                conv(p + ".shortcut.0", co, ci, 1); bn(p + ".shortcut.1", co)
            ci = co
        last = hd[-1]
    else:
        raise ValueError(arch)
    if att:
        conv("attention.conv", 1, last, 1)
    out.append(("projection.0.weight", (emb, last), "lin_w"))
    out.append(("projection.0.bias", (emb,), "bias"))
    bn("projection.1", emb)
    return out


def synthetic_state_dict(arch: str, cfg: dict | None = None, seed: int = 0):
    """Deterministic, version-independent parameters (numpy RandomState, not torch init) so golden
    fixtures need not store weights. Scales are chosen Kaiming-like so activations stay O(1);
    biases / BN affine / running stats are perturbed away from their defaults so that every term of
    the forward and backward is exercised."""
    import numpy as np

    rs = np.random.RandomState(seed)
    sd = {}
    for name, shape, kind in param_shapes(arch, cfg):
        if kind == "conv_w":
            fan_out = shape[0] * shape[2] * shape[3]
            v = rs.standard_normal(shape) * (2.0 / fan_out) ** 0.5
        elif kind == "lin_w":
            v = rs.standard_normal(shape) * 0.05
        elif kind == "bn_w":
            v = 1.0 + 0.1 * rs.standard_normal(shape)
        elif kind == "bias":
            v = 0.05 * rs.standard_normal(shape)
        elif kind == "rm":
            v = 0.1 * rs.standard_normal(shape)
        elif kind == "rv":
            v = 1.0 + 0.1 * rs.uniform(size=shape)
        elif kind == "nbt":
            sd[name] = torch.tensor(0, dtype=torch.long)
            continue
        sd[name] = torch.from_numpy(np.asarray(v, dtype=np.float32))
    return sd
