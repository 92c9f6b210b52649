"""Per-layer timing of the stride-1 3x3 convolutions of cnn_deep at 256 views: round-1 per-tap-gather kernel vs the halo-resident
engine at cluster sizes 1 / 2 / 4 (forward and data gradient). CUDA events, 5 warm-up + 20 timed launches, L2 flushed between
launches by writing a 256 MB buffer. Prints us per launch and algorithmic TFLOP/s."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from phoneme_contrast_b200 import _lib as L
from phoneme_contrast_b200 import ops

DEV = "cuda"
B = int(os.environ.get("HB_BATCH", "256"))
LAYERS = [(20, 51, 64, 64), (10, 26, 128, 128), (5, 13, 256, 256), (3, 7, 512, 512)]
if os.environ.get("HB_LAYERS"):
    LAYERS = [LAYERS[int(i)] for i in os.environ["HB_LAYERS"].split(",")]
flush = torch.empty(64 * 1024 * 1024, device=DEV, dtype=torch.float32)


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


for (H, W, Cin, Cout) in LAYERS:
    g = ops.conv_geom(B, H, W, Cin, Cout, 3, 1, 1)
    x = torch.randn(B, H, W, Cin, device=DEV)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV) * (2.0 / (Cin * 9)) ** 0.5
    bias = torch.randn(Cout, device=DEV)
    cw = ops.ConvWeights(w, g, L.PREC_FP16X2)
    planes = ops.bn_act_split(x)
    st = torch.zeros(2, Cout, device=DEV, dtype=torch.float64)
    dy = torch.randn(B, H, W, Cout, device=DEV) * 1e-6
    amax = dy.abs().max().reshape(1) * 1.5
    scale = 2.0 ** (14 - torch.floor(torch.log2(amax)))
    dyp = ops.bn_act_split((dy * scale).contiguous())
    flops = 2.0 * B * H * W * Cin * Cout * 9
    row = [f"{H}x{W} {Cin}->{Cout}"]
    for name, env in (("old", {"PC_CONV_HALO": "0"}), ("stream", {"PC_CONV_HALO": "1", "PC_HALO_CLUSTER": "1", "PC_HALO_RESIDENT": "0"}),
                      ("default", {"PC_CONV_HALO": "1", "PC_HALO_CLUSTER": "1", "PC_HALO_RESIDENT": "1", "PC_HALO_ALL": "1"}),
                      ("cl2", {"PC_CONV_HALO": "1", "PC_HALO_CLUSTER": "2", "PC_HALO_RESIDENT": "1", "PC_HALO_ALL": "1"}),
                      ("cl4", {"PC_CONV_HALO": "1", "PC_HALO_CLUSTER": "4", "PC_HALO_RESIDENT": "1", "PC_HALO_ALL": "1"})):
        os.environ.update(env)
        tf = timeit(lambda: ops.conv_fwd(planes, cw.wf, bias, g, dict(presplit=True), st, cw.prec_f))
        td = timeit(lambda: ops.conv_dgrad(dyp, cw.wd, g, prec=cw.prec_d, dy_amax=amax, dy_presplit=True))
        row.append(f"{name}: fwd {tf:7.1f} us ({flops / tf / 1e6:6.1f} TF/s) dgrad {td:7.1f} us ({flops / td / 1e6:6.1f} TF/s)")
    print(" | ".join(row), flush=True)
