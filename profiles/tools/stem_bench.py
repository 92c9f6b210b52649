"""Timing of the stem kernels at 256 views (CUDA events, L2 flushed between launches): Gram, statistics, fused forward, pooled backward,
and the round-1 two-pass forward for comparison."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from phoneme_contrast_b200 import _lib as L
from phoneme_contrast_b200 import ops

DEV = "cuda"
B, H, W = int(os.environ.get("SB_BATCH", "256")), 40, 101
flush = torch.empty(64 * 1024 * 1024, device=DEV, dtype=torch.float32)


def timeit(fn, n=10, warm=3):
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


x = torch.randn(B, 1, H, W, device=DEV)
conv = torch.nn.Conv2d(1, 64, 7, padding=3).to(DEV)
bn = torch.nn.BatchNorm2d(64).to(DEV)
st = torch.zeros(2, 64, device=DEV, dtype=torch.float64)
gram = ops.stem_gram(x)
ops.stem_stats_from_gram(gram, conv, B, H, W, st)
co = ops.bn_finalize(st, B * H * W, bn, True)
print("gram            %8.1f us" % timeit(lambda: ops.stem_gram(x)))
print("stats from gram %8.1f us" % timeit(lambda: ops.stem_stats_from_gram(gram, conv, B, H, W, st)))
print("fused forward   %8.1f us" % timeit(lambda: ops.stem_fwd(x, conv, co, want_planes=True)))
g = ops.conv_geom(B, H, W, 1, 64, 7, 1, 3)
y0 = ops.conv_fwd(x, conv.weight, conv.bias, g, None, st, L.PREC_FP32)
print("r1 conv (SIMT)  %8.1f us" % timeit(lambda: ops.conv_fwd(x, conv.weight, conv.bias, g, None, st, L.PREC_FP32)))
print("r1 bn+relu+pool %8.1f us" % timeit(lambda: ops.bn_act_fwd(y0, co, 3, None, want_planes=True)))
p0, am, _ = ops.stem_fwd(x, conv, co, want_planes=True)
dpool = torch.randn_like(p0) * 1e-6
dw, db, dg, dbt = (torch.empty_like(conv.weight), torch.empty(64, device=DEV), torch.empty(64, device=DEV), torch.empty(64, device=DEV))
print("pooled backward %8.1f us" % timeit(lambda: ops.stem_bwd(dpool, p0, am, x, conv, co, gram, dw, db, dg, dbt)))
