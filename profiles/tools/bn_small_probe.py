"""Warm timings (CUDA events, 50 iterations) of the BatchNorm backward passes at the cnn_small shapes: is their time bandwidth or latency?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from phoneme_contrast_b200 import ops
from phoneme_contrast_b200._lib import call, ptr, stream

DEV = "cuda"


def timeit(fn, n=50, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


g = torch.cuda.CUDAGraph()
for (B, H, W, C_, pool) in ((64, 40, 101, 32, 0), (64, 40, 101, 32, 2), (64, 20, 50, 64, 0), (64, 20, 50, 64, 2), (64, 10, 25, 128, 0), (256, 20, 51, 64, 0)):
    y = torch.randn(B, H, W, C_, device=DEV)
    Ho, Wo = ops.pool_dims(H, W, pool)
    dout = torch.randn(B, Ho, Wo, C_, device=DEV) * 1e-5
    bn = torch.nn.BatchNorm2d(C_).to(DEV)
    st = torch.zeros(2, C_, device=DEV, dtype=torch.float64)
    st[0] = y.double().sum((0, 1, 2)); st[1] = (y.double() ** 2).sum((0, 1, 2))
    co = ops.bn_finalize(st, B * H * W, bn, True)
    sums = torch.zeros(2, C_, device=DEV, dtype=torch.float64)
    maxes = torch.zeros(2, device=DEV)
    args = (ptr(dout), ptr(y), B, H, W, C_, ptr(co.scale), ptr(co.shift), ptr(co.mean), ptr(co.invstd), None, pool, None)

    def reduce():
        call("pc_bn_act_bwd_reduce", *args, ptr(sums, torch.float64), ptr(maxes), stream())

    a1 = torch.zeros(1, device=DEV)

    def both():
        ops.bn_act_bwd(dout, y, co, pool, None, None, amax=a1, planes=True)

    # the same inside a captured graph of 20 back-to-back calls (no launch gaps)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        both()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(20):
                both()
    mb = (y.numel() + dout.numel()) * 4 / 1e6
    print(f"{B}x{H}x{W}x{C_} pool {pool}: {mb:.0f} MB in | reduce alone {timeit(reduce):.1f} us | reduce+apply eager {timeit(both):.1f} us | in a graph {timeit(gr.replay, n=10, warm=3) / 20:.1f} us per pair")
