"""Does the halo weight-gradient kernel overlap with the BatchNorm-backward passes / the data-gradient kernel of the next layer when they
run on two streams? Times each alone and both together (CUDA events around the pair) at the block-0 and block-1 shapes of cnn_deep."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from phoneme_contrast_b200 import _lib as L
from phoneme_contrast_b200 import ops

DEV = "cuda"
B = 256


def timeit(fn, n=20, warm=5):
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


for (H, W, C_) in ((20, 51, 64), (10, 26, 128)):
    g = ops.conv_geom(B, H, W, C_, C_, 3, 1, 1)
    x = torch.relu(torch.randn(B, H, W, C_, device=DEV))
    y = torch.randn(B, H, W, C_, device=DEV)
    dout = torch.randn(B, H, W, C_, device=DEV) * 1e-6
    w = torch.randn(C_, C_, 3, 3, device=DEV) * 0.05
    bn = torch.nn.BatchNorm2d(C_).to(DEV)
    st = torch.zeros(2, C_, device=DEV, dtype=torch.float64)
    st[0] = y.double().sum((0, 1, 2)); st[1] = (y.double() ** 2).sum((0, 1, 2))
    co = ops.bn_finalize(st, B * H * W, bn, True)
    cw = ops.ConvWeights(w, g, L.PREC_FP16X2)
    a1 = torch.zeros(1, device=DEV)
    dy_ps, _, _ = ops.bn_act_bwd(dout, y, co, 0, None, None, amax=a1, planes=True)
    x_ps = ops.bn_act_split(x)
    dw = torch.empty_like(w)
    side = torch.cuda.Stream()
    main = torch.cuda.current_stream()

    def wgrad():
        ops.conv_wgrad(x_ps, dy_ps, g, dict(presplit=True), dw=dw, prec=L.PREC_FP16X2, dy_amax=a1, dy_presplit=True, want_db=False)

    a2 = torch.zeros(1, device=DEV)

    def bnbwd():
        ops.bn_act_bwd(dout, y, co, 0, None, None, amax=a2, planes=True)

    def dgrad():
        ops.conv_dgrad(dy_ps, cw.wd, g, prec=cw.prec_d, dy_amax=a1, dy_presplit=True)

    def both(f_main):
        side.wait_stream(main)
        with torch.cuda.stream(side):
            wgrad()
        f_main()
        main.wait_stream(side)

    tw, tb, td = timeit(wgrad), timeit(bnbwd), timeit(dgrad)
    twb, twd = timeit(lambda: both(bnbwd)), timeit(lambda: both(dgrad))
    print(f"{H}x{W}x{C_}: wgrad {tw:.1f} us | bn bwd (reduce+apply) {tb:.1f} | dgrad {td:.1f} | wgrad||bn {twb:.1f} (sum {tw + tb:.1f}, max {max(tw, tb):.1f}) | "
          f"wgrad||dgrad {twd:.1f} (sum {tw + td:.1f})", flush=True)
