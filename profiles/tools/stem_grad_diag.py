"""Diagnostic: per-tensor gradient rel-L2 of the benchmarked cnn_deep step vs the oracle, for the exact-fp32 SIMT path and the
default fp16x2 path (eager), to separate arithmetic error from gate-flip sensitivity (max-pool argmax / ReLU ties)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.test_gpu_r2 import _bench_case, _trainer
from tests.helpers import analytically_zero_grad
from phoneme_contrast_b200.training import get_loss_fn
arch = "phoneme_cnn_deep"
case = _bench_case(arch)
for prec in ("fp32", "fp16x2", "fp16x2"):
    os.environ["PC_PRECISION"] = prec
    m, _ = _trainer(arch, case, False)
    emb = m(case["x"].cuda())
    loss = get_loss_fn("supervised_contrastive", temperature=0.15)(emb, case["y"].cuda())
    loss.backward()
    errs = []
    for n, p in m.named_parameters():
        if analytically_zero_grad(n):
            continue
        g, r = p.grad.cpu().numpy().astype(np.float64), case["grads"][n].astype(np.float64)
        errs.append((np.linalg.norm(g - r) / max(np.linalg.norm(r), 1e-30), n))
    errs.sort(reverse=True)
    e = emb.detach().cpu().numpy()
    print(prec, "emb max rel %.2e loss rel %.2e | worst grads:" % (np.abs(e - case["emb"]).max() / np.abs(case["emb"]).max(), abs(float(loss) - case["loss"]) / abs(case["loss"])),
          ", ".join("%s %.2e" % (n, v) for v, n in errs[:6]))
