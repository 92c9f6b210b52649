"""Diagnostic: error of the bf16 single-product mode vs the oracle at several batch sizes (is north_star's 1e-2 met?)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import nets_oracle, supcon_oracle
from phoneme_contrast_b200.models import model_registry
from phoneme_contrast_b200.training import get_loss_fn
for prec in ("bf16", "fp16x2"):
    os.environ["PC_PRECISION"] = prec
    for arch, B in (("phoneme_cnn_deep", 8), ("phoneme_cnn_deep", 64), ("phoneme_cnn", 64)):
        cfg = {"dropout_rate": 0.0}
        sd = nets_oracle.synthetic_state_dict(arch, cfg, seed=3)
        rs = np.random.RandomState(8)
        x = rs.standard_normal((B, 1, 40, 101)).astype(np.float32)
        y = np.repeat(np.arange(B // 2) // 4, 2).astype(np.int64)
        m = model_registry.create(arch, cfg).cuda(); m.load_state_dict(sd); m.train()
        emb = m(torch.from_numpy(x).cuda())
        loss = get_loss_fn("supervised_contrastive", temperature=0.15)(emb, torch.from_numpy(y).cuda())
        live = {k: v.clone() for k, v in sd.items()}
        e_ref = nets_oracle.forward(arch, live, torch.from_numpy(x), training=True)
        l_ref = supcon_oracle.loss_torch_cpu(e_ref, torch.from_numpy(y), temperature=0.15)
        e = emb.detach().cpu().numpy(); er = e_ref.numpy()
        print(prec, arch, B, "emb rel-L2 %.3e max-abs/max %.3e loss rel %.3e" % (np.linalg.norm(e - er) / np.linalg.norm(er), np.abs(e - er).max() / np.abs(er).max(), abs(float(loss) - float(l_ref)) / abs(float(l_ref))))
