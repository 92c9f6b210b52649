"""pytest plugin used ONLY by tests/test_reference_suite.py: lets the reference's UNMODIFIED test modules (vendored into
oracle/_ref/ref_tests by oracle/build_ref.py) import `src.*` and get the drop-in.

    python -m pytest -p tests.ref_shim oracle/_ref/ref_tests/test_models.py ...

`src.models`, `src.training.losses`, `src.training.trainer`, `src.datasets.transforms` resolve to the matching
phoneme_contrast_b200 modules; `src.utils.logging` (off the hot path) is the reference's own file from oracle/_ref.
The reference tests build CPU tensors and pass device="cpu"; the product has no CPU path by design, so the shim -- not
the product -- moves tensors to cuda:0 at the four seams (model forward, loss forward, transform apply, trainer device)
and back. Nothing here is imported by the package."""
import importlib
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _alias(name, module):
    sys.modules[name] = module
    return module


def _install():
    import torch

    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import phoneme_contrast_b200.datasets as ds
    import phoneme_contrast_b200.datasets.transforms as tr
    import phoneme_contrast_b200.models as models
    import phoneme_contrast_b200.models.phoneme_cnn as cnn
    import phoneme_contrast_b200.training as training
    import phoneme_contrast_b200.training.losses as losses
    import phoneme_contrast_b200.training.trainer as trainer

    dev = torch.device("cuda", 0)
    # ---- seam 1: model forward
    orig_forward = cnn._FusedNet.forward

    def forward(self, x):
        if not x.is_cuda:
            self.to(dev)
            return orig_forward(self, x.to(dev)).cpu()
        return orig_forward(self, x)
    cnn._FusedNet.forward = forward
    # ---- seam 2: loss forward
    for cls in (losses.SupervisedContrastiveLoss, losses.NTXentLoss):
        orig = cls.forward

        def fwd(self, features, labels=None, *a, _orig=orig, **k):
            if isinstance(features, torch.Tensor) and not features.is_cuda:
                labels = labels.to(dev) if isinstance(labels, torch.Tensor) else labels
                return _orig(self, features.to(dev), labels, *a, **k).cpu()
            return _orig(self, features, labels, *a, **k)
        cls.forward = fwd
    # ---- seam 3: transform apply
    orig_apply = tr._apply

    def apply(x, rec, noise):
        if not x.is_cuda:
            return orig_apply(x.to(dev), rec, noise.to(dev) if noise is not None else None).cpu()
        return orig_apply(x, rec, noise)
    tr._apply = apply
    # ---- seam 4: trainer device
    orig_init = trainer.ContrastiveTrainer.__init__

    def init(self, *a, **k):
        if "device" in k and torch.device(k["device"]).type == "cpu":
            k["device"] = dev
            k["model"].to(dev)
        orig_init(self, *a, **k)
    trainer.ContrastiveTrainer.__init__ = init

    src = _alias("src", types.ModuleType("src"))
    src.__path__ = []
    _alias("src.models", models)
    _alias("src.models.phoneme_cnn", cnn)
    _alias("src.training", training)
    _alias("src.training.losses", losses)
    _alias("src.training.trainer", trainer)
    _alias("src.datasets", ds)
    _alias("src.datasets.transforms", tr)
    utils = _alias("src.utils", types.ModuleType("src.utils"))
    utils.__path__ = []
    spec = importlib.util.spec_from_file_location("src.utils.logging", os.path.join(REF, "src", "utils", "logging.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _alias("src.utils.logging", mod)


def pytest_configure(config):
    _install()
