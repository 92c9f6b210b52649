"""The reference's own test modules, UNMODIFIED, against the drop-in (SURVEY.md section 7 stage 0 / VERDICT r1 missing 7).

oracle/build_ref.py vendors tests/test_{models,losses,transforms,trainer}.py of the reference into oracle/_ref/ref_tests
(git-ignored, travels to the GPU box); tests/ref_shim.py aliases `src.*` to phoneme_contrast_b200.*. On the reference itself
the four modules give 24 passed / 3 failed -- the three failures are NTXentLoss called WITHOUT labels, which the reference
leaves unimplemented (losses.py:153-159). The drop-in must reproduce exactly that outcome: same 24 passes, same 3
NotImplementedError failures."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "ref_tests")
MODULES = ("test_models.py", "test_losses.py", "test_transforms.py", "test_trainer.py")
EXPECTED_FAILURES = {"test_losses.py::TestNTXentLoss::test_loss_shape", "test_losses.py::TestNTXentLoss::test_perfect_positive_pairs",
                     "test_losses.py::TestNTXentLoss::test_temperature_effect"}


def test_reference_test_modules_run_unmodified_against_the_drop_in():
    if not all(os.path.exists(os.path.join(REF_TESTS, m)) for m in MODULES):
        pytest.skip("oracle/_ref/ref_tests not present (run oracle/build_ref.py where /root/reference is mounted)")
    cmd = [sys.executable, "-m", "pytest", "-p", "tests.ref_shim", "-p", "no:cacheprovider", "-q", "-rf", "--rootdir", REF_TESTS,
           "-c", os.devnull, *[os.path.join(REF_TESTS, m) for m in MODULES]]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env={**os.environ, "PYTHONPATH": ROOT})
    out = r.stdout + r.stderr
    failed = {re.sub(r"^.*ref_tests/", "", m) for m in re.findall(r"^FAILED (\S+)", out, flags=re.M)}
    m = re.search(r"(\d+) passed", out)
    passed = int(m.group(1)) if m else 0
    assert failed == EXPECTED_FAILURES, out[-4000:]
    assert passed == 24, out[-4000:]
    assert out.count("NotImplementedError: NT-Xent without labels not implemented in this version") >= 3, out[-4000:]
