"""Recipe: vendor the reference's own hot-path Python modules into oracle/_ref/ (TEST / BASELINE INFRASTRUCTURE).

    python oracle/build_ref.py            # run in the build container, where /root/reference is mounted

The reference (brant01/phoneme_contrast) is pure Python with no build step; what "building" it means here is copying the
package directories its hot path imports -- src/{datasets,models,training,utils} -- and the four test modules that pin
the public contracts (tests/test_{models,losses,transforms,trainer}.py) from where they lie under /root/reference into
oracle/_ref/, unmodified. oracle/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored, so
it travels to the GPU box, where /root/reference does not exist. Consumers, and only these:
  * bench.py --impl reference  and bench.py's cpu_baseline leg: time the reference's OWN modules (kind = "reference") on
    the host cores instead of the oracle port;
  * tests/test_reference_suite.py: runs the reference's unmodified test modules against the drop-in through a `src` alias.
The product package never imports anything from here (tests/test_abi_cpu.py checks).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC_ROOT = os.environ.get("PC_REFERENCE_ROOT", "/root/reference")
PACKAGES = ("datasets", "models", "training", "utils")
TESTS = ("test_models.py", "test_losses.py", "test_transforms.py", "test_trainer.py")


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(SRC_ROOT, "src")):
        if verbose:
            print(f"oracle/_ref: {SRC_ROOT} not present; keeping the existing copy" if os.path.isdir(DST) else
                  f"oracle/_ref: {SRC_ROOT} not present and no copy exists (bench falls back to the oracle port)")
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.join(DST, "src"))
    shutil.copy2(os.path.join(SRC_ROOT, "src", "__init__.py"), os.path.join(DST, "src", "__init__.py"))
    n = 1
    for pkg in PACKAGES:
        for dirpath, _dirs, files in os.walk(os.path.join(SRC_ROOT, "src", pkg)):
            rel = os.path.relpath(dirpath, SRC_ROOT)
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            for f in files:
                if f.endswith(".py"):
                    shutil.copy2(os.path.join(dirpath, f), os.path.join(DST, rel, f))
                    n += 1
    os.makedirs(os.path.join(DST, "ref_tests"))
    for t in TESTS:
        shutil.copy2(os.path.join(SRC_ROOT, "tests", t), os.path.join(DST, "ref_tests", t))
        n += 1
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write(f"unmodified copies from {SRC_ROOT} made by oracle/build_ref.py; git-ignored; not product code\n")
    if verbose:
        print(f"oracle/_ref: vendored {n} reference files from {SRC_ROOT}")
    return True


def import_reference():
    """Put oracle/_ref first on sys.path and return True when the reference's `src` package is importable from it."""
    if not os.path.isdir(os.path.join(DST, "src")):
        return False
    for name in [m for m in sys.modules if m == "src" or m.startswith("src.")]:
        del sys.modules[name]
    if DST not in sys.path:
        sys.path.insert(0, DST)
    try:
        import src.models  # noqa: F401
        import src.training.losses  # noqa: F401
        import src.datasets.features  # noqa: F401
        return True
    except Exception as exc:   # missing optional dependency on this box
        print(f"oracle/_ref present but not importable: {exc!r}", file=sys.stderr)
        return False


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
