"""Arithmetic arm of the convolution kernels: one place that decides it.

"bf16x3" (default, the product): tcgen05 tensor cores, 3-term bf16 split, fp32 accumulation -- fp32-class results.
"fp32": exact fp32 on CUDA cores (parity arm / on-device ground truth).  "bf16": single-pass tensor cores (triage).

Every conv-running module carries a `precision` attribute (None = follow the process default); a detector sets all
of its modules at once with `VoxelNet.set_precision()` or through a `precision` key in `test_cfg` / `train_cfg`.
"""
import contextlib

VALID = ("bf16x3", "fp32", "bf16")
_default = "bf16x3"


def check(p):
    if p not in VALID:
        raise ValueError("precision must be one of %s (got %r)" % (VALID, p))
    return p


def default_precision():
    return _default


def set_default_precision(p):
    """Process-wide default for modules whose own `precision` is None."""
    global _default
    _default = check(p)


def resolve(p=None):
    return check(p) if p is not None else _default


def act_fmt(p=None):
    """Inter-layer activation format of a precision: fp32 rows, or split bf16 hi/lo rows on the tensor-core arms."""
    return "fp32" if resolve(p) == "fp32" else "split"


@contextlib.contextmanager
def use_precision(p):
    """with use_precision("fp32"): ...   (tests / A-B comparisons)"""
    global _default
    old = _default
    _default = check(p)
    try:
        yield
    finally:
        _default = old


def set_module_precision(root, p):
    """Set `precision` on every module below `root` that has one (None restores 'follow the default')."""
    if p is not None:
        check(p)
    for m in root.modules():
        if hasattr(m, "precision"):
            m.precision = p
    return root
