"""TEST INFRASTRUCTURE ONLY -- import shims that let the *unmodified* reference run in this container.

The reference (/root/reference, read-only) imports two packages that are not installed here:
`kornia` (models.py:6, lossfunction.py:5, data.py:10) and `skimage.draw.random_shapes` (misc.py:8).
Only `kornia.normalize` is on the hot path (models.py:195-197).  This module registers minimal stand-ins in
`sys.modules` and returns the imported reference modules.  It is used by tests/golden/make_golden.py and by
the `-m "not gpu"` tests that pin oracle/spyramid_oracle.py against the reference; it is never imported by
the product package and never runs on the GPU box (/root/reference does not exist there).
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SPYR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models.py"))


def _install_stubs() -> None:
    import torch

    if "kornia" not in sys.modules:
        kornia = types.ModuleType("kornia")

        def normalize(x, mean, std):
            # kornia.normalize: per-channel (x - mean) / std with 3-vectors (reference models.py:195-197)
            return (x - mean.view(1, -1, 1, 1)) / std.view(1, -1, 1, 1)

        def normalize_min_max(x, min_val=0.0, max_val=1.0, eps=1e-6):
            # only used by data.py:53 (not on the hot path); per-(B,C) min-max rescale
            b, c = x.shape[:2]
            flat = x.reshape(b, c, -1)
            lo = flat.min(dim=-1)[0].view(b, c, 1, 1)
            hi = flat.max(dim=-1)[0].view(b, c, 1, 1)
            return (max_val - min_val) * (x - lo) / (hi - lo + eps) + min_val

        kornia.normalize = normalize
        kornia.normalize_min_max = normalize_min_max
        sys.modules["kornia"] = kornia
    if "skimage" not in sys.modules:
        skimage = types.ModuleType("skimage")
        draw = types.ModuleType("skimage.draw")

        def random_shapes(*args, **kwargs):
            raise RuntimeError("skimage is not installed; supply the uint8 shape image explicitly")

        draw.random_shapes = random_shapes
        skimage.draw = draw
        sys.modules["skimage"] = skimage
        sys.modules["skimage.draw"] = draw
    _ = torch


def import_reference():
    """Returns (models, lossfunction, misc) imported from the reference tree."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference's module names collide with nothing in this repo's top level
    models = importlib.import_module("models")
    lossfunction = importlib.import_module("lossfunction")
    misc = importlib.import_module("misc")
    return models, lossfunction, misc
