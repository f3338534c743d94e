"""TEST INFRASTRUCTURE ONLY -- import shims that let the *unmodified* reference
(/root/reference) be imported in this container, so the oracle restatement and the
golden vectors under tests/golden/ can be pinned against it.

Nothing under moleculardiffusion_mivit_b200/ may import this module.  It is used by
oracle/make_golden.py (fixture generation, run here where /root/reference exists)
and by `-m "not gpu"` tests that skip themselves when the reference is absent (the
GPU box has no /root/reference).

Two third-party imports of helpers/helpersGeneration.py are not installed and there
is no network:
  * andi_datasets.models_phenom (helpersGeneration.py:4)  -- only *called* in the
    unused generateTrajAndVideosBrownian (:403); a stub class is enough.
  * skimage.measure.block_reduce (helpersGeneration.py:5) -- block mean; restated as
    reshape -> func(axis=(1,3)).  skimage.filters.gaussian is only used by the
    out-of-scope denoising variant (:530).
No reference file is copied or modified.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MIVIT_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "helpers"))


def _block_reduce(image, block_size, func, cval=0, func_kwargs=None):
    import numpy as np
    b = int(block_size)
    h, w = image.shape
    assert h % b == 0 and w % b == 0
    return func(image.reshape(h // b, b, w // b, b), axis=(1, 3))


def install_shims():
    if "andi_datasets.models_phenom" not in sys.modules:
        pkg = types.ModuleType("andi_datasets")
        sub = types.ModuleType("andi_datasets.models_phenom")

        class models_phenom:  # noqa: N801 (reference name)
            def single_state(self, *a, **k):
                raise RuntimeError("andi_datasets is not installed (shim)")

        sub.models_phenom = models_phenom
        pkg.models_phenom = sub
        sys.modules["andi_datasets"] = pkg
        sys.modules["andi_datasets.models_phenom"] = sub
    if "skimage" not in sys.modules:
        ski = types.ModuleType("skimage")
        meas = types.ModuleType("skimage.measure")
        meas.block_reduce = _block_reduce
        filt = types.ModuleType("skimage.filters")

        def _gaussian(image, sigma=1, *, mode="nearest", cval=0, preserve_range=False, truncate=4.0, channel_axis=None, **k):
            """skimage.filters.gaussian (absent here, no version pinned by the reference) for a float image: it forwards to
            scipy.ndimage.gaussian_filter with mode='nearest', truncate=4.0 and an output of the image's own float dtype."""
            import numpy as np
            from scipy import ndimage as ndi
            img = np.asarray(image)
            if img.dtype.kind != "f":
                img = img.astype(np.float64)
            out = np.empty_like(img)
            ndi.gaussian_filter(img, sigma, output=out, mode=mode, cval=cval, truncate=truncate)
            return out

        filt.gaussian = _gaussian
        ski.measure = meas
        ski.filters = filt
        sys.modules["skimage"] = ski
        sys.modules["skimage.measure"] = meas
        sys.modules["skimage.filters"] = filt


def import_reference():
    """Returns (helpersGeneration, models) modules of the unmodified reference."""
    if not reference_available():
        raise ImportError("reference not present at %s" % REFERENCE_ROOT)
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    gen = importlib.import_module("helpers.helpersGeneration")
    models = importlib.import_module("helpers.models")
    return gen, models


def import_experiment_settings(name):
    """Import Experiments/<name>/trainSettings<name>.py unmodified (e.g. 'PSFNoise')."""
    import importlib.util
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    path = os.path.join(REFERENCE_ROOT, "Experiments", name, "trainSettings%s.py" % name)
    spec = importlib.util.spec_from_file_location("ref_trainSettings" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
