"""fetal_net — B200-native drop-in for the hot path of GalDude33/Fetal-MRI-Segmentation.

Mirrors the reference package layout for the path it replaces:
  fetal_net.model       builders looked up by `getattr(fetal_net.model, config['model_name'])`
                        (fetal/train_fetal.py:31-32)
  fetal_net.metrics     losses looked up by `getattr(fetal_net.metrics, config['loss'])`
  fetal_net.prediction  patch_wise_prediction and friends (fetal_net/prediction.py)
  fetal_net.training    train_model / load_old_model (fetal_net/training.py)
Numerics run in libfetalb200.so (hand-written sm_100a CUDA behind the C ABI of include/fetal_b200.h).

OVERLAY. This package replaces only the hot path. Everything else of the reference's `fetal_net` (data.py,
generator.py, augment.py, normalize.py, preprocess.py, postprocess.py, utils/utils.py, model/fetal_net*.py ...) is NOT
re-implemented: when FETAL_REFERENCE_ROOT points at a checkout of the reference, its `fetal_net` directory is appended
to this package's search path (and `fetal_net/model`, `fetal_net/utils` to the sub-packages'), so
`import fetal_net.generator`, `from fetal_net.data import open_data_file` or
`from fetal_net.model.fetal_net import fetal_envelope_model` (fetal/train_fetal.py:5-13) resolve to the reference's own
files, while every module defined HERE shadows its reference namesake. With this directory first on PYTHONPATH the
reference's scripts (`python -m fetal.train_fetal`, `python -m fetal.predict`) run unchanged on the B200 path.
"""
import os as _os

__all__ = ["model", "metrics", "prediction", "training"]


def reference_overlay_dir(*parts):
    """<FETAL_REFERENCE_ROOT>/fetal_net/<parts...> if that directory exists, else None."""
    root = _os.environ.get("FETAL_REFERENCE_ROOT")
    if not root:
        return None
    d = _os.path.join(root, "fetal_net", *parts)
    return d if _os.path.isdir(d) else None


_d = reference_overlay_dir()
if _d is not None and _d not in __path__:
    __path__.append(_d)
del _d
