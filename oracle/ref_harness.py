"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Loads the *unmodified* reference `fetal_net/prediction.py` (and `utils/patches.py`)
from /root/reference under a throw-away stub environment, so that
`patch_wise_prediction` (prediction.py:118-210), `get_set_of_patch_indices_full`
(prediction.py:88-95), `batch_iterator` (prediction.py:98-114) and
`get_patch_from_3d_data` (utils/patches.py:57-72) can be executed here and used
to (a) validate the NumPy restatement in `oracle/prediction_oracle.py` and
(b) generate the golden vectors frozen under `tests/golden/`.

/root/reference exists only in the build container; nothing that runs on the
GPU box (`-m gpu` tests, smoke(), bench.py) may import this module.

Stubs needed: Keras/TF/nibabel/tables/nilearn/skimage/imgaug/
SimpleITK are absent; `np.int`/`np.float` were removed from NumPy; and
`fetal_net.utils.utils.list_load` is imported by prediction.py:13 but never
defined by the reference.
"""
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("FETAL_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "fetal_net", "prediction.py"))


class _Anything(types.ModuleType):
    """A module whose every attribute is another permissive stub."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _Anything(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + "()")

    def __mro_entries__(self, bases):  # allows `class X(stub.Base)`
        return (object,)


_STUBBED = [
    "imp", "nibabel", "tables", "keras", "keras.backend", "keras.engine", "keras.engine.network",
    "keras.layers", "keras.layers.merge", "keras.optimizers", "keras.callbacks", "keras.models",
    "keras.losses", "keras.utils", "keras.initializers", "keras.regularizers", "keras.constraints",
    "keras.legacy", "keras.legacy.interfaces", "keras.utils.generic_utils", "keras.utils.conv_utils",
    "keras_contrib", "keras_contrib.layers", "keras_contrib.layers.normalization",
    "tensorflow", "nilearn", "nilearn.image", "nilearn.image.image", "nilearn.image.resampling",
    "skimage", "skimage.exposure", "skimage.transform", "skimage.util", "skimage.filters", "imgaug",
    "imgaug.augmenters", "SimpleITK", "sklearn.preprocessing.data", "matplotlib", "matplotlib.pyplot",
]

_loaded = None
_loaded_gen = None


def load_reference_generator():
    """Return the reference `fetal_net.generator` module (cached): data_generator / add_data / extract_patch
    (generator.py:222-348) run unmodified on NumPy arrays, used to freeze the training-sampler goldens."""
    global _loaded_gen
    if _loaded_gen is None:
        _loaded_gen = _load("fetal_net.generator")
    return _loaded_gen


def load_reference_prediction():
    """Return the reference `fetal_net.prediction` module (cached). Raises if absent."""
    global _loaded
    if _loaded is None:
        _loaded = _load("fetal_net.prediction")
    return _loaded


def _load(module_name):
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    saved_modules = dict(sys.modules)
    saved_path = list(sys.path)
    had_int, had_float = hasattr(np, "int"), hasattr(np, "float")
    try:
        for name in _STUBBED:
            if name not in sys.modules:
                sys.modules[name] = _Anything(name)
        # do NOT alias np.bool (breaks numpy.ma); int/float are what prediction.py/patches.py use
        if not had_int:
            np.int = int
        if not had_float:
            np.float = float
        # drop any already-imported `fetal_net`/`fetal` (ours) so the reference's resolves
        for name in [m for m in sys.modules if m == "fetal_net" or m.startswith("fetal_net.")
                     or m == "fetal" or m.startswith("fetal.")]:
            del sys.modules[name]
        sys.path.insert(0, REFERENCE_ROOT)
        # fetal_net/__init__ pulls in the whole Keras model zoo: stub the model package instead
        sys.modules["fetal_net.model"] = _Anything("fetal_net.model")
        utils_utils = importlib.import_module("fetal_net.utils.utils")
        if not hasattr(utils_utils, "list_load"):
            utils_utils.list_load = lambda *a, **k: []
        mod = importlib.import_module(module_name)
        patches = importlib.import_module("fetal_net.utils.patches")
        mod._ref_patches = patches
        return mod
    finally:
        # restore interpreter state: the reference modules stay reachable only via `_loaded`
        sys.path[:] = saved_path
        for name in list(sys.modules):
            if name not in saved_modules:
                del sys.modules[name]
        for name, mod in saved_modules.items():
            sys.modules[name] = mod


class FunctionModel:
    """Minimal Keras-Model duck type: what prediction.py:118-210,354-361 touches."""

    def __init__(self, fn, output_shape):
        self._fn = fn
        self.output_shape = tuple(output_shape)

    def predict(self, data):
        return self._fn(np.asarray(data))
