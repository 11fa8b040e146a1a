"""fetal_net.utils — `patches` is the B200-side mirror; the reference's other helpers (utils.py, threaded_generator.py,
sitk_utils.py ...) resolve through the overlay when FETAL_REFERENCE_ROOT is set (see fetal_net/__init__.py), and the
names the reference's own `fetal_net/utils/__init__.py` re-exports are forwarded lazily."""
from .. import reference_overlay_dir as _overlay

_d = _overlay("utils")
if _d is not None and _d not in __path__:
    __path__.append(_d)
del _d

# fetal_net/utils/__init__.py:1-2 of the reference
_FORWARDED = {"crop_img_to": "nilearn_custom_utils.nilearn_utils", "crop_img": "nilearn_custom_utils.nilearn_utils",
              "pickle_dump": "utils", "pickle_load": "utils", "read_image": "utils"}


def __getattr__(name):
    if name in _FORWARDED and _overlay("utils") is not None:
        import importlib
        return getattr(importlib.import_module(__name__ + "." + _FORWARDED[name]), name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
