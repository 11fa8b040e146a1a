"""Builders exported under the reference's names (fetal_net/model/__init__.py:3-18)."""
from .unet3d import unet_model_3d, unet_model_2d, isensee2017_model_3d, isensee2017_model, Model  # noqa: F401
from .. import reference_overlay_dir as _overlay

# overlay: sub-modules this package does not define (fetal_net.model.fetal_net, .discriminator, .norm ... - Keras
# models outside the hot path) resolve to the reference checkout when FETAL_REFERENCE_ROOT is set
_d = _overlay("model")
if _d is not None and _d not in __path__:
    __path__.append(_d)
del _d


def _not_built(name, why, ref_module=None):
    """A builder outside the B200 hot path: forwarded to the reference's own (Keras) builder when the overlay is
    active and Keras is importable there, otherwise a clear NotImplementedError."""
    def fn(*a, **k):
        if ref_module is not None and _overlay("model") is not None:
            import importlib
            try:
                mod = importlib.import_module(__name__ + "." + ref_module)
            except ImportError as e:
                raise NotImplementedError("%s: %s (and the reference's own builder could not be imported: %s)"
                                          % (name, why, e)) from e
            return getattr(mod, name)(*a, **k)
        raise NotImplementedError("%s: %s" % (name, why))
    fn.__name__ = name
    return fn


_OUT = "outside the B200 hot path (SURVEY.md §2: classifier / adversarial models are out of scope)"
fetal_envelope_model = _not_built("fetal_envelope_model", _OUT, "fetal_net")
fetal_origin_model = _not_built("fetal_origin_model", _OUT, "fetal_net_skip")
fetal_origin2_model = _not_built("fetal_origin2_model", _OUT, "fetal_net_skip2")
fetal_origin3_model = _not_built("fetal_origin3_model", _OUT, "fetal_net_skip3")
norm_net_model = _not_built("norm_net_model", _OUT, "norm.NormNet")
discriminator_image_2d = _not_built("discriminator_image_2d", _OUT, "discriminator.discriminator_image_2d")
discriminator_image_3d = _not_built("discriminator_image_3d", _OUT, "discriminator.discriminator_image_3d")
