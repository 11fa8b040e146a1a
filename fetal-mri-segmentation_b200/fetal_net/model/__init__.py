"""Builders exported under the reference's names (fetal_net/model/__init__.py:3-18)."""
from .unet3d import unet_model_3d, unet_model_2d, isensee2017_model_3d, Model  # noqa: F401


def _not_built(name, why):
    def fn(*a, **k):
        raise NotImplementedError("%s: %s" % (name, why))
    fn.__name__ = name
    return fn


_NEXT = "on the §8 'next' list of SURVEY.md — not built yet in the B200 path"
_OUT = "outside the B200 hot path (SURVEY.md §2: classifier / adversarial models are out of scope)"
isensee2017_model = _not_built("isensee2017_model", _NEXT)
fetal_envelope_model = _not_built("fetal_envelope_model", _OUT)
fetal_origin_model = _not_built("fetal_origin_model", _OUT)
fetal_origin2_model = _not_built("fetal_origin2_model", _OUT)
fetal_origin3_model = _not_built("fetal_origin3_model", _OUT)
norm_net_model = _not_built("norm_net_model", _OUT)
discriminator_image_2d = _not_built("discriminator_image_2d", _OUT)
discriminator_image_3d = _not_built("discriminator_image_3d", _OUT)
