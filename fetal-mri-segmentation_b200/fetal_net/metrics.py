"""Loss / metric names of the reference (fetal_net/metrics.py), as looked up by
`getattr(fetal_net.metrics, config['loss'])` (fetal/train_fetal.py:31).

The callables evaluate on host NumPy arrays (they are identifiers + evaluation helpers); the
training loss itself is computed on the device by libfetalb200 (fm_train_step) — the builders
compare `loss_function` by identity with `dice_coefficient_loss`, as the reference does
(unet3d/unet.py:82-83). Built on the device path: `dice_coefficient_loss`, `dice_and_xent` (also as a
`functools.partial` with `xent_weight`) and `dice_and_xent_mask` (the closure factory that takes the weight-mask
input, isensee2017.py:85-88); see `device_loss_spec`.
"""
import functools

import numpy as np


def dice_coefficient(y_true, y_pred, smooth=1.):
    # metrics.py:11-15 — flattens everything including the batch axis
    t = np.asarray(y_true, dtype=np.float64).ravel()
    p = np.asarray(y_pred, dtype=np.float64).ravel()
    return (2. * np.sum(t * p) + smooth) / (np.sum(t) + np.sum(p) + smooth)


def dice_coefficient_loss(y_true, y_pred):
    # metrics.py:31-32
    return -dice_coefficient(y_true, y_pred)


def vod_coefficient(y_true, y_pred, binarize=True, smooth=1.):
    # metrics.py:18-28
    t = np.asarray(y_true, dtype=np.float64).ravel()
    p = np.asarray(y_pred, dtype=np.float64).ravel()
    if binarize:
        t = (t > 0.5).astype(np.float64)
        p = (p > 0.5).astype(np.float64)
    inter = np.sum(t * p)
    union = np.sum(t) + np.sum(p) - inter
    return (inter + smooth) / (union + smooth)


def vod_coefficient_loss(y_true, y_pred):
    return -vod_coefficient(y_true, y_pred, binarize=False)


def binary_accuracy(y_true, y_pred):
    # Keras: mean(equal(y_true, round(y_pred)))
    return float(np.mean(np.asarray(y_true) == np.round(np.asarray(y_pred))))


def weighted_dice_coefficient(y_true, y_pred, axis=(-3, -2, -1), smooth=0.00001):
    # metrics.py:39-52 (host evaluation only)
    t = np.asarray(y_true, dtype=np.float64)
    p = np.asarray(y_pred, dtype=np.float64)
    return float(np.mean(2. * (np.sum(t * p, axis=axis) + smooth / 2) /
                         (np.sum(t, axis=axis) + np.sum(p, axis=axis) + smooth)))


def weighted_dice_coefficient_loss(y_true, y_pred):
    return -weighted_dice_coefficient(y_true, y_pred)


def _not_built(name):
    def fn(*a, **k):
        raise NotImplementedError(
            "%s is outside the B200 hot path (SURVEY.md §8: optional losses are not in BASELINE's configs); "
            "only dice_coefficient_loss trains on the device" % name)
    fn.__name__ = name
    return fn


def binary_crossentropy(y_true, y_pred):
    # Keras K.binary_crossentropy (TF backend) on probabilities: clip to [eps, 1 - eps] with both bounds formed in
    # float32 (1 - eps = 1 - 2^-23), elementwise
    t = np.asarray(y_true, dtype=np.float64)
    p = np.clip(np.asarray(y_pred, dtype=np.float64), float(np.float32(1e-7)), float(np.float32(1.) - np.float32(1e-7)))
    return -(t * np.log(p) + (1. - t) * np.log1p(-p))


def weighted_cross_entropy_loss(y_true, y_pred, weight_mask=None):
    # metrics.py:73-77
    xent = binary_crossentropy(y_true, y_pred)
    if weight_mask is not None:
        xent = np.asarray(weight_mask, dtype=np.float64).reshape(xent.shape) * xent
    return float(np.mean(xent))


def dice_and_xent(y_true, y_pred, xent_weight=1.0, weight_mask=None):
    # metrics.py:68-70
    return dice_coefficient_loss(y_true, y_pred) + xent_weight * weighted_cross_entropy_loss(y_true, y_pred, weight_mask)


def dice_and_xent_mask(weight_mask, xent_weight=1.0, dist_sigma=3):
    """metrics.py:89-95: a loss closure over the weight-mask input. With `weight_mask` an array the closure evaluates on
    the host; the builders call it with the mask INPUT placeholder (isensee2017.py:85-88) - here `None` - and only read
    its parameters (`device_loss_spec`): on the device the mask arrives per step as the second input."""
    def _loss(y_true, y_pred):
        w = None if weight_mask is None else np.exp(-np.asarray(weight_mask, dtype=np.float64) / dist_sigma)
        return dice_and_xent(y_true, y_pred, xent_weight=xent_weight, weight_mask=w)
    _loss.xent_weight, _loss.dist_sigma, _loss.is_dice_and_xent_mask = float(xent_weight), float(dist_sigma), True
    return _loss


def device_loss_spec(loss_function, has_mask_input=False):
    """-> (kind, xent_weight, dist_sigma) for fm_model_set_loss, or None when the loss is not built on the device."""
    if loss_function is dice_coefficient_loss:
        return 0, 0.0, 0.0
    if loss_function is dice_and_xent:
        return 1, 1.0, 0.0
    if isinstance(loss_function, functools.partial) and loss_function.func is dice_and_xent and not loss_function.args \
            and set(loss_function.keywords) <= {"xent_weight"}:
        return 1, float(loss_function.keywords.get("xent_weight", 1.0)), 0.0
    if loss_function is dice_and_xent_mask and has_mask_input:
        return 2, 1.0, 3.0                                   # the factory's defaults, as isensee2017.py:88 calls it
    if getattr(loss_function, "is_dice_and_xent_mask", False) and has_mask_input:
        return 2, loss_function.xent_weight, loss_function.dist_sigma
    return None


# names the reference exports (metrics.py:97-100, config_utils.py:73-79)
dice_coef = dice_coefficient
dice_coef_loss = dice_coefficient_loss
binary_crossentropy_loss = _not_built("binary_crossentropy_loss")
focal_loss = _not_built("focal_loss")
double_dice_loss = _not_built("double_dice_loss")
