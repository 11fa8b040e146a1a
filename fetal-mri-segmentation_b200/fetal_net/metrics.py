"""Loss / metric names of the reference (fetal_net/metrics.py), as looked up by
`getattr(fetal_net.metrics, config['loss'])` (fetal/train_fetal.py:31).

The callables evaluate on host NumPy arrays (they are identifiers + evaluation helpers); the
training loss itself is computed on the device by libfetalb200 (fm_train_step) — the builders
compare `loss_function` by identity with `dice_coefficient_loss`, as the reference does
(unet3d/unet.py:82-83), and only the soft-Dice loss is built on the device path.
"""
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


# names the reference exports (metrics.py:97-100, config_utils.py:73-79)
dice_coef = dice_coefficient
dice_coef_loss = dice_coefficient_loss
binary_crossentropy_loss = _not_built("binary_crossentropy_loss")
focal_loss = _not_built("focal_loss")
dice_and_xent = _not_built("dice_and_xent")
dice_and_xent_mask = _not_built("dice_and_xent_mask")
double_dice_loss = _not_built("double_dice_loss")
