"""Host mirror of the reference's patch slicing (fetal_net/utils/patches.py:57-91) — API compatibility
for callers that slice on the host; the sliding-window path itself gathers on the device
(fm_gather_patches / fm_patchwise_predict)."""
import numpy as np


def get_patch_from_3d_data(data, patch_shape, patch_index):
    patch_index = np.asarray(patch_index, dtype=np.int16)     # patches.py:65
    patch_shape = np.asarray(patch_shape)
    image_shape = data.shape[-3:]
    if np.any(patch_index < 0) or np.any((patch_index + patch_shape) > image_shape):
        data, patch_index = fix_out_of_bound_patch_attempt(data, patch_shape, patch_index)
    return data[..., patch_index[0]:patch_index[0] + patch_shape[0],
                patch_index[1]:patch_index[1] + patch_shape[1],
                patch_index[2]:patch_index[2] + patch_shape[2]]


def fix_out_of_bound_patch_attempt(data, patch_shape, patch_index, ndim=3):
    image_shape = data.shape[-ndim:]
    before = np.abs((patch_index < 0) * patch_index)
    after = np.abs(((patch_index + patch_shape) > image_shape) * ((patch_index + patch_shape) - image_shape))
    pads = [[0, 0]] * (data.ndim - ndim) + np.stack([before, after], axis=1).tolist()
    return np.pad(data, pads, mode="edge"), patch_index + before
