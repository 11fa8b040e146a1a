"""Host-side patch slicing with the reference's entry points (fetal_net/utils/patches.py:57-91), kept for callers
that slice on the host; the sliding-window path itself gathers on the device (fm_gather_patches /
fm_patchwise_predict).

Behaviour mirrored: the corner is cast to int16 first (patches.py:65); a patch that sticks out of the last three
axes is completed with the nearest edge sample (`np.pad(mode='edge')` in the reference). Here that is done by
clamping the per-axis index ranges, which yields the same array without building the padded copy."""
import numpy as np


def _axis_indices(corner, extent, size):
    """Indices of one patch axis with out-of-range positions clamped to the border (= edge padding)."""
    return np.clip(np.arange(int(corner), int(corner) + int(extent)), 0, size - 1)


def get_patch_from_3d_data(data, patch_shape, patch_index):
    corner = np.asarray(patch_index, dtype=np.int16)
    out = np.asarray(data)
    first = out.ndim - 3
    for a in range(3):
        size = out.shape[first + a]
        c, p = int(corner[a]), int(patch_shape[a])
        if c >= 0 and c + p <= size:
            out = out[(slice(None),) * (first + a) + (slice(c, c + p),)]        # plain view, like the reference
        else:
            out = np.take(out, _axis_indices(c, p, size), axis=first + a)
    return out


def fix_out_of_bound_patch_attempt(data, patch_shape, patch_index, ndim=3):
    """Returns (edge-padded data, corner shifted into it) for a patch that leaves the last `ndim` axes."""
    corner = np.asarray(patch_index)
    extent = np.asarray(patch_shape)
    sizes = np.asarray(data.shape[-ndim:])
    low = np.maximum(-corner, 0)
    high = np.maximum(corner + extent - sizes, 0)
    widths = [(0, 0)] * (data.ndim - ndim) + [(int(l), int(h)) for l, h in zip(low, high)]
    return np.pad(data, widths, mode="edge"), corner + low
