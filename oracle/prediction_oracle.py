"""TEST INFRASTRUCTURE ONLY (the product path must never import `oracle/`).

NumPy restatement of the reference's sliding-window inference:

  * `patch_plan`            <- get_set_of_patch_indices_full      fetal_net/prediction.py:88-95
                               + overlap arithmetic                fetal_net/prediction.py:129-137,161-163
  * `pad_volume`            <- the two np.pad calls                fetal_net/prediction.py:138-146 (truth: 148-154)
  * `extract_patch`         <- get_patch_from_3d_data              fetal_net/utils/patches.py:57-72
                               fix_out_of_bound_patch_attempt      fetal_net/utils/patches.py:75-91
  * `patch_wise_prediction` <- patch_wise_prediction               fetal_net/prediction.py:118-210
                               batch_iterator                      fetal_net/prediction.py:98-114

Pinned: `tests/test_oracle_pinning.py` runs this file against the unmodified
reference code (oracle/ref_harness.py) where /root/reference exists, and against
the goldens frozen from it in tests/golden/ everywhere else.
"""
import itertools

import numpy as np


def is_3d_model(output_shape):
    # prediction.py:129
    return int(np.sum(np.array(output_shape[1:]) > 1)) > 2


def prediction_shape_of(output_shape):
    # prediction.py:131-134
    if is_3d_model(output_shape):
        return tuple(output_shape[-3:])
    return tuple(output_shape[-3:-1]) + (1,)


def axis_starts(stop, step):
    # prediction.py:91-94 (start is always 0 in the live path)
    starts = list(range(0, stop + 1, step))
    if stop % step > 0:
        starts.append(stop)
    return starts


def compute_overlap(patch_shape, prediction_shape, overlap_factor):
    # prediction.py:135-137 (astype(np.int) == truncation toward zero)
    min_overlap = np.subtract(patch_shape, prediction_shape)
    max_overlap = np.subtract(patch_shape, (1, 1, 1))
    return min_overlap + (overlap_factor * (max_overlap - min_overlap)).astype(int)


def halo_pad(patch_shape, prediction_shape):
    # prediction.py:139-140
    return [(int(np.ceil(d / 2)), int(np.floor(d / 2))) for d in np.subtract(patch_shape, prediction_shape)]


def fit_pad(patch_shape, padded_shape):
    # prediction.py:142-143
    return [(int(np.ceil(d / 2)), int(np.floor(d / 2)))
            for d in np.maximum(np.subtract(patch_shape, padded_shape), 0)]


def pad_volume(data0, patch_shape, prediction_shape, truth0=None):
    """Returns (padded data, pad_for_fit, padded truth or None). prediction.py:138-154."""
    hp = halo_pad(patch_shape, prediction_shape)
    d = np.pad(data0, hp, mode="constant", constant_values=np.percentile(data0, q=1))
    pf = fit_pad(patch_shape, d.shape)
    d = np.pad(d, pf, "constant", constant_values=np.percentile(d, q=1))
    t = None
    if truth0 is not None:
        t = np.pad(truth0, hp, mode="constant", constant_values=0)
        t = np.pad(t, pf, "constant", constant_values=0)
    return d, pf, t


def patch_plan(padded_shape, patch_shape, prediction_shape, overlap_factor):
    """(n,3) int array of patch corners, x-major product order. prediction.py:88-95,161-163."""
    overlap = compute_overlap(patch_shape, prediction_shape, overlap_factor)
    stop = np.subtract(padded_shape, patch_shape)
    step = np.subtract(patch_shape, overlap)
    per_axis = [axis_starts(int(s), int(st)) for s, st in zip(stop, step)]
    return np.array(list(itertools.product(*per_axis)))


def extract_patch(data, patch_shape, patch_index):
    # patches.py:57-91. patch_index is cast to int16 by the reference (patches.py:65).
    patch_index = np.asarray(patch_index, dtype=np.int16)
    patch_shape = np.asarray(patch_shape)
    image_shape = data.shape[-3:]
    if np.any(patch_index < 0) or np.any((patch_index + patch_shape) > image_shape):
        pad_before = np.abs((patch_index < 0) * patch_index)
        pad_after = np.abs(((patch_index + patch_shape) > image_shape) * ((patch_index + patch_shape) - image_shape))
        pad_args = np.stack([pad_before, pad_after], axis=1).tolist()
        pad_args = [[0, 0]] * (data.ndim - 3) + pad_args
        data = np.pad(data, pad_args, mode="edge")
        patch_index = patch_index + pad_before
    return data[..., patch_index[0]:patch_index[0] + patch_shape[0],
                patch_index[1]:patch_index[1] + patch_shape[1],
                patch_index[2]:patch_index[2] + patch_shape[2]]


def count_map(out_shape3, pred_shape, indices):
    cnt = np.zeros(tuple(out_shape3), dtype=np.int16)
    px, py, pz = pred_shape
    for x, y, z in indices:
        cnt[x:x + px, y:y + py, z:z + pz] += 1
    return cnt


def patch_wise_prediction(model, data, patch_shape, overlap_factor=0, batch_size=5,
                          truth_data=None, prev_truth_index=None, prev_truth_size=None):
    """prediction.py:118-210 without the producer thread / tqdm (both value-neutral)."""
    output_shape = model.output_shape
    is3d = is_3d_model(output_shape)
    pred_shape = prediction_shape_of(output_shape)
    data0, pad_for_fit, truth0 = pad_volume(data[0], patch_shape, pred_shape,
                                            None if truth_data is None else truth_data[0])
    indices = patch_plan(data0.shape, patch_shape, pred_shape, overlap_factor)
    truth_patch_shape = None if truth0 is None else list(patch_shape[:2]) + [prev_truth_size]

    data_shape = list(np.asarray(data.shape[-3:]) + np.sum(pad_for_fit, -1))
    data_shape += [output_shape[1]] if is3d else [output_shape[-1]]
    out = np.zeros(data_shape)
    cnt = np.zeros(data_shape, dtype=np.int16)

    for b0 in range(0, len(indices), batch_size):
        idx = indices[b0:b0 + batch_size]
        batch = []
        for ci in idx:
            p = extract_patch(data0, patch_shape, ci)
            if truth0 is not None:
                ti = list(ci[:2]) + [ci[2] + prev_truth_index]
                p = np.concatenate([p, extract_patch(truth0, truth_patch_shape, ti)], axis=-1)
            batch.append(p)
        batch = np.asarray(batch)
        if is3d:
            batch = np.expand_dims(batch, 1)
        pred = model.predict(batch)
        pred = pred.transpose([0, 2, 3, 4, 1]) if is3d else np.expand_dims(pred, -2)
        for pp, (x, y, z) in zip(pred, idx):
            xl, yl, zl = pp.shape[:-1]
            out[x:x + xl, y:y + yl, z:z + zl, :] += pp
            cnt[x:x + xl, y:y + yl, z:z + zl] += 1

    assert np.all(cnt > 0), 'Found zeros in count'
    if np.sum(pad_for_fit) > 0:
        sl = tuple(slice(p[0] or None, -p[1] if p[1] else None) for p in pad_for_fit)
        out, cnt = out[sl], cnt[sl]
    assert np.array_equal(cnt.shape[:-1], data[0].shape), 'prediction shape wrong'
    return out / cnt
