"""Training batches cut on the device.

Device-resident counterpart of the reference's training generator (`data_generator` -> `add_data` -> `extract_patch` ->
`get_patch_from_3d_data`, `convert_data`; fetal_net/generator.py:222-348,380-401) for the `augment=None` path: the cases
are uploaded once (`fm_volset_set_case`), the HOST keeps making every random decision with the reference's own call
sequence on the global `random` / `np.random` generators -

    index  = next(index_generator)                  # random.sample per epoch (np.random.seed() first), or in order
    corner = [np.random.randint(0, high) for high in truth.shape - patch_shape]          # generator.py:266-269
    drop_easy_patches: np.random.random() against 1 - |mean(truth[16:-16, 16:-16, :]) - 0.5|      # :307-310
    skip_blank: keep the sample only if np.any(truth != 0)                                        # :323

- and the device cuts the whole batch with one kernel (`sample_patches_kernel`, csrc/sampler.cu): the data patch, the
target slice(s) at `truth_index`, the previous-truth slice(s) at `prev_truth_index` appended as extra input channels
(generator.py:305-306), slices that stick out of the volume completed with the nearest edge sample
(utils/patches.py:75-91). Same seeds => the same (case, corner) sequence and bit-identical float32 batches as the
reference generator (tests/golden/sampler_golden.npz is frozen from the unmodified reference).

`DeviceSampler` is an iterator yielding `(x, y)` like the reference generator (so `fit_generator` takes it unchanged);
`train_on_next_batch(model)` skips the host round trip altogether (`fm_train_step_sampled`: the batch goes straight
into the model's input buffers; 32 bytes of arguments per sample cross the bus).

Not covered here, by design: the reference's nilearn / imgaug augmentation pipeline (`augment_data`,
generator.py:271-295), `truth_downsample`, masks and `categorical=True` - those stay on the reference's host generator
(reachable through the overlay, see fetal_net/__init__.py). The cheap per-sample augmentations the kernel offers (axis
flips, intensity scale, additive Gaussian noise) are an extension, off by default.
"""
import ctypes
import random

import numpy as np

from . import _lib


def random_list_generator(index_list):
    """generator.py:201-204, including its np.random.seed() at the start of every pass."""
    while True:
        np.random.seed()
        yield from random.sample(index_list, len(index_list))


def list_generator(index_list):
    """generator.py:207-209."""
    while True:
        yield from index_list


class PatchDraws:
    """The host half: which (case, corner) samples make up the next batch. No device needed."""

    def __init__(self, truth_list, index_list, batch_size=1, patch_shape=None, shuffle_index_list=True,
                 skip_blank=True, truth_index=-1, truth_size=1, prev_truth_index=None, prev_truth_size=None,
                 drop_easy_patches=False, is3d=False, categorical=False, augment=None, truth_downsample=None):
        if augment is not None:
            raise NotImplementedError("DeviceSampler covers the augment=None path; the nilearn/imgaug pipeline of "
                                      "generator.py:271-295 stays on the reference's host generator")
        if categorical:
            raise NotImplementedError("categorical=True (to_categorical targets) is not part of the soft-Dice path")
        if truth_downsample is not None and truth_downsample > 1:
            raise NotImplementedError("truth_downsample is not supported by the device sampler")
        if patch_shape is None:
            raise ValueError("patch_shape is required")
        self.patch_shape = tuple(int(v) for v in patch_shape)
        self.batch_size = int(batch_size)
        self.skip_blank = bool(skip_blank)
        self.truth_index, self.truth_size = int(truth_index), int(truth_size)
        self.prev_truth_index = None if prev_truth_index is None else int(prev_truth_index)
        self.prev_truth_size = 0 if prev_truth_index is None else int(1 if prev_truth_size is None else prev_truth_size)
        self.drop_easy_patches = bool(drop_easy_patches)
        self.is3d = bool(is3d)
        self._index_generator = (random_list_generator if shuffle_index_list else list_generator)(list(index_list))
        # host copies of the truth volumes: the skip_blank / drop_easy_patches decisions are host decisions
        self._truth = [np.asarray(t) for t in truth_list]
        self._shapes = [t.shape[-3:] for t in self._truth]

    # ---- the host decisions, in the reference's order --------------------------------------------------------------
    def _truth_patch(self, case, corner, index, size):
        """get_patch_from_3d_data(truth, patch_shape[:-1] + (size,), corner + (0, 0, index)) by clamped indices."""
        t = self._truth[case]
        lo = (corner[0], corner[1], corner[2] + index)
        hi = (lo[0] + self.patch_shape[0], lo[1] + self.patch_shape[1], lo[2] + size)
        if min(lo) >= 0 and all(h <= n for h, n in zip(hi, t.shape)):
            return t[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]        # inside the volume: a view, no gather
        ix = np.clip(np.arange(corner[0], corner[0] + self.patch_shape[0]), 0, t.shape[0] - 1)
        iy = np.clip(np.arange(corner[1], corner[1] + self.patch_shape[1]), 0, t.shape[1] - 1)
        iz = np.clip(np.arange(corner[2] + index, corner[2] + index + size), 0, t.shape[2] - 1)
        return t[np.ix_(ix, iy, iz)]

    def draw(self):
        """One batch worth of accepted samples: (cases int32 [B], corners int32 [B, 3])."""
        cases, corners = [], []
        while len(cases) < self.batch_size:
            index = next(self._index_generator)
            shape = self._shapes[index]
            corner = [np.random.randint(low=0, high=high) for high in np.subtract(shape, self.patch_shape)]
            if self.drop_easy_patches or self.skip_blank:
                truth = self._truth_patch(index, corner, self.truth_index, self.truth_size)
                if self.drop_easy_patches:
                    truth_mean = np.mean(truth[16:-16, 16:-16, :])
                    if 1 - np.abs(truth_mean - 0.5) < np.random.random():
                        continue
                if self.skip_blank and not np.any(truth != 0):
                    continue
            cases.append(index)
            corners.append(corner)
        return np.asarray(cases, np.int32), np.asarray(corners, np.int32).reshape(-1, 3)

    def _args(self):
        return (self.truth_index, self.truth_size, self.prev_truth_index if self.prev_truth_size else 0,
                self.prev_truth_size)


class DeviceSampler(PatchDraws):
    """data_generator(data_file, index_list, ...) with the cases resident in HBM.

    data_list / truth_list: per-case arrays [X, Y, Z] (what `data_file.root.data[i]` / `.truth[i]` return after
    DataFileDummy / pad_samples). Arguments carry the reference's names and defaults (generator.py:222-226).
    """

    def __init__(self, data_list, truth_list, index_list, device_augment=None, device=None, **kw):
        if len(data_list) != len(truth_list):
            raise ValueError("data_list and truth_list differ in length")
        PatchDraws.__init__(self, truth_list, index_list, **kw)
        self.device_augment = device_augment
        self._lib = _lib.load()
        self._ctx = _lib.get_context(device)
        h = _lib.c_vp()
        _lib.check(self._lib.fm_volset_create(self._ctx.handle, len(data_list), ctypes.byref(h)))
        self._handle = h
        for i, (d, t) in enumerate(zip(data_list, truth_list)):
            d32, t32 = _lib.f32c(d), _lib.f32c(t)
            if d32.shape != t32.shape or d32.ndim != 3:
                raise ValueError("case %d: data %s / truth %s must be equal-shaped 3-D volumes" % (i, d32.shape, t32.shape))
            _lib.check(self._lib.fm_volset_set_case(h, i, _lib.fptr(d32), _lib.fptr(t32), _lib.i32ptr(_lib.i32x(d32.shape))))

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                self._lib.fm_volset_destroy(h)
            except Exception:
                pass

    def _aug_array(self, n):
        """fm_sample_aug[n] from `device_augment` = dict(flip=bool, intensity=(lo, hi), noise_sigma=float) or None."""
        a = self.device_augment
        if not a:
            return None
        arr = (_lib.SampleAug * n)()
        for b in range(n):
            arr[b].flip = int(np.random.randint(0, 8)) if a.get("flip") else 0
            lo, hi = a.get("intensity", (1.0, 1.0))
            arr[b].intensity_scale = float(np.random.uniform(lo, hi))
            arr[b].noise_sigma = float(a.get("noise_sigma", 0.0))
            arr[b].noise_seed = int(np.random.randint(0, 2 ** 31 - 1))
        return arr

    def gather(self, cases, corners):
        """The batch of the given (case, corner) samples, cut on the device and copied back: x, y shaped like
        convert_data's output (generator.py:380-401)."""
        cases, corners = _lib.i32x(cases), _lib.i32x(corners)
        n = cases.size
        p0, p1, p2 = self.patch_shape
        x = np.empty((n, p0, p1, p2 + self.prev_truth_size), np.float32)
        y = np.empty((n, p0, p1, self.truth_size), np.float32)
        ti, ts, pi, ps = self._args()
        _lib.check(self._lib.fm_volset_gather(self._handle, _lib.i32ptr(cases), _lib.i32ptr(corners), self._aug_array(n),
                                              n, _lib.i32ptr(_lib.i32x(self.patch_shape)), ti, ts, pi, ps,
                                              _lib.fptr(x), _lib.fptr(y)))
        if self.is3d:
            x, y = np.expand_dims(x, 1), np.expand_dims(y, 1)
        return x, y

    def __iter__(self):
        return self

    def __next__(self):
        cases, corners = self.draw()
        return self.gather(cases, corners)

    def train_on_next_batch(self, model):
        """next(generator) + model.train_on_batch(x, y) without the batch ever visiting the host. Returns the
        metrics of Model.train_on_batch ([loss, dice_coefficient, vod_coefficient, binary_accuracy])."""
        cases, corners = self.draw()
        return model.train_on_sampled_batch(self, cases, corners)
