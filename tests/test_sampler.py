"""Training sampler: the host draws and the device gather against batches frozen from the
UNMODIFIED reference generator (tests/golden/make_golden.py::make_sampler_golden, fetal_net/generator.py:222-348)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fetal-mri-segmentation_b200"))

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "sampler_golden.npz"))
CONFIGS = {   # the data_generator arguments of make_sampler_golden
    "u3d": ((16, 16, 8), dict(truth_index=0, truth_size=8, is3d=True, skip_blank=True)),
    "u25d": ((24, 24, 5), dict(truth_index=2, truth_size=1, prev_truth_index=1, prev_truth_size=1, skip_blank=True)),
    "u25d_edge": ((24, 24, 5), dict(truth_index=5, truth_size=1, prev_truth_index=-1, prev_truth_size=2, skip_blank=False)),
    "u2d_easy": ((34, 34, 3), dict(truth_index=1, truth_size=1, skip_blank=False, drop_easy_patches=True)),
}


def cases():
    n = int(GOLD["n_cases"])
    return [GOLD["data_%d" % i] for i in range(n)], [GOLD["truth_%d" % i] for i in range(n)]


def host_cut(data, truth, case, corner, patch, truth_index, truth_size, prev_truth_index=None, prev_truth_size=None,
             is3d=False, **_):
    """Independent NumPy statement of extract_patch + the concatenation of generator.py:305-306 (np.pad mode='edge'
    on the out-of-range side, then a plain slice)."""
    def cut(vol, z0, nz):
        lo = [corner[0], corner[1], z0]
        sz = [patch[0], patch[1], nz]
        before = [max(0, -l) for l in lo]
        after = [max(0, l + s - d) for l, s, d in zip(lo, sz, vol.shape)]
        v = np.pad(vol, list(zip(before, after)), mode="edge")
        lo = [l + b for l, b in zip(lo, before)]
        return v[lo[0]:lo[0] + sz[0], lo[1]:lo[1] + sz[1], lo[2]:lo[2] + sz[2]]
    x = cut(data[case], corner[2], patch[2])
    y = cut(truth[case], corner[2] + truth_index, truth_size)
    if prev_truth_index is not None:
        x = np.concatenate([x, cut(truth[case], corner[2] + prev_truth_index, prev_truth_size)], axis=-1)
    if is3d:
        x, y = x[None], y[None]
    return x, y


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_host_draws_reproduce_the_reference_generator(name):
    """Same seed => PatchDraws makes the reference's np.random calls in the reference's order (corner draws, the
    drop_easy_patches draw, skip_blank rejections): the samples it names, cut on the host, ARE the golden batches."""
    from fetal_net.device_sampler import PatchDraws
    data, truth = cases()
    patch, kw = CONFIGS[name]
    np.random.seed(77)
    d = PatchDraws(truth, [2, 0, 1], batch_size=3, patch_shape=patch, shuffle_index_list=False, **kw)
    for b in range(3):
        cs, corners = d.draw()
        assert cs.dtype == np.int32 and corners.shape == (3, 3)
        xs, ys = zip(*[host_cut(data, truth, int(c), [int(v) for v in k], patch, **kw) for c, k in zip(cs, corners)])
        assert np.array_equal(np.stack(xs), GOLD["%s_x%d" % (name, b)]), (name, b)
        assert np.array_equal(np.stack(ys), GOLD["%s_y%d" % (name, b)]), (name, b)


def test_patch_draws_rejects_what_the_device_path_does_not_cover():
    from fetal_net.device_sampler import PatchDraws
    _, truth = cases()
    for bad in (dict(augment={"flip": True}), dict(categorical=True), dict(truth_downsample=2)):
        with pytest.raises(NotImplementedError):
            PatchDraws(truth, [0], patch_shape=(8, 8, 8), **bad)
    with pytest.raises(ValueError):
        PatchDraws(truth, [0])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_device_sampler_bit_exact_against_reference_generator(name):
    from fetal_net.device_sampler import DeviceSampler
    data, truth = cases()
    patch, kw = CONFIGS[name]
    np.random.seed(77)
    s = DeviceSampler(data, truth, [2, 0, 1], batch_size=3, patch_shape=patch, shuffle_index_list=False, **kw)
    for b in range(3):
        x, y = next(s)
        assert x.dtype == np.float32 and np.array_equal(x, GOLD["%s_x%d" % (name, b)]), (name, b)
        assert np.array_equal(y, GOLD["%s_y%d" % (name, b)]), (name, b)


@pytest.mark.gpu
def test_device_augmentations_flip_scale():
    """Flips act on data, previous truth and target alike; the intensity scale on the data channels only."""
    from fetal_net import _lib
    from fetal_net.device_sampler import DeviceSampler
    data, truth = cases()
    patch, kw = CONFIGS["u25d"]
    s = DeviceSampler(data, truth, [0], batch_size=1, patch_shape=patch, shuffle_index_list=False, **kw)
    cs, corners = np.array([1], np.int32), np.array([[5, 7, 9]], np.int32)
    x0, y0 = s.gather(cs, corners)
    for flip in range(8):
        aug = (_lib.SampleAug * 1)()
        aug[0].flip, aug[0].intensity_scale, aug[0].noise_sigma, aug[0].noise_seed = flip, 1.5, 0.0, 0
        s._aug_array = lambda n, a=aug: a
        x, y = s.gather(cs, corners)
        ex, ey = x0.copy(), y0.copy()
        ex[..., :patch[2]] = ex[..., :patch[2]] * np.float32(1.5)
        if flip & 1:
            ex, ey = ex[:, ::-1], ey[:, ::-1]
        if flip & 2:
            ex, ey = ex[:, :, ::-1], ey[:, :, ::-1]
        if flip & 4:
            ex = np.concatenate([ex[..., :patch[2]][..., ::-1], ex[..., patch[2]:][..., ::-1]], -1)
            ey = ey[..., ::-1]
        assert np.array_equal(x, ex) and np.array_equal(y, ey), flip


@pytest.mark.gpu
def test_train_on_sampled_batch_equals_train_on_batch():
    """fm_train_step_sampled == next(generator) + train_on_batch: same losses, step after step (2.5D model with a
    previous-truth channel, deterministic kernels not needed: the bound is the run-to-run spread)."""
    from fetal_net.device_sampler import DeviceSampler
    from fetal_net.model import unet_model_2d
    rng = np.random.default_rng(5)
    data = [rng.standard_normal((48, 48, 12)).astype(np.float32) for _ in range(2)]
    truth = [(rng.random((48, 48, 12)) > 0.6).astype(np.float32) for _ in range(2)]
    kw = dict(batch_size=4, patch_shape=(32, 32, 5), shuffle_index_list=False, truth_index=2, truth_size=1,
              prev_truth_index=1, prev_truth_size=1, skip_blank=False)
    losses = []
    for on_device in (False, True):
        np.random.seed(3)
        s = DeviceSampler(data, truth, [0, 1], **kw)
        model = unet_model_2d(input_shape=(32, 32, 6), n_base_filters=16, depth=3, initial_learning_rate=1e-3)
        model.init_glorot_uniform(seed=11)
        ls = []
        for _ in range(4):
            if on_device:
                ls.append(s.train_on_next_batch(model)[0])
            else:
                x, y = next(s)
                ls.append(model.train_on_batch(x, y)[0])
        losses.append(ls)
    assert np.allclose(losses[0], losses[1], atol=2e-3), losses
    assert losses[0][-1] < losses[0][0]
