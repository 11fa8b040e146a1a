"""Generates tests/golden/*.npz by running the UNMODIFIED reference code
(/root/reference/fetal_net/prediction.py, utils/patches.py) under the stub harness of
oracle/ref_harness.py. Run in the build container only (the reference tree does not travel):

    python tests/golden/make_golden.py

Everything written here is deterministic IEEE arithmetic (NumPy float32 mul/add, float64 sums), so
the fixtures are machine-independent. The fake "models" are position-dependent so that overlapping
patches contribute DIFFERENT values to a voxel (an identity model cannot catch index/order bugs).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_harness import FunctionModel, load_reference_prediction  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ramp_model(patch_shape, channels=1):
    """predict(batch[B,1,x,y,z]) -> float32 [B,C,x,y,z]: batch * ramp_c + 0.125 * ramp_c (exact fp32 ops)."""
    px, py, pz = patch_shape
    g = np.meshgrid(np.arange(px), np.arange(py), np.arange(pz), indexing="ij")
    ramps = []
    for c in range(channels):
        r = (1.0 + (g[0] * 3 + g[1] * 5 + g[2] * 7 + c * 11) % 13).astype(np.float32) / np.float32(16.0)
        ramps.append(r)
    ramp = np.stack(ramps)[None]                                  # [1,C,x,y,z]

    def fn(batch):
        b = np.asarray(batch).astype(np.float32)                  # Keras casts the feed to float32
        return (b * ramp + np.float32(0.125) * ramp).astype(np.float32)

    return fn, (None, channels) + tuple(patch_shape)


def ramp_model_2d(hw, n_in):
    """2D / 2.5D fake model: predict(batch[B,H,W,n_in]) -> float32 [B,H,W,1]; a fixed-order float32 combination
    of all input channels (slices and previous-truth slices), position dependent."""
    h, w = hw
    g = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    ramp = ((1.0 + (g[0] * 3 + g[1] * 5) % 13).astype(np.float32) / np.float32(16.0))[None, :, :, None]

    def fn(batch):
        b = np.asarray(batch).astype(np.float32)
        acc = np.zeros(b.shape[:3] + (1,), np.float32)
        for c in range(n_in):                                       # fixed summation order
            acc = (acc + b[..., c:c + 1] * np.float32(0.25 * (c + 1))).astype(np.float32)
        return (acc * ramp + np.float32(0.125) * ramp).astype(np.float32)

    return fn, (None, h, w, 1)


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    pred = load_reference_prediction()
    patches = pred._ref_patches
    fx = {}

    # ---- (1) plans: get_set_of_patch_indices_full with the overlap arithmetic of prediction.py:135-137
    plan_cases = [
        ("cfg1_f05", (256, 256, 64), (64, 64, 64), (64, 64, 64), 0.5),
        ("cfg1_f0", (256, 256, 64), (64, 64, 64), (64, 64, 64), 0.0),
        ("cfg1_f09", (256, 256, 64), (64, 64, 64), (64, 64, 64), 0.9),
        ("cfg3_f05", (256, 256, 64), (128, 128, 64), (128, 128, 64), 0.5),
        ("cfg5_f05", (512, 512, 128), (128, 128, 64), (128, 128, 64), 0.5),
        ("cfg4_2d_f05", (256, 256, 68), (256, 256, 5), (256, 256, 1), 0.5),
        ("odd_f03", (70, 50, 41), (32, 16, 8), (32, 16, 8), 0.3),
        ("odd_f077", (97, 33, 40), (32, 32, 16), (32, 32, 16), 0.77),
    ]
    for name, padded, patch, pshape, f in plan_cases:
        min_overlap = np.subtract(patch, pshape)
        max_overlap = np.subtract(patch, (1, 1, 1))
        overlap = min_overlap + (f * (max_overlap - min_overlap)).astype(int)
        idx = pred.get_set_of_patch_indices_full((0, 0, 0), np.subtract(padded, patch), np.subtract(patch, overlap))
        fx["plan/%s/args" % name] = np.array(list(padded) + list(patch) + list(pshape), np.int64)
        fx["plan/%s/f" % name] = np.float64(f)
        fx["plan/%s/n" % name] = np.int64(len(idx))
        if len(idx) <= 4096:
            fx["plan/%s/idx" % name] = idx.astype(np.int32)
        fx["plan/%s/sha" % name] = np.array(sha16(idx.astype(np.int32)))

    # ---- (2) cfg-1 count map (SURVEY.md §8c golden 5): run the reference accumulate loop on ones
    ones_model = FunctionModel(lambda b: np.ones((len(b), 1, 64, 64, 64), np.float32), (None, 1, 64, 64, 64))
    # count = sum of ones predictions before the divide: recover it by calling with a constant volume and
    # instrumenting through a second pass: patch_wise_prediction returns sum/count == 1, so compute the
    # count with the reference's own index list instead.
    idx = pred.get_set_of_patch_indices_full((0, 0, 0), (192, 192, 0), (33, 33, 33))
    cnt = np.zeros((256, 256, 64), np.int16)
    for x, y, z in idx:
        cnt[x:x + 64, y:y + 64, z:z + 64] += 1          # prediction.py:193
    fx["count/cfg1/sha"] = np.array(sha16(cnt))
    vals, freq = np.unique(cnt, return_counts=True)
    fx["count/cfg1/hist"] = np.stack([vals.astype(np.int64), freq.astype(np.int64)])
    fx["count/cfg1/axis_x"] = cnt[:, 0, 0].copy()
    out = pred.patch_wise_prediction(ones_model, np.zeros((1, 256, 256, 64), np.float32), patch_shape=(64, 64, 64),
                                     overlap_factor=0.5)
    assert out.shape == (256, 256, 64, 1) and np.all(out == 1.0)

    # ---- (3) full patch_wise_prediction runs with position-dependent fake models
    run_cases = [
        # name, volume shape, patch, overlap_factor, batch_size, channels
        ("ramp_a", (1, 40, 36, 20), (16, 16, 16), 0.5, 5, 1),
        ("ramp_fit", (1, 30, 12, 20), (16, 16, 16), 0.5, 3, 1),      # pad_for_fit on y (12 < 16)
        ("ramp_f08", (1, 24, 24, 17), (8, 8, 8), 0.8, 7, 1),
        ("ramp_c2", (1, 20, 24, 16), (8, 16, 8), 0.3, 4, 2),         # two output channels
        ("ramp_f0", (1, 32, 32, 16), (16, 16, 16), 0.0, 5, 1),
    ]
    for name, vshape, patch, f, bs, ch in run_cases:
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        vol = rng.standard_normal(vshape).astype(np.float32)
        fn, oshape = ramp_model(patch, ch)
        out = pred.patch_wise_prediction(FunctionModel(fn, oshape), vol, patch_shape=patch, overlap_factor=f,
                                         batch_size=bs)
        fx["run/%s/vol" % name] = vol
        fx["run/%s/patch" % name] = np.array(patch, np.int64)
        fx["run/%s/f" % name] = np.float64(f)
        fx["run/%s/batch" % name] = np.int64(bs)
        fx["run/%s/channels" % name] = np.int64(ch)
        fx["run/%s/out" % name] = out                                 # float64 [X,Y,Z,C]

    # ---- (3b) 2D / 2.5D runs (prediction.py:131-141: prediction_shape (H,W,1), z halo (ceil,floor)((D-1)/2)),
    #           with and without previous-slice truth conditioning (prediction.py:106-110,148-156)
    run2d = [
        # name, volume shape, patch (H,W,D), overlap_factor, batch, prev_truth_index, prev_truth_size
        ("s2d_plain", (1, 24, 20, 9), (16, 16, 5), 0.5, 5, None, None),
        ("s2d_truth1", (1, 24, 20, 9), (16, 16, 5), 0.5, 4, 1, 1),
        ("s2d_truth2", (1, 16, 16, 7), (16, 16, 3), 0.0, 3, 0, 2),
        ("s2d_fit", (1, 12, 20, 6), (16, 16, 5), 0.7, 5, 1, 1),          # x needs pad_for_fit
    ]
    for name, vshape, patch, f, bs, pti, pts in run2d:
        rng = np.random.default_rng(sum(map(ord, name)))
        vol = rng.standard_normal(vshape).astype(np.float32)
        truth = (rng.random(vshape) < 0.4).astype(np.float32) if pts else None
        fn, oshape = ramp_model_2d(patch[:2], patch[2] + (pts or 0))
        out = pred.patch_wise_prediction(FunctionModel(fn, oshape), vol, patch_shape=patch, overlap_factor=f,
                                         batch_size=bs, truth_data=truth, prev_truth_index=pti, prev_truth_size=pts)
        fx["run2d/%s/vol" % name] = vol
        if truth is not None:
            fx["run2d/%s/truth" % name] = truth
        fx["run2d/%s/patch" % name] = np.array(patch, np.int64)
        fx["run2d/%s/f" % name] = np.float64(f)
        fx["run2d/%s/batch" % name] = np.int64(bs)
        fx["run2d/%s/prev" % name] = np.array([-1 if pti is None else pti, 0 if pts is None else pts], np.int64)
        fx["run2d/%s/out" % name] = out

    # ---- (4) get_patch_from_3d_data incl. the out-of-bounds edge-pad branch (patches.py:57-91)
    rng = np.random.default_rng(7)
    data = rng.standard_normal((2, 12, 10, 9)).astype(np.float32)
    corners = [(0, 0, 0), (4, 2, 1), (8, 6, 5), (-2, 0, 0), (9, 7, 6), (-1, -3, 6)]
    fx["patch/data"] = data
    fx["patch/corners"] = np.array(corners, np.int64)
    fx["patch/shape"] = np.array((4, 4, 4), np.int64)
    for i, c in enumerate(corners):
        fx["patch/out%d" % i] = np.ascontiguousarray(patches.get_patch_from_3d_data(data, (4, 4, 4), np.array(c)))

    # ---- (5) test-time augmentation wrappers (prediction.py:65-85, 364-369, 25-62)
    rng = np.random.default_rng(11)
    vol = rng.standard_normal((1, 24, 20, 16)).astype(np.float32)
    fn, oshape = ramp_model((8, 8, 8), 1)
    flips = pred.predict_flips(vol, FunctionModel(fn, oshape), 0.5, {"patch_shape": [8, 8], "patch_depth": 8})
    fx["tta/flips/vol"] = vol
    fx["tta/flips/out"] = np.stack(flips)                              # [8,X,Y,Z] float64, powerset order

    pdata = rng.standard_normal((1, 8, 8, 6)).astype(np.float32)
    fn, oshape = ramp_model((8, 8, 6), 2)
    fx["tta/perm/data"] = pdata
    fx["tta/perm/out"] = np.asarray(pred.predict_with_permutations(FunctionModel(fn, oshape), pdata))   # [C,x,y,z]
    vol = rng.standard_normal((1, 20, 12, 9)).astype(np.float32)
    fx["tta/perm_pw/vol"] = vol
    fx["tta/perm_pw/out"] = pred.patch_wise_prediction(FunctionModel(fn, oshape), vol, patch_shape=(8, 8, 6),
                                                       overlap_factor=0.5, batch_size=4, permute=True)

    # predict_augment: skimage is absent, so the reference's contrast_augment (augment.py:123-126) gets the
    # restated rescale_intensity injected — the RNG draw order, flips, transposes and scipy rotations are the
    # reference's own code.
    def rescale(d, lo, hi):
        d = np.asarray(d)
        omin, omax = d.min(), d.max()
        return (((np.clip(d, lo, hi) - lo) / (hi - lo)) * (omax - omin) + omin).astype(d.dtype)
    pred.contrast_augment = rescale
    vol = rng.standard_normal((1, 16, 16, 8)).astype(np.float32)
    fn, oshape = ramp_model((8, 8, 8), 1)
    np.random.seed(1234)
    fx["tta/augment/vol"] = vol
    fx["tta/augment/out"] = pred.predict_augment(vol, FunctionModel(fn, oshape), 0.5, (8, 8, 8), num_augments=1)

    np.savez_compressed(os.path.join(OUT, "prediction_golden.npz"), **fx)
    print("wrote", os.path.join(OUT, "prediction_golden.npz"), len(fx), "arrays")




def make_sampler_golden():
    """tests/golden/sampler_golden.npz: batches of the UNMODIFIED reference training generator
    (fetal_net/generator.py:222-348, augment=None) on small synthetic cases, with the (case, corner) draws recovered
    from the same seeds. shuffle_index_list=False: random_list_generator re-seeds np.random from the OS every pass
    (generator.py:202), which no golden can follow."""
    from oracle.ref_harness import load_reference_generator
    gen = load_reference_generator()

    class Root:
        pass

    rng = np.random.default_rng(1234)
    shapes = [(40, 36, 20), (37, 41, 23), (48, 35, 19)]
    data = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    truth = []
    for s in shapes:                                   # blobs with empty borders so that skip_blank rejects some
        t = np.zeros(s, np.float32)
        c = [v // 2 for v in s]
        t[c[0] - 8:c[0] + 9, c[1] - 7:c[1] + 8, c[2] - 4:c[2] + 5] = 1.0
        truth.append(t)
    df = Root()
    df.root = df
    df.data, df.truth, df.mask = data, truth, None
    fx = {"n_cases": np.int32(len(shapes))}
    for i in range(len(shapes)):
        fx["data_%d" % i], fx["truth_%d" % i] = data[i], truth[i]
    configs = {
        # name: (patch_shape, kwargs of data_generator)
        "u3d": ((16, 16, 8), dict(truth_index=0, truth_size=8, is3d=True, skip_blank=True)),
        "u25d": ((24, 24, 5), dict(truth_index=2, truth_size=1, prev_truth_index=1, prev_truth_size=1, skip_blank=True)),
        "u25d_edge": ((24, 24, 5), dict(truth_index=5, truth_size=1, prev_truth_index=-1, prev_truth_size=2,
                                         skip_blank=False)),
        "u2d_easy": ((34, 34, 3), dict(truth_index=1, truth_size=1, skip_blank=False, drop_easy_patches=True)),
    }
    for name, (patch, kw) in configs.items():
        np.random.seed(77)
        g = gen.data_generator(df, [2, 0, 1], batch_size=3, patch_shape=patch, shuffle_index_list=False, augment=None,
                               categorical=False, **kw)
        for b in range(3):
            x, y = next(g)
            fx["%s_x%d" % (name, b)] = np.asarray(x, np.float32)
            fx["%s_y%d" % (name, b)] = np.asarray(y, np.float32)
    np.savez_compressed(os.path.join(OUT, "sampler_golden.npz"), **fx)
    print("sampler golden:", {k: v.shape for k, v in fx.items() if k.startswith("u") and k.endswith("0")})


if __name__ == "__main__":
    if "--sampler" in sys.argv:          # python tests/golden/make_golden.py --sampler : only the sampler fixtures
        make_sampler_golden()
    else:
        main()
        make_sampler_golden()
