"""patch_wise_prediction on the GPU: index mapping / overlap-add must be BIT-EXACT against the goldens
frozen from the unmodified reference (tests/golden/make_golden.py) and against the oracle."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import prediction_oracle as po
from oracle import unet_oracle as uo
from oracle.ref_harness import FunctionModel
from tests.golden.make_golden import ramp_model, ramp_model_2d

pytestmark = pytest.mark.gpu


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def run_names(golden):
    return sorted({k.split("/")[1] for k in golden if k.startswith("run/")})


def test_golden_runs_bit_exact(golden, ctx):
    """Generic-model path: gather (GPU) -> model.predict (host fake model) -> overlap-add/average (GPU)."""
    from fetal_net.prediction import patch_wise_prediction
    for name in run_names(golden):
        vol = golden["run/%s/vol" % name]
        patch = tuple(int(v) for v in golden["run/%s/patch" % name])
        fn, oshape = ramp_model(patch, int(golden["run/%s/channels" % name]))
        out = patch_wise_prediction(FunctionModel(fn, oshape), vol, patch_shape=patch,
                                    overlap_factor=float(golden["run/%s/f" % name]),
                                    batch_size=int(golden["run/%s/batch" % name]))
        ref = golden["run/%s/out" % name]
        assert out.dtype == np.float64 and out.shape == ref.shape, name
        assert np.array_equal(out, ref), (name, float(np.abs(out - ref).max()))


def test_golden_2d_and_truth_runs_bit_exact(golden, ctx):
    """2D / 2.5D path of patch_wise_prediction incl. previous-slice truth conditioning (prediction.py:106-110)."""
    from fetal_net.prediction import patch_wise_prediction
    names = sorted({k.split("/")[1] for k in golden if k.startswith("run2d/")})
    for name in names:
        vol = golden["run2d/%s/vol" % name]
        patch = tuple(int(v) for v in golden["run2d/%s/patch" % name])
        pti, pts = [int(v) for v in golden["run2d/%s/prev" % name]]
        truth = golden.get("run2d/%s/truth" % name)
        fn, oshape = ramp_model_2d(patch[:2], patch[2] + pts)
        out = patch_wise_prediction(FunctionModel(fn, oshape), vol, patch_shape=patch,
                                    overlap_factor=float(golden["run2d/%s/f" % name]),
                                    batch_size=int(golden["run2d/%s/batch" % name]), truth_data=truth,
                                    prev_truth_index=pti if truth is not None else None,
                                    prev_truth_size=pts if truth is not None else None)
        ref = golden["run2d/%s/out" % name]
        assert out.shape == ref.shape and np.array_equal(out, ref), (name, float(np.abs(out - ref).max()))


def test_cfg1_count_map_bit_exact(golden, ctx):
    """49 patches of 64^3 on 256x256x64, overlap_factor 0.5: counts and 1/count weights (SURVEY §8c golden 5)."""
    from fetal_net import _lib
    from fetal_net.prediction import patch_plan
    lib = _lib.load()
    idx = patch_plan((256, 256, 64), (64, 64, 64), (64, 64, 64), 0.5)
    assert len(idx) == 49
    preds = np.ones((49, 64, 64, 64, 1), np.float32)
    out = np.empty((256, 256, 64, 1), np.float64)
    cnt = np.empty((256, 256, 64), np.int16)
    _lib.check(lib.fm_reassemble(ctx.handle, _lib.fptr(preds), _lib.i32ptr(idx), 49, _lib.i32ptr(_lib.i32x((64, 64, 64))),
                                 1, _lib.i32ptr(_lib.i32x((256, 256, 64))), _lib.dptr(out), _lib.i16ptr(cnt)))
    assert sha16(cnt) == "ba96762db2430b62" == str(golden["count/cfg1/sha"])
    assert np.all(out == 1.0)


def test_gather_matches_oracle_padding(ctx):
    from fetal_net import _lib
    from fetal_net.prediction import _geometry, patch_plan
    lib = _lib.load()

    class M:
        output_shape = (None, 1, 16, 16, 16)
    rng = np.random.default_rng(5)
    vol = rng.standard_normal((1, 30, 12, 20)).astype(np.float32)     # y needs pad_for_fit
    g = _geometry(M, vol, (16, 16, 16), 0.5)
    idx = patch_plan(g["padded"], (16, 16, 16), (16, 16, 16), 0.5)
    out = np.empty((len(idx), 16, 16, 16), np.float32)
    _lib.check(lib.fm_gather_patches(ctx.handle, _lib.fptr(vol[0].copy()), _lib.i32ptr(_lib.i32x(vol.shape[1:])),
                                     _lib.i32ptr(_lib.i32x(g["halo"])), _lib.i32ptr(_lib.i32x(g["fit"])),
                                     _lib.dptr(np.asarray(g["pad"], np.float64)), _lib.i32ptr(idx), len(idx),
                                     _lib.i32ptr(_lib.i32x((16, 16, 16))), _lib.fptr(out)))
    d0, _, _ = po.pad_volume(vol[0], (16, 16, 16), (16, 16, 16))
    for i, c in enumerate(idx):
        ref = po.extract_patch(d0, (16, 16, 16), c).astype(np.float32)     # Keras casts the feed to float32
        assert np.array_equal(out[i], ref), i


@pytest.fixture(scope="module")
def native():
    from fetal_net.model import unet_model_3d
    from tests.test_gpu_model import decisive_weights
    w = decisive_weights(uo.unet3d_layers(4, 16), seed=3)
    m = unet_model_3d(input_shape=(1, 32, 32, 32), n_base_filters=16)
    m.set_named_weights(w)
    return m, w


def test_native_pipeline_equals_own_predict_plus_oracle_reassembly(native):
    """fm_patchwise_predict (fused device pipeline) == oracle control flow driven by the same network's
    .predict: bit-exact, because patch coordinates, accumulation order and the fp64 divide are identical."""
    from fetal_net.prediction import patch_wise_prediction
    model, _ = native
    rng = np.random.default_rng(6)
    for vshape, f, bs in [((1, 72, 40, 32), 0.5, 5), ((1, 40, 20, 48), 0.3, 4)]:
        vol = rng.standard_normal(vshape).astype(np.float32)
        out = patch_wise_prediction(model, vol, patch_shape=(32, 32, 32), overlap_factor=f, batch_size=bs)
        ref = po.patch_wise_prediction(model, vol, (32, 32, 32), overlap_factor=f, batch_size=bs)
        assert out.shape == vshape[1:] + (1,) and out.dtype == np.float64
        assert np.array_equal(out, ref), float(np.abs(out - ref).max())
        # batch size does not change the result (reference property)
        out2 = patch_wise_prediction(model, vol, patch_shape=(32, 32, 32), overlap_factor=f, batch_size=64)
        assert np.array_equal(out, out2)


def test_native_pipeline_close_to_fp32_oracle(native):
    from fetal_net.prediction import patch_wise_prediction
    model, w = native
    rng = np.random.default_rng(7)
    vol = rng.standard_normal((1, 64, 48, 32)).astype(np.float32)
    out = patch_wise_prediction(model, vol, patch_shape=(32, 32, 32), overlap_factor=0.5)
    ref = po.patch_wise_prediction(uo.OracleModel(w, (1, 32, 32, 32)), vol, (32, 32, 32), overlap_factor=0.5)
    d = np.abs(out - ref)
    # bf16 tolerance of tests/test_gpu_model.py (random decisive weights: logit std ~2.7, 1.5-3 % logit error)
    assert d.max() <= 0.12 and d.mean() <= 0.006, (d.max(), d.mean())
    soft = (2 * (out * ref).sum() + 1) / ((out * out).sum() + (ref * ref).sum() + 1)
    assert soft >= 0.999, soft


def test_sharded_partial_sums_add_up(native):
    from fetal_net.prediction import patch_wise_prediction
    model, _ = native
    rng = np.random.default_rng(8)
    vol = rng.standard_normal((1, 72, 40, 32)).astype(np.float32)
    full = patch_wise_prediction(model, vol, patch_shape=(32, 32, 32), overlap_factor=0.5)
    parts = [patch_wise_prediction(model, vol, patch_shape=(32, 32, 32), overlap_factor=0.5, shard=(r, 3))
             for r in range(3)]
    s = parts[0][0] + parts[1][0] + parts[2][0]
    cnt = parts[0][1]
    assert all(np.array_equal(cnt, p[1]) for p in parts)
    assert np.array_equal(s / cnt[..., None], full)


# ---- native 2D / 2.5D model ----------------------------------------------------------------------

@pytest.fixture(scope="module")
def native2d():
    from fetal_net.model import unet_model_2d
    layers = uo.unet2d_layers(4, 32, 6)
    w = uo.glorot_uniform_weights(layers, seed=4, ndim=2)
    rng = np.random.default_rng(5)
    for k in w:
        w[k] = (w[k] * np.sqrt(2.0) * 1.15).astype(np.float32) if k.endswith("/kernel") else \
            (0.05 * rng.standard_normal(w[k].shape)).astype(np.float32)
    m = unet_model_2d(input_shape=(32, 32, 6), n_base_filters=32, depth=4)
    m.set_named_weights(w)
    return m, w


def test_unet2d_forward_matches_oracle(native2d):
    model, w = native2d
    assert model.count_params() == 5441569
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 32, 32, 6)).astype(np.float32)
    p = model.predict(x)
    assert p.shape == (3, 32, 32, 1)
    with torch.no_grad():
        ref = uo.unet2d_forward(torch.as_tensor(x), w).numpy()
    logit = lambda q: np.log(np.clip(q.astype(np.float64), 1e-7, 1 - 1e-7) / np.clip(1 - q.astype(np.float64), 1e-7, 1))
    rel = np.linalg.norm(logit(p) - logit(ref)) / np.linalg.norm(logit(ref))
    assert rel <= 0.03 and np.abs(p - ref).mean() <= 0.006, (rel, np.abs(p - ref).mean())
    assert np.array_equal(model.predict(x[1:2]), p[1:2])      # batch invariance / determinism


def test_unet2d_train_step_matches_oracle(native2d):
    from fetal_net.model import unet_model_2d
    _, w0 = native2d
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 32, 32, 6)).astype(np.float32)
    t = (rng.random((2, 32, 32, 1)) < 0.3).astype(np.float32)
    model = unet_model_2d(input_shape=(32, 32, 6), n_base_filters=32, depth=4, initial_learning_rate=1e-4)
    model.set_named_weights(w0)
    ref = uo.train_step(uo.unet2d_forward, x, t, {k: v.copy() for k, v in w0.items()}, {}, 1e-4)
    got = model.train_on_batch(x, t)
    assert got[0] == pytest.approx(ref["loss"], abs=3e-3), (got, ref["loss"])
    grads = model.get_gradients()
    bad = []
    for l, gk in zip(model.layers, grads[0::2]):
        r = ref["grads"][l["name"] + "/kernel"].astype(np.float64).ravel()
        g = gk.astype(np.float64).ravel()
        cos = float(g @ r / max(np.linalg.norm(g) * np.linalg.norm(r), 1e-300))
        if cos < (0.95 if l["name"] in ("enc0a", "enc0b") else 0.99):
            bad.append((l["name"], cos))
    assert not bad, bad


def test_unet2d_spatial_dropout_train_step_matches_oracle(native2d):
    """unet_model_2d(dropout_rate=0.25): SpatialDropout2D behind enc<d>a / dec<d>a in training steps
    (unet/unet.py:60-61,76-77). The keep masks come from the library's counter-based hash, restated in the oracle
    (library_dropout_scales), so loss and gradients of a step WITH dropout are compared; inference ignores it."""
    from fetal_net.model import unet_model_2d
    _, w0 = native2d
    depth, nf, rate, seed, B = 4, 32, 0.25, 1234, 2
    rng = np.random.default_rng(7)
    x = rng.standard_normal((B, 32, 32, 6)).astype(np.float32)
    t = (rng.random((B, 32, 32, 1)) < 0.3).astype(np.float32)
    model = unet_model_2d(input_shape=(32, 32, 6), n_base_filters=nf, depth=depth, initial_learning_rate=1e-4,
                          dropout_rate=rate, dropout_seed=seed)
    plain = unet_model_2d(input_shape=(32, 32, 6), n_base_filters=nf, depth=depth, initial_learning_rate=1e-4)
    model.set_named_weights(w0)
    plain.set_named_weights(w0)
    assert np.array_equal(model.predict(x), plain.predict(x))           # identity at inference
    drop = {}
    for d in range(depth):                                              # slot d: enc<d>a (nf * 2^d channels)
        drop["enc%d" % d] = uo.library_dropout_scales(B * (nf << d), rate, seed + d).reshape(B, -1)
    for d in range(depth - 1):                                          # slot depth + d: dec<d>a (2 nf * 2^d channels)
        drop["dec%d" % d] = uo.library_dropout_scales(B * (2 * nf << d), rate, seed + depth + d).reshape(B, -1)
    assert 0.1 < np.mean([np.mean(v == 0) for v in drop.values()]) < 0.4
    ref = uo.train_step(lambda xt, prm: uo.unet2d_forward(xt, prm, depth=depth, drop=drop), x, t,
                        {k: v.copy() for k, v in w0.items()}, {}, 1e-4)
    # yardstick = the storage format: the same oracle with bf16-rounded activations / weights against itself in fp32
    bf = uo.train_step(lambda xt, prm: uo.unet2d_forward(xt, prm, depth=depth, drop=drop, quant=uo.bf16_round), x, t,
                       {k: v.copy() for k, v in w0.items()}, {}, 1e-4)
    got = model.train_on_batch(x, t)
    assert got[0] == pytest.approx(ref["loss"], abs=3e-3), (got, ref["loss"])
    assert abs(got[0] - plain.train_on_batch(x, t)[0]) > 1e-4           # the masks did change the forward pass
    grads = model.get_gradients()
    bad = []
    for l, gk in zip(model.layers, grads[0::2]):
        name = l["name"] + "/kernel"
        r = ref["grads"][name].astype(np.float64).ravel()
        g = gk.astype(np.float64).ravel()
        b = bf["grads"][name].astype(np.float64).ravel()
        cos = float(g @ r / max(np.linalg.norm(g) * np.linalg.norm(r), 1e-300))
        floor = min(0.95 if l["name"] in ("enc0a", "enc0b") else 0.99,
                    float(b @ r / max(np.linalg.norm(b) * np.linalg.norm(r), 1e-300)) - 0.02)
        if cos < floor:
            bad.append((l["name"], cos, floor))
    assert not bad, bad


def test_native_2d_pipeline_with_truth_equals_own_predict_plus_oracle_reassembly(native2d):
    """config 4 shape class: 5 slices + 1 previous-truth slice as channels, z-step 1, z halo (2,2)."""
    from fetal_net.prediction import patch_wise_prediction
    model, _ = native2d
    rng = np.random.default_rng(6)
    vol = rng.standard_normal((1, 48, 32, 10)).astype(np.float32)
    truth = (rng.random(vol.shape) < 0.4).astype(np.float32)
    out = patch_wise_prediction(model, vol, patch_shape=(32, 32, 5), overlap_factor=0.5, batch_size=7,
                                truth_data=truth, prev_truth_index=1, prev_truth_size=1)
    ref = po.patch_wise_prediction(model, vol, (32, 32, 5), overlap_factor=0.5, batch_size=7, truth_data=truth,
                                   prev_truth_index=1, prev_truth_size=1)
    assert out.shape == (48, 32, 10, 1) and out.dtype == np.float64
    assert np.array_equal(out, ref), float(np.abs(out - ref).max())


def test_tta_wrappers_match_reference_goldens(golden, ctx):
    """predict_flips / patch_wise_prediction(permute=True) / predict_augment (prediction.py:25-85,364-369):
    host re-indexing around the device pipeline; goldens frozen from the reference's own functions."""
    from fetal_net import prediction as P
    fn, oshape = ramp_model((8, 8, 8), 1)
    flips = P.predict_flips(golden["tta/flips/vol"], FunctionModel(fn, oshape), 0.5,
                            {"patch_shape": [8, 8], "patch_depth": 8})
    assert len(flips) == 8
    assert np.array_equal(np.stack(flips), golden["tta/flips/out"])

    fn2, oshape2 = ramp_model((8, 8, 6), 2)
    out = P.patch_wise_prediction(FunctionModel(fn2, oshape2), golden["tta/perm_pw/vol"], patch_shape=(8, 8, 6),
                                  overlap_factor=0.5, batch_size=4, permute=True)
    ref = golden["tta/perm_pw/out"]
    assert out.shape == ref.shape and out.dtype == np.float64
    np.testing.assert_allclose(out, ref, rtol=2e-6, atol=2e-6)    # fp32 mean over 48 keys, then fp64 overlap-add

    np.random.seed(1234)                                           # same draws as make_golden.py
    aug = P.predict_augment(golden["tta/augment/vol"], FunctionModel(fn, oshape), 0.5, (8, 8, 8), num_augments=1)
    ref = golden["tta/augment/out"]
    assert aug.shape == ref.shape
    np.testing.assert_allclose(aug, ref, rtol=1e-9, atol=1e-9)


def test_predict_flips_native_model_consistency(ctx):
    """Native model through predict_flips: entry () equals the plain call bit-for-bit (inference kernels are
    reproducible), and every entry is the un-flipped prediction of the flipped volume."""
    from fetal_net import prediction as P
    from fetal_net.model import unet_model_3d
    model = unet_model_3d(input_shape=(1, 16, 16, 16), n_base_filters=16, depth=2)
    model.init_glorot_uniform(seed=3)
    vol = np.random.default_rng(5).standard_normal((1, 24, 16, 16)).astype(np.float32)
    cfg = {"patch_shape": [16, 16], "patch_depth": 16}
    # a 3D volume: with a [1,X,Y,Z] array the reference flips input axes (C,X,Y) but un-flips prediction axes
    # (X,Y,Z) (prediction.py:72,78) - frozen as is in the golden test above
    flips = P.predict_flips(vol[0], model, 0.5, cfg)
    plain = P.patch_wise_prediction(model, vol, (16, 16, 16), overlap_factor=0.5).squeeze()
    assert np.array_equal(flips[0], plain)
    fl = P.patch_wise_prediction(model, np.flip(vol, 2)[...], (16, 16, 16), overlap_factor=0.5).squeeze()
    assert np.array_equal(flips[2], np.flip(fl, 1))
