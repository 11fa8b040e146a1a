"""patch_wise_prediction on the GPU: index mapping / overlap-add must be BIT-EXACT against the goldens
frozen from the unmodified reference (tests/golden/make_golden.py) and against the oracle."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import prediction_oracle as po
from oracle import unet_oracle as uo
from oracle.ref_harness import FunctionModel
from tests.golden.make_golden import ramp_model

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
