"""Network-level parity at the REAL BASELINE.json shapes (the per-op and small-shape tests live in test_gpu_ops.py /
test_gpu_model.py / test_gpu_prediction.py). Everything goes through the reference-facing Python API on the default
(benchmarked) kernels and is compared with the fp32 CPU oracle on the same seeded inputs and weights.

  configs[1]  soft-Dice training step, batch 8 x 1x64x64x64, depth 4, 16 base filters
  configs[0]  patch_wise_prediction of one 1x256x256x64 volume, patch 64^3, overlap_factor 0.5 (49 patches)
  configs[2]  Isensee-2017 (depth 5, 16 filters, 3 segmentation levels) forward on 1x128x128x64
  configs[3]  2.5D U-Net (32 filters) forward on 256x256x(5 slices + 1 previous-truth slice)
  configs[4]  per-GPU shard of the data-parallel step: 128x128x64 patches (one sample here; the collectives are
              covered by tests/test_gpu_distributed.py and bench.py's dp_parity record)

Stated tolerances are those of tests/test_gpu_model.py (bf16 storage, fp32 accumulation): logits relative L2 error
<= 3 % (Isensee 4 %), mean |dp| <= 0.006 (Isensee 0.008), soft Dice >= 0.999, loss within 3e-3, per-layer gradient
cosine >= 0.99 (>= 0.95 / 0.97 for the two earliest layers).
"""
import numpy as np
import pytest
import torch

from oracle import prediction_oracle as po
from oracle import unet_oracle as uo
from tests.test_gpu_model import blob_target, decisive_weights, isensee_weights

pytestmark = pytest.mark.gpu


def _logit(q):
    q = np.clip(q.astype(np.float64), 1e-7, 1 - 1e-7)
    return np.log(q / (1 - q))


def _forward_close(p, ref, rel_tol, mean_tol):
    rel = float(np.linalg.norm(_logit(p) - _logit(ref)) / np.linalg.norm(_logit(ref)))
    soft = float((2 * (p.astype(np.float64) * ref).sum() + 1) / ((p.astype(np.float64) ** 2).sum() + (ref.astype(np.float64) ** 2).sum() + 1))
    mean = float(np.abs(p - ref).mean())
    assert rel <= rel_tol and mean <= mean_tol and soft >= 0.999, dict(rel=rel, mean=mean, soft=soft)
    return rel, mean, soft


def test_cfg2_train_step_8x64cube_matches_oracle():
    """BASELINE configs[1]: one full-size training step on the default (non-deterministic-order) kernels."""
    from fetal_net.model import unet_model_3d
    w0 = decisive_weights(uo.unet3d_layers(4, 16))
    rng = np.random.default_rng(1)
    x = rng.standard_normal((8, 1, 64, 64, 64)).astype(np.float32)
    t = blob_target(x.shape, rng)
    model = unet_model_3d(input_shape=(1, 64, 64, 64), n_base_filters=16, depth=4, initial_learning_rate=1e-4)
    model.set_named_weights(w0)
    ref = uo.unet3d_train_step(x, t, {k: v.copy() for k, v in w0.items()}, {}, 1e-4)
    got = model.train_on_batch(x, t)
    assert got[0] == pytest.approx(ref["loss"], abs=3e-3), (got, ref["loss"])
    assert got[1] == pytest.approx(ref["binary_accuracy"], abs=5e-3)
    assert got[2] == pytest.approx(ref["vod_coefficient"], abs=5e-3)
    grads = model.get_gradients()
    floor = {"enc0a": 0.95, "enc0b": 0.97}
    bad = []
    for l, gk, gb in zip(model.layers, grads[0::2], grads[1::2]):
        for kind, g in (("kernel", gk), ("bias", gb)):
            r = ref["grads"]["%s/%s" % (l["name"], kind)].astype(np.float64).ravel()
            g = g.astype(np.float64).ravel()
            cos = float(g @ r / max(np.linalg.norm(g) * np.linalg.norm(r), 1e-300))
            ratio = float(np.linalg.norm(g) / max(np.linalg.norm(r), 1e-300))
            if not (cos >= floor.get(l["name"], 0.99) and 0.92 <= ratio <= 1.08):
                bad.append((l["name"], kind, round(cos, 4), round(ratio, 4)))
    assert not bad, bad
    # the Keras-Adam update landed: every kernel moved by at most lr, and the packs used by the next forward follow it
    for l, k in zip(model.layers, model.get_weights()[0::2]):
        assert 0 < np.max(np.abs(k - w0[l["name"] + "/kernel"])) <= 1.01e-4, l["name"]


def test_cfg1_patchwise_256x256x64_matches_oracle():
    """BASELINE configs[0]: 49 patches of 64^3, overlap_factor 0.5, against the oracle network driven through the
    reference's own sliding-window control flow (oracle/prediction_oracle.py, pinned bit-exact to the reference)."""
    from fetal_net.model import unet_model_3d
    from fetal_net.prediction import patch_wise_prediction
    w = decisive_weights(uo.unet3d_layers(4, 16))
    model = unet_model_3d(input_shape=(1, 64, 64, 64), n_base_filters=16, depth=4)
    model.set_named_weights(w)
    vol = np.random.default_rng(0).standard_normal((1, 256, 256, 64)).astype(np.float32)
    out = patch_wise_prediction(model, vol, patch_shape=(64, 64, 64), overlap_factor=0.5, batch_size=49)
    assert out.shape == (256, 256, 64, 1) and out.dtype == np.float64
    ref = po.patch_wise_prediction(uo.OracleModel(w, (1, 64, 64, 64)), vol, (64, 64, 64), overlap_factor=0.5, batch_size=7)
    assert ref.shape == out.shape
    _forward_close(out.astype(np.float32), ref.astype(np.float32), 0.03, 0.006)
    # index mapping / reassembly weights: the GPU pipeline equals the oracle's overlap-add of ITS OWN patch outputs
    # bit for bit (the reference batch size of 5 and ours of 49 must not matter either)
    chk = po.patch_wise_prediction(model, vol, (64, 64, 64), overlap_factor=0.5, batch_size=5)
    assert np.array_equal(out, chk), float(np.abs(out - chk).max())


def test_cfg3_isensee_forward_128x128x64_matches_oracle():
    """BASELINE configs[2]: Isensee-2017 residual U-Net, depth 5, 16 base filters, 3 summed segmentation levels."""
    from fetal_net.model import isensee2017_model_3d
    shape, depth, nseg = (1, 128, 128, 64), 5, 3
    layers = uo.isensee3d_layers(depth, 16, nseg)
    w = isensee_weights(layers, seed=7)
    model = isensee2017_model_3d(input_shape=shape, n_base_filters=16, depth=depth, n_segmentation_levels=nseg)
    model.set_named_weights(w)
    x = np.random.default_rng(0).standard_normal((1,) + shape).astype(np.float32)
    p = model.predict(x)
    with torch.no_grad():
        ref = uo.isensee3d_forward(torch.as_tensor(x), w, depth=depth, n_segmentation_levels=nseg).numpy()
    assert p.shape == ref.shape == (1, 1, 128, 128, 64)
    _forward_close(p, ref, 0.04, 0.008)


def test_cfg4_unet2d_forward_256x256x6_matches_oracle():
    """BASELINE configs[3]: 2.5D U-Net, 5 slices + 1 previous-truth channel, 32 base filters, 256 x 256."""
    from fetal_net.model import unet_model_2d
    layers = uo.unet2d_layers(4, 32, 6)
    w = uo.glorot_uniform_weights(layers, seed=3, ndim=2)
    rng = np.random.default_rng(4)
    for k in w:
        if k.endswith("/kernel"):
            w[k] = (w[k] * np.sqrt(2.0) * 1.2).astype(np.float32)
        else:
            w[k] = (0.05 * rng.standard_normal(w[k].shape)).astype(np.float32)
    model = unet_model_2d(input_shape=(256, 256, 6), n_base_filters=32, depth=4)
    model.set_named_weights(w)
    x = rng.standard_normal((2, 256, 256, 6)).astype(np.float32)
    x[..., 5] = (rng.random(x.shape[:-1]) < 0.3)                      # binary previous-truth slice
    p = model.predict(x)
    with torch.no_grad():
        ref = uo.unet2d_forward(torch.as_tensor(x), w, depth=4).numpy()
    assert p.shape == ref.shape == (2, 256, 256, 1)
    _forward_close(p, ref, 0.03, 0.006)


def test_cfg5_train_step_128x128x64_patch_matches_oracle():
    """BASELINE configs[4]'s patch shape (128x128x64) through one training step (single sample here)."""
    from fetal_net.model import unet_model_3d
    w0 = decisive_weights(uo.unet3d_layers(4, 16))
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1, 1, 128, 128, 64)).astype(np.float32)
    t = blob_target(x.shape, rng)
    model = unet_model_3d(input_shape=(1, 128, 128, 64), n_base_filters=16, depth=4, initial_learning_rate=1e-4)
    model.set_named_weights(w0)
    ref = uo.unet3d_train_step(x, t, {k: v.copy() for k, v in w0.items()}, {}, 1e-4)
    got = model.train_on_batch(x, t)
    assert got[0] == pytest.approx(ref["loss"], abs=3e-3), (got, ref["loss"])
    grads = model.get_gradients()
    floor = {"enc0a": 0.95, "enc0b": 0.97}
    bad = []
    for l, gk in zip(model.layers, grads[0::2]):
        r = ref["grads"]["%s/kernel" % l["name"]].astype(np.float64).ravel()
        g = gk.astype(np.float64).ravel()
        cos = float(g @ r / max(np.linalg.norm(g) * np.linalg.norm(r), 1e-300))
        if cos < floor.get(l["name"], 0.99):
            bad.append((l["name"], round(cos, 4)))
    assert not bad, bad
