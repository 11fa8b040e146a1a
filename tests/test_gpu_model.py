"""Network-level parity on the GPU: unet_model_3d forward / training step through the reference-facing
Python API (fetal_net.model.unet_model_3d -> Model.predict / train_on_batch) against the fp32 oracle.

Stated tolerances (every activation stored as bf16, fp32 accumulation, 15 layers deep; the same fp32
oracle with its activations/weights rounded to bf16 lands at 1.5-1.7 % relative logit error, so the
bound is the storage format's, not the kernels'):
  logits         relative L2 error ||z - z_oracle|| / ||z_oracle|| <= 3 %
  probabilities  mean |p - p_oracle| <= 0.006, soft-Dice(p, p_oracle) >= 0.999
  masks          Dice(p > 0.5, p_oracle > 0.5) >= 0.999 on a TRAINED network (segmentation-like output;
                 on random weights + noise input every voxel sits near the decision boundary and even the
                 bf16-rounded oracle only reaches 0.996) - the north_star bar
  loss           |loss - loss_oracle| <= 3e-3
  gradients      per-layer cosine similarity vs the fp32 oracle >= 0.99 (>= 0.95 for enc0a, >= 0.97 for
                 enc0b) and norm ratio within 8 %. The two earliest layers sit at the end of back-propagation
                 on a noise input with random weights, where ReLU-mask / max-pool routing decisions flip on
                 bf16-rounded activations: the fp32 oracle with only its FORWARD activations rounded to bf16
                 scores 0.968 (enc0a) / 0.984 (enc0b) against itself in fp32 - the bound is the storage format.
"""
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as uo

pytestmark = pytest.mark.gpu


def decisive_weights(layers, seed=0):
    """Glorot-uniform with gain sqrt(2) (variance preserving through ReLU) + small biases, so the logits are
    O(1) and masks are decisive; plain glorot at 15 layers gives p = 0.5 +- 0.01 (useless for Dice)."""
    w = uo.glorot_uniform_weights(layers, seed=seed)
    rng = np.random.default_rng(seed + 1)
    for k in w:
        if k.endswith("/kernel"):
            w[k] = (w[k] * np.sqrt(2.0) * 1.2).astype(np.float32)
        else:
            w[k] = (0.05 * rng.standard_normal(w[k].shape)).astype(np.float32)
    return w


def blob_target(shape, rng):
    B, _, X, Y, Z = shape
    g = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij")
    t = np.zeros(shape, np.float32)
    for b in range(B):
        c = rng.uniform(0.3, 0.7, 3) * [X, Y, Z]
        r = rng.uniform(0.2, 0.35) * min(X, Y, Z)
        t[b, 0] = ((g[0] - c[0]) ** 2 + (g[1] - c[1]) ** 2 + (g[2] - c[2]) ** 2) < r * r
    return t


def mask_dice(a, b):
    a, b = a > 0.5, b > 0.5
    return (2.0 * (a & b).sum() + 1.0) / (a.sum() + b.sum() + 1.0)


@pytest.fixture(scope="module")
def setup():
    from fetal_net.model import unet_model_3d
    layers = uo.unet3d_layers(4, 16)
    w = decisive_weights(layers)
    model = unet_model_3d(input_shape=(1, 32, 32, 32), n_base_filters=16, depth=4, initial_learning_rate=1e-4)
    model.set_named_weights(w)
    return model, w


def test_layer_table_and_param_count(setup):
    model, w = setup
    assert model.count_params() == 4079713
    assert [l["name"] for l in model.layers] == [n for n, *_ in uo.unet3d_layers(4, 16)]
    assert model.output_shape == (None, 1, 32, 32, 32)


def test_weights_roundtrip_exact(setup):
    model, w = setup
    got = model.get_weights()
    for l, k, b in zip(model.layers, got[0::2], got[1::2]):
        assert np.array_equal(k, w[l["name"] + "/kernel"]) and np.array_equal(b, w[l["name"] + "/bias"])


def test_forward_matches_oracle(setup):
    model, w = setup
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 1, 32, 32, 32)).astype(np.float32)
    p = model.predict(x)
    assert p.shape == (3, 1, 32, 32, 32) and p.dtype == np.float32
    with torch.no_grad():
        ref = uo.unet3d_forward(torch.as_tensor(x), w).numpy()
    d = np.abs(p - ref)
    frac_decisive = np.mean(np.abs(ref - 0.5) > 0.05)
    assert frac_decisive > 0.5, "test weights are not decisive (%.2f)" % frac_decisive
    logit = lambda q: np.log(np.clip(q.astype(np.float64), 1e-7, 1 - 1e-7) / np.clip(1 - q.astype(np.float64), 1e-7, 1))
    zr, zg = logit(ref), logit(p)
    rel = np.linalg.norm(zg - zr) / np.linalg.norm(zr)
    assert rel <= 0.03, rel
    assert d.mean() <= 0.006, (d.max(), d.mean())
    soft = (2 * (p * ref).sum() + 1) / ((p * p).sum() + (ref * ref).sum() + 1)
    assert soft >= 0.999, soft
    # batch invariance / determinism: sample 1 alone gives bit-identical output
    assert np.array_equal(model.predict(x[1:2]), p[1:2])


def test_train_step_matches_oracle(setup):
    from fetal_net.model import unet_model_3d
    _, w0 = setup
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 1, 32, 32, 32)).astype(np.float32)
    t = blob_target(x.shape, rng)
    model = unet_model_3d(input_shape=(1, 32, 32, 32), n_base_filters=16, depth=4, initial_learning_rate=1e-4)
    model.set_named_weights(w0)
    w = {k: v.copy() for k, v in w0.items()}
    state = {}
    ref = uo.unet3d_train_step(x, t, w, state, 1e-4)
    got = model.train_on_batch(x, t)
    assert got[0] == pytest.approx(ref["loss"], abs=3e-3), (got, ref["loss"])
    assert got[1] == pytest.approx(ref["binary_accuracy"], abs=5e-3)
    assert got[2] == pytest.approx(ref["vod_coefficient"], abs=5e-3)
    grads = model.get_gradients()
    floor = {"enc0a": 0.95, "enc0b": 0.97}
    report, bad = [], []
    for l, gk, gb in zip(model.layers, grads[0::2], grads[1::2]):
        for kind, g in (("kernel", gk), ("bias", gb)):
            r = ref["grads"]["%s/%s" % (l["name"], kind)].astype(np.float64).ravel()
            g = g.astype(np.float64).ravel()
            cos = float(g @ r / max(np.linalg.norm(g) * np.linalg.norm(r), 1e-300))
            ratio = float(np.linalg.norm(g) / max(np.linalg.norm(r), 1e-300))
            report.append((l["name"], kind, round(cos, 4), round(ratio, 4)))
            if not (cos >= floor.get(l["name"], 0.99) and 0.92 <= ratio <= 1.08):
                bad.append(report[-1])
    assert not bad, bad
    # Adam moved every weight by at most lr (first Keras-Adam step is lr * sign-like)
    new = model.get_weights()
    for l, k in zip(model.layers, new[0::2]):
        assert np.max(np.abs(k - w0[l["name"] + "/kernel"])) <= 1.01e-4


def test_loss_curve_tracks_oracle(setup):
    from fetal_net.model import unet_model_3d
    _, w0 = setup
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 1, 32, 32, 32)).astype(np.float32)
    t = blob_target(x.shape, rng)
    model = unet_model_3d(input_shape=(1, 32, 32, 32), n_base_filters=16, depth=4, initial_learning_rate=3e-4)
    model.set_named_weights(w0)
    w = {k: v.copy() for k, v in w0.items()}
    state = {}
    ours, refs = [], []
    for _ in range(6):
        refs.append(uo.unet3d_train_step(x, t, w, state, 3e-4)["loss"])
        ours.append(model.train_on_batch(x, t)[0])
    assert refs[-1] < refs[0] - 0.01 and ours[-1] < ours[0] - 0.01, (ours, refs)   # it learns
    assert np.max(np.abs(np.array(ours) - np.array(refs))) <= 0.02, (ours, refs)


def test_trained_network_mask_dice_vs_fp32_oracle():
    """north_star bar: Dice >= 0.999 between our mask and the fp32 reference's with the same weights, on a
    network that actually segments (trained here on the GPU for ~150 steps on a synthetic blob task)."""
    from fetal_net.model import unet_model_3d
    rng = np.random.default_rng(4)
    model = unet_model_3d(input_shape=(1, 32, 32, 32), n_base_filters=16, depth=4, initial_learning_rate=1e-3)
    model.init_glorot_uniform(seed=5)

    def batch(n):
        t = blob_target((n, 1, 32, 32, 32), rng)
        x = ((2 * t - 1) * 0.7 + 0.6 * rng.standard_normal(t.shape)).astype(np.float32)
        return x, t
    losses = []
    for _ in range(150):
        x, t = batch(4)
        losses.append(model.train_on_batch(x, t)[0])
    assert losses[-1] < -0.85, losses[::15]
    x, t = batch(3)
    p = model.predict(x)
    names = [l["name"] for l in model.layers]
    ws = model.get_weights()
    w = {}
    for n, k, b in zip(names, ws[0::2], ws[1::2]):
        w[n + "/kernel"], w[n + "/bias"] = k, b
    with torch.no_grad():
        ref = uo.unet3d_forward(torch.as_tensor(x), w).numpy()
    assert mask_dice(ref, t) > 0.9                       # the fp32 oracle agrees it is a segmentation
    md = mask_dice(p, ref)
    assert md >= 0.999, (md, float(np.abs(p - ref).max()))


_PIPELINE_SCRIPT = r"""
import json, sys
import numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
from fetal_net.model import unet_model_3d
from oracle import unet_oracle as uo
from tests.test_gpu_model import decisive_weights, blob_target
w = decisive_weights(uo.unet3d_layers(4, 16))
rng = np.random.default_rng(21)
xs = [rng.standard_normal((2, 1, 32, 32, 32)).astype(np.float32) for _ in range(4)]
ts = [blob_target(x.shape, rng) for x in xs]
losses, weights = [], []
for pinned in (False, True):
    model = unet_model_3d(input_shape=(1, 32, 32, 32), n_base_filters=16, depth=4, initial_learning_rate=1e-4)
    model.set_named_weights(w)
    ls = []
    for x, t in zip(xs, ts):
        if pinned:
            xp, tp = torch.as_tensor(x).pin_memory(), torch.as_tensor(t).pin_memory()
            ls.append(model.train_on_batch(xp.numpy(), tp.numpy())[0])
            xp.zero_(); tp.zero_()                      # inputs are free for reuse on return
        else:
            ls.append(model.train_on_batch(x, t)[0])
    losses.append(ls)
    weights.append(model.get_weights())
    p = model.predict(xs[0])
dw = max(float(np.abs(a - b).max()) for a, b in zip(weights[0], weights[1]))
print(json.dumps(dict(losses=losses, dw=dw)))
"""


def test_pipelined_train_step_with_pinned_inputs():
    """fm_train_step with page-locked inputs returns after the forward statistics are on the host and lets the rest of
    the step overlap the next upload: same losses / weights as the synchronous (pageable) route, and later calls
    (get_weights, predict) see the finished update. Runs in a subprocess with FETAL_B200_DETERMINISTIC=1 so that the
    two routes use bit-reproducible kernels and only the fp32 red.add order of the weight / bias gradients differs.
    That order is usually the same from run to run (tools/pipeline_noise.py: six runs agree to 1e-7) but not always,
    and Adam turns a gradient that is rounding noise around zero into a full +-lr step, so the bound is 1e-3 in the
    loss; a race in the pipelining (the inputs are zeroed right after every call) would show as O(0.1)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FETAL_B200_DETERMINISTIC="1")
    r = subprocess.run([sys.executable, "-c", _PIPELINE_SCRIPT, root, os.path.join(root, "fetal-mri-segmentation_b200")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    a, b = out["losses"]
    assert np.allclose(a, b, atol=1e-3), out
    assert out["dw"] <= 1.5e-3, out               # <= 4 Adam steps of 1e-4 in either direction


def test_evaluate_matches_host_metrics(setup):
    import fetal_net.metrics as fm
    model, _ = setup
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 1, 32, 32, 32)).astype(np.float32)
    t = blob_target(x.shape, rng)
    p = model.predict(x)
    loss, acc, vod = model.test_on_batch(x, t)
    assert loss == pytest.approx(fm.dice_coefficient_loss(t, p), abs=1e-5)
    assert acc == pytest.approx(fm.binary_accuracy(t, p), abs=1e-6)
    assert vod == pytest.approx(fm.vod_coefficient(t, p), abs=1e-5)


def test_backward_without_forward_is_an_error(setup):
    from fetal_net import _lib
    model, _ = setup
    model.predict(np.zeros((1, 1, 32, 32, 32), np.float32))
    with pytest.raises(_lib.FetalB200Error):
        _lib.check(_lib.load().fm_train_backward(model._h))


def test_bad_shapes_are_rejected():
    from fetal_net import _lib
    from fetal_net.model import unet_model_3d
    with pytest.raises(_lib.FetalB200Error):
        unet_model_3d(input_shape=(1, 30, 32, 32), n_base_filters=16)     # not divisible by 2^(depth-1)
    with pytest.raises(_lib.FetalB200Error):
        unet_model_3d(input_shape=(2, 32, 32, 32), n_base_filters=16)     # multi-modality not built


# ---- Isensee-2017 residual 3D U-Net (forward / inference) ------------------------------------------

def isensee_weights(layers, seed=0):
    w = uo.glorot_uniform_weights(layers, seed=seed)
    rng = np.random.default_rng(seed + 1)
    for name, cin, cout, k in layers:
        w[name + "/bias"] = (0.1 * rng.standard_normal(cout)).astype(np.float32)
        if not name.endswith("_seg"):
            w[name + "/gamma"] = (1.0 + 0.2 * rng.standard_normal(cout)).astype(np.float32)
            w[name + "/beta"] = (0.2 * rng.standard_normal(cout)).astype(np.float32)
        else:
            w[name + "/kernel"] = (w[name + "/kernel"] * 3.0).astype(np.float32)
    return w


@pytest.mark.parametrize("shape,depth,nseg", [((1, 32, 32, 32), 4, 2), ((1, 64, 64, 32), 5, 3)])
def test_isensee_forward_matches_oracle(shape, depth, nseg):
    from fetal_net.model import isensee2017_model_3d
    layers = uo.isensee3d_layers(depth, 16, nseg)
    w = isensee_weights(layers, seed=7)
    model = isensee2017_model_3d(input_shape=shape, n_base_filters=16, depth=depth, n_segmentation_levels=nseg)
    assert [l["name"] for l in model.layers if not l["is_norm"]] == [n for n, *_ in layers]
    model.set_named_weights(w)
    if depth == 5 and nseg == 3:
        conv_params = sum(l["k"] ** 3 * l["cin"] * l["cout"] + l["cout"] for l in model.layers if not l["is_norm"])
        assert conv_params == 8263619
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2,) + shape).astype(np.float32)
    p = model.predict(x)
    with torch.no_grad():
        ref = uo.isensee3d_forward(torch.as_tensor(x), w, depth=depth, n_segmentation_levels=nseg).numpy()
    assert p.shape == ref.shape
    logit = lambda q: np.log(np.clip(q.astype(np.float64), 1e-7, 1 - 1e-7) / np.clip(1 - q.astype(np.float64), 1e-7, 1))
    rel = np.linalg.norm(logit(p) - logit(ref)) / np.linalg.norm(logit(ref))
    # instance normalisation re-scales every block, so bf16 storage error does not shrink with depth: 4 % bound
    assert rel <= 0.04 and np.abs(p - ref).mean() <= 0.008, (rel, float(np.abs(p - ref).mean()))
    assert np.array_equal(model.predict(x[1:2]), p[1:2])              # per-sample statistics: batch invariant


def _blob_truth(x):
    """A smooth-ish binary target correlated with the input (so gradients are not degenerate)."""
    t = torch.nn.functional.avg_pool3d(torch.as_tensor(x), 5, stride=1, padding=2).numpy()
    return (t > 0.05).astype(np.float32)


@pytest.mark.parametrize("shape,depth,nseg", [((1, 32, 32, 16), 3, 2), ((1, 32, 32, 32), 4, 3)])
def test_isensee_train_step_matches_oracle(shape, depth, nseg):
    """fwd + Dice + bwd + Adam of the Isensee net vs torch autograd on the fp32 restatement (dropout off): loss,
    per-layer gradient direction (kernels, gamma, beta), updated weights, and a falling loss."""
    from fetal_net.model import isensee2017_model_3d
    layers = uo.isensee3d_layers(depth, 16, nseg)
    w = isensee_weights(layers, seed=11)
    model = isensee2017_model_3d(input_shape=shape, n_base_filters=16, depth=depth, n_segmentation_levels=nseg,
                                 dropout_rate=0, initial_learning_rate=1e-3)
    model.set_named_weights(w)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2,) + shape).astype(np.float32)
    t = _blob_truth(x)
    wo = {k: v.copy() for k, v in w.items()}
    ref = uo.train_step(lambda xt, prm: uo.isensee3d_forward(xt, prm, depth=depth, n_segmentation_levels=nseg),
                        x, t, wo, {}, 1e-3)
    loss, acc, vod = model.train_on_batch(x, t)
    assert abs(loss - ref["loss"]) <= 4e-3, (loss, ref["loss"])
    grads = model.get_gradients()
    cos = {}
    for l, gk, gb in zip(model.layers, grads[0::2], grads[1::2]):
        if l["is_norm"]:
            base = l["name"][:-len("_norm")]
            pairs = [(base + "/gamma", gk), (base + "/beta", gb)]
        else:
            rk = ref["grads"][l["name"] + "/kernel"]                   # Keras layout, like get_gradients()
            assert rk.shape == gk.shape, (l["name"], rk.shape, gk.shape)
            cos[l["name"] + "/kernel"] = float((gk * rk).sum() / (np.linalg.norm(gk) * np.linalg.norm(rk) + 1e-30))
            if l["name"].endswith("_seg"):
                pairs = [(l["name"] + "/bias", gb)]
            else:
                # a bias in front of an instance norm has an exactly-zero gradient: only rounding noise remains
                assert np.abs(gb).max() <= 1e-2 * max(np.abs(gk).max(), 1e-12), l["name"]
                pairs = []
        for name, g in pairs:
            r = ref["grads"][name]
            cos[name] = float((g * r).sum() / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30))
    worst = min(cos.items(), key=lambda kv: kv[1])
    assert worst[1] >= 0.97, (worst, sorted(cos.items(), key=lambda kv: kv[1])[:6])
    assert np.median(list(cos.values())) >= 0.99
    # a few more steps: the loss falls like the oracle's
    losses = [loss]
    for _ in range(5):
        losses.append(model.train_on_batch(x, t)[0])
    assert losses[-1] < losses[0] - 1e-3, losses


def test_isensee_dropout_training_only():
    from fetal_net.model import isensee2017_model_3d
    shape = (1, 32, 32, 16)
    kw = dict(input_shape=shape, n_base_filters=16, depth=3, n_segmentation_levels=1, initial_learning_rate=1e-3)
    a = isensee2017_model_3d(dropout_rate=0.0, **kw)
    b = isensee2017_model_3d(dropout_rate=0.5, **kw)
    a.init_glorot_uniform(seed=2)
    b.set_weights(a.get_weights())
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2,) + shape).astype(np.float32)
    t = _blob_truth(x)
    assert np.array_equal(a.predict(x), b.predict(x))                 # identity at inference
    # validation metrics (fit_generator's test_on_batch) come from the inference kernels: equal to host metrics of predict
    from fetal_net import metrics as hm
    ev = b.test_on_batch(x, t)
    p = b.predict(x)
    assert abs(ev[0] - hm.dice_coefficient_loss(t, p)) <= 1e-5 and abs(ev[2] - hm.vod_coefficient(t, p)) <= 1e-5, ev
    la, lb = a.train_on_batch(x, t)[0], b.train_on_batch(x, t)[0]
    assert np.isfinite(lb) and la != lb                               # masks change the training forward
    ga, gb = a.get_gradients(), b.get_gradients()
    assert all(np.isfinite(g).all() for g in gb)
    assert any(not np.allclose(p, q) for p, q in zip(ga, gb))


@pytest.mark.parametrize("tag", ["unet3d_d4_nf16", "isensee3d_d3_nf8_seg2", "unet2d_d3_nf16"])
def test_gpu_matches_keras_fixture(tag):
    """CUDA path against REAL Keras output (tests/golden/keras_fixture.npz, written by tools/export_keras_fixture.py
    where Keras/TF exist): same weights, same input -> north_star's bar (logits within the stated bf16 tolerance, soft
    Dice >= 0.999), and the loss of one train_on_batch within 3e-3. Skips while the fixture is absent."""
    from oracle import keras_fixture as kf
    from fetal_net.model import unet_model_3d, unet_model_2d, isensee2017_model_3d
    z = kf.load()
    if z is None or tag + "/names" not in z.files:
        pytest.skip("tests/golden/keras_fixture.npz absent (needs Keras/TF: tools/export_keras_fixture.py)")
    x, t, ref = z[tag + "/x"], z[tag + "/t"], z[tag + "/predict"]
    lr = float(z[tag + "/lr"])
    if tag.startswith("unet3d"):
        layers = uo.unet3d_layers(4, 16)
        model = unet_model_3d(input_shape=x.shape[1:], depth=4, n_base_filters=16, initial_learning_rate=lr)
    elif tag.startswith("isensee"):
        layers = uo.isensee3d_layers(3, 8, 2)
        model = isensee2017_model_3d(input_shape=x.shape[1:], depth=3, n_base_filters=8, n_segmentation_levels=2,
                                     dropout_rate=0.0, initial_learning_rate=lr)
    else:
        layers = uo.unet2d_layers(3, 16, 6)
        model = unet_model_2d(input_shape=x.shape[1:], depth=3, n_base_filters=16, initial_learning_rate=lr)
    model.set_named_weights(kf.named_weights(z, tag, layers))
    p = model.predict(x)
    lg = lambda q: np.log(np.clip(q.astype(np.float64), 1e-7, 1 - 1e-7) / np.clip(1 - q.astype(np.float64), 1e-7, 1))
    err = np.linalg.norm(lg(p) - lg(ref)) / np.linalg.norm(lg(ref))
    soft = (2 * (p * ref).sum() + 1) / ((p * p).sum() + (ref * ref).sum() + 1)
    assert err <= 0.04 and np.abs(p - ref).mean() <= 0.006 and soft >= 0.999, (err, soft)
    loss = model.train_on_batch(x, t)[0]
    assert abs(loss - float(z[tag + "/train_metrics"][0])) <= 3e-3


def _grad_cosines(model, ref_grads):
    cos = {}
    grads = model.get_gradients()
    for l, gk, gb in zip(model.layers, grads[0::2], grads[1::2]):
        if l["is_norm"]:
            base = l["name"][:-len("_norm")]
            pairs = [(base + "/gamma", gk), (base + "/beta", gb)]
        else:
            pairs = [(l["name"] + "/kernel", gk)]
        for name, g in pairs:
            r = ref_grads[name].astype(np.float64).ravel()
            g = g.astype(np.float64).ravel()
            cos[name] = float(g @ r / max(np.linalg.norm(g) * np.linalg.norm(r), 1e-300))
    return cos


def test_unet3d_dice_and_xent_train_step_matches_oracle(setup):
    """loss_function=dice_and_xent (fetal_net/metrics.py:68-78, config_utils.py:77): the combined loss, the soft Dice
    metric Keras appends for a non-Dice loss, and the gradients against torch autograd on the fp32 restatement."""
    import functools
    from fetal_net import metrics as fmet
    from fetal_net.model import unet_model_3d
    _, w0 = setup
    rng = np.random.default_rng(21)
    x = rng.standard_normal((2, 1, 32, 32, 32)).astype(np.float32)
    t = blob_target(x.shape, rng)
    for loss_fn, xw in ((fmet.dice_and_xent, 1.0), (functools.partial(fmet.dice_and_xent, xent_weight=0.25), 0.25)):
        model = unet_model_3d(input_shape=(1, 32, 32, 32), n_base_filters=16, depth=4, initial_learning_rate=1e-4,
                              loss_function=loss_fn)
        assert model.metrics_names == ['loss', 'binary_accuracy', 'vod_coefficient', 'dice_coefficient']
        model.set_named_weights(w0)
        w = {k: v.copy() for k, v in w0.items()}
        ref = uo.train_step(lambda xt, prm: uo.unet3d_forward(xt, prm, depth=4), x, t, w, {}, 1e-4,
                            loss_fn=lambda tt, pp: uo.dice_and_xent(tt, pp, xw))
        got = model.train_on_batch(x, t)
        assert len(got) == 4
        assert got[0] == pytest.approx(ref["loss"], abs=4e-3), (got, ref["loss"])
        p_ref = ref["pred"]
        assert got[3] == pytest.approx(fmet.dice_coefficient(t, p_ref), abs=3e-3)
        # host evaluation helper == oracle loss on the oracle's own prediction
        assert loss_fn(t, p_ref) == pytest.approx(ref["loss"], abs=1e-5)
        cos = _grad_cosines(model, ref["grads"])
        floor = {"enc0a/kernel": 0.95, "enc0b/kernel": 0.97}
        bad = {k: v for k, v in cos.items() if v < floor.get(k, 0.99)}
        assert not bad, bad
        # evaluate() reports the same combined loss on the inference kernels
        ev = model.test_on_batch(x, t)
        assert len(ev) == 4 and np.isfinite(ev).all()


def test_isensee_dice_and_xent_mask_two_input_model():
    """isensee2017_model_3d(..., loss_function=dice_and_xent_mask, mask_shape=input_shape) (train_fetal.py:31-39 with
    config['weight_mask'], isensee2017.py:85-88): the model takes [x, weight_mask]; loss and gradients against torch
    autograd with weight exp(-mask / 3)."""
    from fetal_net import metrics as fmet
    from fetal_net.model import isensee2017_model_3d
    shape, depth, nseg = (1, 32, 32, 16), 3, 2
    layers = uo.isensee3d_layers(depth, 16, nseg)
    w = isensee_weights(layers, seed=13)
    model = isensee2017_model_3d(input_shape=shape, n_base_filters=16, depth=depth, n_segmentation_levels=nseg,
                                 dropout_rate=0, initial_learning_rate=1e-3, loss_function=fmet.dice_and_xent_mask,
                                 mask_shape=shape)
    model.set_named_weights(w)
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2,) + shape).astype(np.float32)
    t = _blob_truth(x)
    mask = (8.0 * rng.random((2,) + shape)).astype(np.float32)          # distance-to-border map
    with pytest.raises(AssertionError):
        model.train_on_batch(x, t)                                       # the mask input is mandatory
    wo = {k: v.copy() for k, v in w.items()}
    mt = torch.as_tensor(mask)
    ref = uo.train_step(lambda xt, prm: uo.isensee3d_forward(xt, prm, depth=depth, n_segmentation_levels=nseg),
                        x, t, wo, {}, 1e-3, loss_fn=lambda tt, pp: uo.dice_and_xent(tt, pp, 1.0, mt, 3.0))
    got = model.train_on_batch([x, mask], t)
    assert len(got) == 4 and abs(got[0] - ref["loss"]) <= 5e-3, (got, ref["loss"])
    assert fmet.dice_and_xent_mask(mask)(t, ref["pred"]) == pytest.approx(ref["loss"], abs=1e-5)
    cos = _grad_cosines(model, ref["grads"])
    worst = min(cos.items(), key=lambda kv: kv[1])
    assert worst[1] >= 0.97 and np.median(list(cos.values())) >= 0.99, worst
    # predict ignores the mask input; a different mask changes the loss
    assert np.array_equal(model.predict([x, mask]), model.predict(x))
    ev0 = model.test_on_batch([x, mask], t)
    ev1 = model.test_on_batch([x, np.zeros_like(mask)], t)
    assert ev1[0] > ev0[0] + 1e-3                                        # weight 1 everywhere > exp(-mask/3)


# ---- Isensee-2017 in 2D: isensee2017_model (fetal_net/model/unet/isensee.py) -------------------------------------

def _isensee2d_weights(layers, seed):
    w = uo.glorot_uniform_weights(layers, seed=seed, ndim=2)
    rng = np.random.default_rng(seed + 1)
    for name, cin, cout, k in layers:
        w[name + "/bias"] = (0.1 * rng.standard_normal(cout)).astype(np.float32)
        if not name.endswith("_seg"):
            w[name + "/gamma"] = (1.0 + 0.2 * rng.standard_normal(cout)).astype(np.float32)
            w[name + "/beta"] = (0.2 * rng.standard_normal(cout)).astype(np.float32)
        else:
            w[name + "/kernel"] = (w[name + "/kernel"] * 3.0).astype(np.float32)
    return w


@pytest.mark.parametrize("shape,depth,nseg,summation", [((32, 32, 5), 3, 3, False), ((64, 64, 6), 4, 2, True)])
def test_isensee2d_forward_and_train_step_match_oracle(shape, depth, nseg, summation):
    """The 2D Isensee builder (config_utils.py:66-69 model_name 'isensee2017_model'): Conv2D blocks with instance norm,
    strides (2,2), UpSampling2D; summation=False keeps only the finest head (the reference default), summation=True
    sums the heads coarse to fine. Forward, loss and gradients against torch autograd on the fp32 restatement."""
    from fetal_net.model import isensee2017_model
    heads = nseg if summation else 1
    layers = uo.isensee2d_layers(depth, 16, heads, shape[2])
    w = _isensee2d_weights(layers, seed=5)
    model = isensee2017_model(input_shape=shape, n_base_filters=16, depth=depth, n_segmentation_levels=nseg,
                              summation=summation, dropout_rate=0, initial_learning_rate=1e-3)
    assert [l["name"] for l in model.layers if not l["is_norm"]] == [n for n, *_ in layers]
    assert model.layers[0]["kshape"] == (3, 3, shape[2], 16) and model.output_shape == (None,) + shape[:2] + (1,)
    assert model.count_params() == sum(v.size for v in w.values())
    model.set_named_weights(w)
    back = model.get_weights()
    assert np.array_equal(back[0], w["l0_in/kernel"])                        # (3,3,Cin,Cout) round trip, Cin padded inside
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2,) + shape).astype(np.float32)
    t = (torch.nn.functional.avg_pool2d(torch.as_tensor(x[..., shape[2] // 2])[:, None], 5, stride=1, padding=2)
         .numpy()[:, 0, :, :, None] > 0.05).astype(np.float32)
    fwd = lambda xt, prm: uo.isensee2d_forward(xt, prm, depth=depth, n_heads=heads)
    p = model.predict(x)
    with torch.no_grad():
        ref = fwd(torch.as_tensor(x), w).numpy()
    assert p.shape == ref.shape == (2,) + shape[:2] + (1,)
    logit = lambda q: np.log(np.clip(q.astype(np.float64), 1e-7, 1 - 1e-7) / np.clip(1 - q.astype(np.float64), 1e-7, 1))
    rel = np.linalg.norm(logit(p) - logit(ref)) / np.linalg.norm(logit(ref))
    assert rel <= 0.04 and np.abs(p - ref).mean() <= 0.008, (rel, float(np.abs(p - ref).mean()))
    refstep = uo.train_step(fwd, x, t, {k: v.copy() for k, v in w.items()}, {}, 1e-3)
    loss = model.train_on_batch(x, t)[0]
    assert abs(loss - refstep["loss"]) <= 4e-3, (loss, refstep["loss"])
    cos = _grad_cosines(model, refstep["grads"])
    worst = min(cos.items(), key=lambda kv: kv[1])
    assert worst[1] >= 0.97 and np.median(list(cos.values())) >= 0.99, (worst, sorted(cos.items(), key=lambda kv: kv[1])[:5])
    losses = [loss] + [model.train_on_batch(x, t)[0] for _ in range(5)]
    assert losses[-1] < losses[0] - 1e-3, losses


def test_isensee2d_dropout_and_patchwise():
    """SpatialDropout2D of the context modules (unet/isensee.py:105): identity at inference, active in training; and
    the 2.5D sliding window over a volume drives the 2D Isensee model like the 2D U-Net."""
    from fetal_net.model import isensee2017_model
    from fetal_net.prediction import patch_wise_prediction
    from oracle import prediction_oracle as po
    kw = dict(input_shape=(32, 32, 5), n_base_filters=16, depth=3, initial_learning_rate=1e-3)
    a = isensee2017_model(dropout_rate=0.0, **kw)
    b = isensee2017_model(dropout_rate=0.5, **kw)
    a.init_glorot_uniform(seed=2)
    b.set_weights(a.get_weights())
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 32, 32, 5)).astype(np.float32)
    t = (rng.random((2, 32, 32, 1)) < 0.3).astype(np.float32)
    assert np.array_equal(a.predict(x), b.predict(x))
    la, lb = a.train_on_batch(x, t)[0], b.train_on_batch(x, t)[0]
    assert np.isfinite(lb) and la != lb
    vol = rng.standard_normal((1, 48, 32, 9)).astype(np.float32)
    out = patch_wise_prediction(a, vol, patch_shape=(32, 32, 5), overlap_factor=0.5, batch_size=4)
    ref = po.patch_wise_prediction(a, vol, (32, 32, 5), overlap_factor=0.5, batch_size=4)
    assert out.shape == (48, 32, 9, 1) and np.array_equal(out, ref)


# ---- deconvolution=True: Deconvolution3D / Deconvolution2D instead of UpSampling (unet3d/unet.py:57-59,132-136) ---------

def _cos(g, r):
    g, r = g.astype(np.float64).ravel(), r.astype(np.float64).ravel()
    return float(g @ r / max(np.linalg.norm(g) * np.linalg.norm(r), 1e-300))


def _unet_grad_check(model, ref_grads, bf16_grads):
    """Per-layer gradient direction against the fp32 oracle. The yardstick is the storage format: `bf16_grads` are the
    gradients of the SAME fp32 oracle with its stored activations / weights / gradients rounded to bfloat16; a layer must
    reach min(0.99, that oracle's own cosine against fp32 - 0.02) (0.95 at most for the two earliest layers, as in
    test_train_step_matches_oracle), and the norm must agree within 10 %."""
    bad = []
    grads = model.get_gradients()
    for l, gk, gb in zip(model.layers, grads[0::2], grads[1::2]):
        for kind, g in (("kernel", gk), ("bias", gb)):
            name = "%s/%s" % (l["name"], kind)
            r = ref_grads[name]
            cos = _cos(g, r)
            # training passes are not bit-reproducible run to run: 0.02 of head-room below the yardstick
            floor = min(0.95 if l["name"] in ("enc0a", "enc0b") else 0.99, _cos(bf16_grads[name], r) - 0.02)
            ratio = float(np.linalg.norm(g.astype(np.float64)) / max(np.linalg.norm(r.astype(np.float64)), 1e-300))
            if not (cos >= floor and 0.9 <= ratio <= 1.1):
                bad.append((name, round(cos, 4), round(floor, 4), round(ratio, 4)))
    return bad


def test_unet3d_deconvolution_matches_oracle():
    from fetal_net.model import unet_model_3d
    layers = uo.unet3d_layers(4, 16, deconvolution=True)
    w = decisive_weights(layers, seed=4)
    for d in range(3):                                            # a linear layer: keep the variance (fan-in = C per class)
        w["up%d/kernel" % d] = (w["up%d/kernel" % d] * 2.0).astype(np.float32)
    model = unet_model_3d(input_shape=(1, 32, 32, 32), n_base_filters=16, depth=4, initial_learning_rate=1e-4,
                          deconvolution=True)
    assert [l["name"] for l in model.layers] == [n for n, *_ in layers]
    up2 = [l for l in model.layers if l["name"] == "up2"][0]
    assert up2["kshape"] == (2, 2, 2, 256, 256) and up2["keras_name"] == "conv3d_transpose_1"
    assert [l["keras_name"] for l in model.layers if l["name"] in ("dec2a", "final")] == ["conv3d_9", "conv3d_15"]
    assert model.count_params() == sum(v.size for v in w.values())
    model.set_named_weights(w)
    assert np.array_equal(model.get_weights()[2 * up2["index"]], w["up2/kernel"])
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 1, 32, 32, 32)).astype(np.float32)
    t = blob_target(x.shape, rng)
    with torch.no_grad():
        ref = uo.unet3d_forward(torch.as_tensor(x), w).numpy()
    p = model.predict(x)
    lg = lambda q: np.log(np.clip(q.astype(np.float64), 1e-7, 1 - 1e-7) / np.clip(1 - q.astype(np.float64), 1e-7, 1))
    err = np.linalg.norm(lg(p) - lg(ref)) / np.linalg.norm(lg(ref))
    soft = (2 * (p * ref).sum() + 1) / ((p * p).sum() + (ref * ref).sum() + 1)
    assert err <= 0.03 and np.abs(p - ref).mean() <= 0.006 and soft >= 0.999, (err, soft)
    refstep = uo.unet3d_train_step(x, t, {k: v.copy() for k, v in w.items()}, {}, 1e-4)
    bfstep = uo.unet3d_train_step(x, t, {k: v.copy() for k, v in w.items()}, {}, 1e-4, quant=uo.bf16_round)
    got = model.train_on_batch(x, t)
    assert got[0] == pytest.approx(refstep["loss"], abs=3e-3), (got, refstep["loss"])
    bad = _unet_grad_check(model, refstep["grads"], bfstep["grads"])
    assert not bad, bad


def test_unet2d_deconvolution_matches_oracle():
    from fetal_net.model import unet_model_2d
    depth, nf = 3, 32
    layers = uo.unet2d_layers(depth, nf, 6, deconvolution=True)
    w = uo.glorot_uniform_weights(layers, seed=9, ndim=2)
    rng = np.random.default_rng(10)
    for k in w:
        if k.endswith("/kernel"):
            w[k] = (w[k] * (2.0 if k.startswith("up") else np.sqrt(2.0) * 1.2)).astype(np.float32)
        else:
            w[k] = (0.05 * rng.standard_normal(w[k].shape)).astype(np.float32)
    model = unet_model_2d(input_shape=(32, 32, 6), n_base_filters=nf, depth=depth, initial_learning_rate=1e-4,
                          deconvolution=True)
    assert [l["name"] for l in model.layers] == [n for n, *_ in layers]
    assert [l["kshape"] for l in model.layers if l["name"] == "up0"] == [(2, 2, 128, 128)]   # channels of dec1b
    model.set_named_weights(w)
    x = rng.standard_normal((2, 32, 32, 6)).astype(np.float32)
    t = (rng.random((2, 32, 32, 1)) < 0.3).astype(np.float32)
    fwd = lambda xt, prm: uo.unet2d_forward(xt, prm, depth=depth)
    with torch.no_grad():
        ref = fwd(torch.as_tensor(x), w).numpy()
    p = model.predict(x)
    lg = lambda q: np.log(np.clip(q.astype(np.float64), 1e-7, 1 - 1e-7) / np.clip(1 - q.astype(np.float64), 1e-7, 1))
    err = np.linalg.norm(lg(p) - lg(ref)) / np.linalg.norm(lg(ref))
    assert err <= 0.03 and np.abs(p - ref).mean() <= 0.006, err
    refstep = uo.train_step(fwd, x, t, {k: v.copy() for k, v in w.items()}, {}, 1e-4)
    bfstep = uo.train_step(lambda xt, prm: uo.unet2d_forward(xt, prm, depth=depth, quant=uo.bf16_round), x, t,
                           {k: v.copy() for k, v in w.items()}, {}, 1e-4)
    got = model.train_on_batch(x, t)
    assert got[0] == pytest.approx(refstep["loss"], abs=3e-3), (got, refstep["loss"])
    bad = _unet_grad_check(model, refstep["grads"], bfstep["grads"])
    assert not bad, bad


# ---- batch_normalization=True: Conv -> BatchNormalization(axis=1) -> ReLU blocks (unet3d/unet.py:102-113) ---------------

def _bn_weights(layers, ndim, seed):
    w = uo.glorot_uniform_weights(layers, seed=seed, ndim=ndim)
    w.update(uo.bn_params(layers))
    rng = np.random.default_rng(seed + 1)
    for name, cin, cout, k in layers:
        w[name + "/bias"] = (0.1 * rng.standard_normal(cout)).astype(np.float32)
        if k == 3:
            w[name + "/gamma"] = (1.0 + 0.2 * rng.standard_normal(cout)).astype(np.float32)
            w[name + "/beta"] = (0.2 * rng.standard_normal(cout)).astype(np.float32)
            w[name + "/moving_mean"] = (0.1 * rng.standard_normal(cout)).astype(np.float32)
            w[name + "/moving_variance"] = (0.5 + rng.random(cout)).astype(np.float32)
    w["final/kernel"] = (w["final/kernel"] * 3.0).astype(np.float32)
    return w


def _bn_grad_check(model, ref_grads, bf_grads):
    bad = []
    grads = model.get_gradients()
    for l, gk, gb in zip(model.layers, grads[0::2], grads[1::2]):
        if l["is_moving"]:
            assert not gk.any() and not gb.any(), l["name"]          # non-trainable
            continue
        if l["is_norm"]:
            base = l["name"][:-len("_norm")]
            pairs = [(base + "/gamma", gk), (base + "/beta", gb)]
        elif (l["name"] + "/gamma") in ref_grads:                    # conv bias in front of a BN: analytically zero
            assert np.abs(gb).max() <= 2e-2 * max(np.abs(gk).max(), 1e-12), l["name"]
            pairs = [(l["name"] + "/kernel", gk)]
        else:
            pairs = [(l["name"] + "/kernel", gk), (l["name"] + "/bias", gb)]
        for name, g in pairs:
            r = ref_grads[name]
            cos = _cos(g, r)
            # batch statistics couple every voxel of a channel: 0.02 below the bf16-storage oracle's own score
            floor = min(0.99, _cos(bf_grads[name], r) - 0.02)
            if cos < floor:
                bad.append((name, round(cos, 4), round(floor, 4)))
    return bad


@pytest.mark.parametrize("deconvolution", [False, True])
def test_unet3d_batch_normalization_matches_oracle(deconvolution):
    """unet_model_3d(batch_normalization=True): inference on the moving statistics, a training step on the batch
    statistics (loss, gradients incl. gamma / beta), the Keras moving-average update, and inference after it."""
    from fetal_net.model import unet_model_3d
    depth, nf, shape = 3, 16, (1, 32, 32, 16)
    layers = uo.unet3d_layers(depth, nf, deconvolution=deconvolution)
    w = _bn_weights(layers, 3, seed=21)
    model = unet_model_3d(input_shape=shape, n_base_filters=nf, depth=depth, initial_learning_rate=1e-3,
                          batch_normalization=True, deconvolution=deconvolution)
    names = [l["name"] for l in model.layers]
    assert names[:3] == ["enc0a", "enc0a_norm", "enc0a_moving"] and names[-1] == "final"
    bnl = [l for l in model.layers if l["name"] == "enc1a_moving"][0]
    assert bnl["keras_name"] == "batch_normalization_3" and bnl["keys"] == ("/moving_mean:0", "/moving_variance:0")
    assert model.count_params() == sum(v.size for v in w.values())
    model.set_named_weights(w)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((4,) + shape).astype(np.float32)
    t = _blob_truth(x)
    lg = lambda q: np.log(np.clip(q.astype(np.float64), 1e-7, 1 - 1e-7) / np.clip(1 - q.astype(np.float64), 1e-7, 1))

    def check_predict(weights):
        with torch.no_grad():
            ref = uo.unet3d_forward(torch.as_tensor(x), weights, depth=depth).numpy()
        p = model.predict(x)
        rel = np.linalg.norm(lg(p) - lg(ref)) / np.linalg.norm(lg(ref))
        # 5 %: every block is re-scaled by 1 / sqrt(moving variance), which amplifies the bf16 rounding of the raw conv
        # output (the instance-normalised nets carry 4 % for the same reason)
        assert rel <= 0.05 and np.abs(p - ref).mean() <= 0.008, (rel, float(np.abs(p - ref).mean()))

    check_predict(w)                                                   # moving statistics
    wo = {k: v.copy() for k, v in w.items()}
    ref = uo.unet3d_train_step(x, t, wo, {}, 1e-3, depth=depth)
    bf = uo.unet3d_train_step(x, t, {k: v.copy() for k, v in w.items()}, {}, 1e-3, depth=depth, quant=uo.bf16_round)
    got = model.train_on_batch(x, t)
    assert abs(got[0] - ref["loss"]) <= 4e-3, (got, ref["loss"])
    bad = _bn_grad_check(model, ref["grads"], bf["grads"])
    assert not bad, bad
    # moving statistics after the step: 0.99 * old + 0.01 * batch value (variance with Keras' sample-size correction)
    new = dict(zip([l["name"] for l in model.layers], zip(model.get_weights()[0::2], model.get_weights()[1::2])))
    for name, cin, cout, k in layers:
        if k != 3:
            continue
        mm, mv = new[name + "_moving"]
        assert np.abs(mm - wo[name + "/moving_mean"]).max() <= 2e-3 * max(1.0, np.abs(wo[name + "/moving_mean"]).max()), name
        assert np.abs(mv / wo[name + "/moving_variance"] - 1).max() <= 5e-3, name
        assert np.abs(mm - w[name + "/moving_mean"]).max() > 0          # it did move
    # inference after the step runs on the UPDATED weights and moving statistics: compare against the oracle loaded with
    # the library's own post-step state (the first Adam step moves every weight by ~lr in the direction of the
    # gradient's sign, so the two sides' weights differ by up to 2 lr wherever a tiny gradient flips sign)
    own = {}
    ws = model.get_weights()
    for l, a, b in zip(model.layers, ws[0::2], ws[1::2]):
        if l["is_moving"]:
            base = l["name"][:-len("_moving")]
            own[base + "/moving_mean"], own[base + "/moving_variance"] = a, b
        elif l["is_norm"]:
            base = l["name"][:-len("_norm")]
            own[base + "/gamma"], own[base + "/beta"] = a, b
        else:
            own[l["name"] + "/kernel"], own[l["name"] + "/bias"] = a, b
    check_predict(own)
    losses = [got[0]] + [model.train_on_batch(x, t)[0] for _ in range(5)]
    assert losses[-1] < losses[0] - 1e-3, losses


def test_unet2d_batch_normalization_matches_oracle():
    from fetal_net.model import unet_model_2d
    depth, nf = 3, 32
    layers = uo.unet2d_layers(depth, nf, 6)
    w = _bn_weights(layers, 2, seed=31)
    model = unet_model_2d(input_shape=(32, 32, 6), n_base_filters=nf, depth=depth, initial_learning_rate=1e-3,
                          batch_normalization=True)
    model.set_named_weights(w)
    rng = np.random.default_rng(6)
    x = rng.standard_normal((4, 32, 32, 6)).astype(np.float32)
    t = (rng.random((4, 32, 32, 1)) < 0.3).astype(np.float32)
    lg = lambda q: np.log(np.clip(q.astype(np.float64), 1e-7, 1 - 1e-7) / np.clip(1 - q.astype(np.float64), 1e-7, 1))
    with torch.no_grad():
        ref = uo.unet2d_forward(torch.as_tensor(x), w, depth=depth).numpy()
    p = model.predict(x)
    rel = np.linalg.norm(lg(p) - lg(ref)) / np.linalg.norm(lg(ref))
    assert rel <= 0.04 and np.abs(p - ref).mean() <= 0.008, rel
    upd = {}
    fwd = lambda xt, prm: uo.unet2d_forward(xt, prm, depth=depth, training=True, bn_updates=upd)
    refstep = uo.train_step(fwd, x, t, {k: v.copy() for k, v in w.items()}, {}, 1e-3)
    bfstep = uo.train_step(lambda xt, prm: uo.unet2d_forward(xt, prm, depth=depth, training=True, quant=uo.bf16_round),
                           x, t, {k: v.copy() for k, v in w.items()}, {}, 1e-3)
    got = model.train_on_batch(x, t)
    assert abs(got[0] - refstep["loss"]) <= 4e-3, (got, refstep["loss"])
    bad = _bn_grad_check(model, refstep["grads"], bfstep["grads"])
    assert not bad, bad
    new = dict(zip([l["name"] for l in model.layers], zip(model.get_weights()[0::2], model.get_weights()[1::2])))
    for name in ("enc0a", "dec0b"):
        assert np.abs(new[name + "_moving"][1] / upd[name + "/moving_variance"] - 1).max() <= 5e-3, name


def test_fast_inference_mode_stays_within_tolerance(setup):
    """Model.set_fast_inference (fm_model_set_inference_mode): the three-issuer conv mode in inference differs from
    the reproducible default only in the accumulation order - last-bit flips of bf16 activations that the decisive
    test weights amplify to the same order as the stated tolerance against the oracle (mean |dp| <= 0.006); switching back restores bit-identical results."""
    model, w = setup
    rng = np.random.default_rng(11)
    x = rng.standard_normal((2, 1, 32, 32, 32)).astype(np.float32)
    p0 = model.predict(x)
    assert np.array_equal(model.predict(x), p0)
    model.set_fast_inference(True)
    p1 = model.predict(x)
    model.set_fast_inference(False)
    assert np.array_equal(model.predict(x), p0)
    with torch.no_grad():
        ref = uo.unet3d_forward(torch.as_tensor(x), w).numpy()
    # (single voxels next to a decision boundary move by a few percent when a bf16 activation flips its last bit)
    assert np.abs(p1 - p0).mean() <= 0.006 and np.abs(p1 - p0).max() <= 0.1 and np.abs(p1 - ref).mean() <= 0.006
