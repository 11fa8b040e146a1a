"""CPU-side checks: the C-ABI library loads and exports every symbol include/fetal_b200.h declares,
the host-only entry point fm_patch_plan reproduces the reference plans, and the Python mirror of the
reference interface keeps its names/signatures. No GPU compute here."""
import ctypes
import hashlib
import inspect
import os
import re

import numpy as np
import pytest

from oracle import prediction_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fetal_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from fetal_net import _lib
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), "libfetalb200.so does not export %s" % s
        assert s in _lib.SIGNATURES, "ctypes binding missing for %s" % s
    assert sorted(_lib.SIGNATURES) == syms


def test_missing_library_fails_loudly(monkeypatch):
    from fetal_net import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libfetalb200.so")
    with pytest.raises(_lib.FetalB200Error):
        _lib.load()


def test_ctx_create_without_gpu_reports_error(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.fm_ctx_create(0, ctypes.byref(h))
    assert rc != 0 and lib.fm_last_error()


def test_fm_patch_plan_matches_golden_and_oracle(golden):
    from fetal_net.prediction import patch_plan, get_set_of_patch_indices_full
    names = sorted({k.split("/")[1] for k in golden if k.startswith("plan/")})
    for name in names:
        a = golden["plan/%s/args" % name]
        padded, patch, pshape = tuple(int(v) for v in a[0:3]), tuple(int(v) for v in a[3:6]), tuple(int(v) for v in a[6:9])
        f = float(golden["plan/%s/f" % name])
        idx = patch_plan(padded, patch, pshape, f)
        assert idx.dtype == np.int32 and idx.shape == (int(golden["plan/%s/n" % name]), 3), name
        assert sha16(idx) == str(golden["plan/%s/sha" % name]), name
        assert np.array_equal(idx, po.patch_plan(padded, patch, pshape, f)), name
        ov = po.compute_overlap(patch, pshape, f)
        assert np.array_equal(idx, get_set_of_patch_indices_full((0, 0, 0), np.subtract(padded, patch),
                                                                 np.subtract(patch, ov))), name


def test_fm_patch_plan_edge_cases():
    from fetal_net import _lib
    from fetal_net.prediction import patch_plan
    # volume == patch: one patch; overlap 1.0 -> step 1 (predict.main default)
    assert patch_plan((64, 64, 64), (64, 64, 64), (64, 64, 64), 0.5).tolist() == [[0, 0, 0]]
    idx = patch_plan((20, 16, 16), (16, 16, 16), (16, 16, 16), 1.0)
    assert idx[:, 0].tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(_lib.FetalB200Error):      # padded smaller than the patch
        patch_plan((8, 64, 64), (64, 64, 64), (64, 64, 64), 0.5)
    # randomised agreement with the oracle
    rng = np.random.default_rng(0)
    for _ in range(200):
        patch = tuple(int(v) for v in rng.integers(1, 20, 3))
        padded = tuple(int(p + v) for p, v in zip(patch, rng.integers(0, 40, 3)))
        f = float(rng.choice([0.0, 0.25, 0.5, 0.77, 0.9, 1.0, rng.random()]))
        assert np.array_equal(patch_plan(padded, patch, patch, f), po.patch_plan(padded, patch, patch, f))


def test_geometry_matches_oracle_padding():
    from fetal_net.prediction import _geometry

    class M:
        output_shape = (None, 1, 16, 16, 16)
    rng = np.random.default_rng(1)
    for vshape in [(1, 30, 12, 20), (1, 16, 16, 16), (1, 7, 40, 15)]:
        vol = rng.standard_normal(vshape).astype(np.float32)
        g = _geometry(M, vol, (16, 16, 16), 0.5)
        d0, pf, _ = po.pad_volume(vol[0], (16, 16, 16), (16, 16, 16))
        assert g["padded"] == d0.shape and [tuple(p) for p in g["fit"]] == [tuple(p) for p in pf]
        # the virtual padding reproduces the reference's padded array exactly
        rebuilt = np.full(g["padded"], g["pad"][1], np.float64)
        sl = tuple(slice(a, a + s) for (a, _), s in zip(g["fit"], vol.shape[1:]))
        rebuilt[sl] = vol[0]
        assert np.array_equal(rebuilt.astype(np.float32), d0.astype(np.float32))


def test_reference_api_surface():
    import fetal_net.metrics as fm
    import fetal_net.model as fmod
    import fetal_net.prediction as fp
    import fetal_net.training as ft
    # names looked up by getattr in fetal/train_fetal.py:31-32 and exported by fetal_net/model/__init__.py:3-18
    for n in ["unet_model_3d", "isensee2017_model_3d", "unet_model_2d", "isensee2017_model", "fetal_envelope_model",
              "fetal_origin_model", "fetal_origin2_model", "fetal_origin3_model", "norm_net_model",
              "discriminator_image_2d", "discriminator_image_3d"]:
        assert callable(getattr(fmod, n)), n
    for n in ["dice_coefficient", "dice_coefficient_loss", "dice_coef", "dice_coef_loss", "vod_coefficient",
              "vod_coefficient_loss", "weighted_dice_coefficient", "weighted_dice_coefficient_loss",
              "binary_crossentropy_loss", "focal_loss", "dice_and_xent", "dice_and_xent_mask"]:
        assert callable(getattr(fm, n)), n
    # signatures (prediction.py:118-119, unet3d/unet.py:17-20, training.py:89-92)
    sig = inspect.signature(fp.patch_wise_prediction)
    assert list(sig.parameters)[:9] == ["model", "data", "patch_shape", "overlap_factor", "batch_size", "permute",
                                        "truth_data", "prev_truth_index", "prev_truth_size"]
    assert sig.parameters["overlap_factor"].default == 0 and sig.parameters["batch_size"].default == 5
    sig = inspect.signature(fmod.unet_model_3d)
    assert sig.parameters["depth"].default == 4 and sig.parameters["n_base_filters"].default == 32
    assert sig.parameters["initial_learning_rate"].default == 0.00001
    assert sig.parameters["loss_function"].default is fm.dice_coefficient_loss
    assert list(inspect.signature(ft.train_model).parameters)[:6] == [
        "model", "model_file", "training_generator", "validation_generator", "steps_per_epoch", "validation_steps"]
    sig = inspect.signature(fmod.isensee2017_model)               # unet/isensee.py:14-16
    assert sig.parameters["n_segmentation_levels"].default == 3 and sig.parameters["summation"].default is False
    assert sig.parameters["dropout_rate"].default == 0.3 and sig.parameters["initial_learning_rate"].default == 5e-4
    sig = inspect.signature(fmod.isensee2017_model_3d)            # unet3d/isensee2017.py:15-18
    assert sig.parameters["mask_shape"].default is None and sig.parameters["n_segmentation_levels"].default == 1


def test_host_metrics_known_answers():
    import fetal_net.metrics as fm
    d = np.zeros((1, 3, 10, 10, 10), np.float32)
    d[0, 0, :5] = 1
    d[0, 1, 5:] = 1
    d[0, 2, :, :5] = 1
    assert fm.dice_coefficient(d, d) == pytest.approx(1.0)
    assert fm.dice_coefficient(d, np.zeros_like(d)) == pytest.approx(1.0 / 1501.0)
    assert fm.dice_coefficient_loss(d, d) == pytest.approx(-1.0)
    # reference test/test_metrics.py:19-38 (weighted dice known answers)
    assert fm.weighted_dice_coefficient(d, d) == pytest.approx(1.0)
    e = d.copy()
    e[0, 0] = 0
    assert fm.weighted_dice_coefficient(d, e) == pytest.approx(2 / 3, abs=1e-5)
    assert fm.vod_coefficient(d, d) == pytest.approx(1.0)


def test_callbacks_order():
    # reference test/test_training.py:9-15
    from fetal_net.training import EarlyStopping, ReduceLROnPlateau, get_callbacks
    cbs = get_callbacks("model", early_stopping_patience=3)
    assert isinstance(cbs[2], ReduceLROnPlateau) and isinstance(cbs[3], EarlyStopping)


def test_permutation_helpers_match_golden_and_reference(golden):
    """predict_with_permutations / permute_data / reverse_permute_data (prediction.py:364-369, augment.py:380-469):
    pure host code around model.predict — checked against the golden frozen from the reference and, where the
    reference tree exists, against the live functions."""
    from tests.golden.make_golden import ramp_model
    from oracle.ref_harness import FunctionModel, reference_available, load_reference_prediction
    from fetal_net import prediction as P
    keys = P.generate_permutation_keys()
    assert len(keys) == 48
    data = golden["tta/perm/data"]
    for k in keys:
        assert np.array_equal(P.reverse_permute_data(P.permute_data(data, k), k), data)
    fn, oshape = ramp_model((8, 8, 6), 2)
    out = P.predict_with_permutations(FunctionModel(fn, oshape), data)
    ref = golden["tta/perm/out"]
    assert out.shape == ref.shape
    np.testing.assert_allclose(out, ref, rtol=2e-6, atol=2e-6)
    batch = np.stack([data, data[:, ::-1]])
    both = P.predict(FunctionModel(fn, oshape), batch, permute=True)
    assert both.shape == (2,) + ref.shape
    np.testing.assert_allclose(both[0], ref, rtol=2e-6, atol=2e-6)
    if reference_available():
        rp = load_reference_prediction()
        assert rp.generate_permutation_keys() == keys
        for k in list(keys)[:12]:
            assert np.array_equal(rp.permute_data(data, k), P.permute_data(data, k))
            assert np.array_equal(rp.reverse_permute_data(data, k), P.reverse_permute_data(data, k))
        np.testing.assert_allclose(rp.predict_with_permutations(FunctionModel(fn, oshape), data), out,
                                   rtol=2e-6, atol=2e-6)


def test_rescale_intensity_restatement():
    # skimage.exposure.rescale_intensity(in_range=(lo,hi), out_range='image') semantics used by augment.py:123-126
    from fetal_net.prediction import rescale_intensity_to_image_range
    d = np.array([-2.0, -1.0, 0.0, 1.0, 4.0], np.float32)
    out = rescale_intensity_to_image_range(d, -1.0, 1.0)
    assert out.dtype == np.float32
    np.testing.assert_allclose(out, [-2.0, -2.0, 1.0, 4.0, 4.0])


def test_host_patch_slicing_matches_reference_goldens(golden):
    """fetal_net.utils.patches.get_patch_from_3d_data incl. the out-of-bounds edge completion (patches.py:57-91)."""
    from fetal_net.utils.patches import fix_out_of_bound_patch_attempt, get_patch_from_3d_data
    data = golden["patch/data"]
    shape = tuple(int(v) for v in golden["patch/shape"])
    for i, c in enumerate(golden["patch/corners"]):
        got = get_patch_from_3d_data(data, shape, c)
        assert np.array_equal(got, golden["patch/out%d" % i]), (i, c)
        padded, fixed = fix_out_of_bound_patch_attempt(data, np.asarray(shape), np.asarray(c))
        sl = tuple(slice(int(f), int(f) + s) for f, s in zip(fixed, shape))
        assert np.array_equal(padded[(Ellipsis,) + sl], golden["patch/out%d" % i])


def test_keras_h5_bridge_host_parts(tmp_path):
    """fetal_net.keras_h5: file sniffing, creation-order renumbering, and the guarded h5py path
    (h5py is absent here: the error must say how to convert the checkpoint instead)."""
    from fetal_net import keras_h5
    npz = tmp_path / "w.h5"                              # the ModelCheckpoint pattern keeps the .h5 name
    with open(npz, "wb") as f:
        np.savez(f, a=np.zeros(3))
    assert not keras_h5.is_hdf5(str(npz))
    fake = tmp_path / "keras.h5"
    fake.write_bytes(keras_h5.HDF5_MAGIC + b"\0" * 64)
    assert keras_h5.is_hdf5(str(fake))
    entries = [("conv", np.zeros((3, 3, 3, 1, 16), np.float32), np.zeros(16, np.float32)),
               ("norm", np.ones(16, np.float32), np.zeros(16, np.float32)),
               ("conv", np.zeros((3, 3, 3, 16, 16), np.float32), np.zeros(16, np.float32)),
               ("conv", np.zeros((3, 3, 6, 32), np.float32), np.zeros(32, np.float32))]
    arrays = keras_h5.to_npz_arrays(entries)
    assert sorted(arrays) == ["conv2d_3/bias:0", "conv2d_3/kernel:0", "conv3d_1/bias:0", "conv3d_1/kernel:0",
                              "conv3d_2/bias:0", "conv3d_2/kernel:0", "instance_normalization_1/beta:0",
                              "instance_normalization_1/gamma:0"]
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="export_keras_weights"):
            keras_h5.read_keras_h5_weights(str(fake))


def test_keras_h5_creation_order_for_isensee_heads():
    """Keras writes `layer_names` in model.layers order (sorted by graph depth): for isensee2017 with 2 segmentation
    levels the level-1 head Conv3D(n_labels, 1) (created right after u1_loc2) is listed after the later-created level-0
    decoder convs. The bridge must restore creation order from the auto-name suffixes so kernels line up with our table."""
    from fetal_net import keras_h5
    from oracle import unet_oracle as uo
    layers = uo.isensee3d_layers(3, 16, 2)                   # creation order: (..., u1_loc2, u1_seg, u0_up, ...)
    names = [n for n, *_ in layers]
    assert names.index("u1_seg") < names.index("u0_up")
    entries = []
    for i, (name, cin, cout, k) in enumerate(layers, start=1):
        entries.append(("conv", "conv3d_%d" % i, np.full((k, k, k, cin, cout), float(i), np.float32),
                        np.zeros(cout, np.float32)))
    # depth-sorted file order: the coarse head moves behind the level-0 decoder convs; 'conv3d_10' < 'conv3d_9' lexically
    seg = entries.pop(names.index("u1_seg"))
    entries.insert(len(entries) - 1, seg)
    shuffled = [e[1] for e in entries]
    assert shuffled != ["conv3d_%d" % i for i in range(1, len(layers) + 1)]
    arrays = keras_h5.to_npz_arrays(keras_h5.creation_order(entries))
    for i, (name, cin, cout, k) in enumerate(layers, start=1):
        kern = arrays["conv3d_%d/kernel:0" % i]
        assert kern.shape == (k, k, k, cin, cout) and float(kern.flat[0]) == float(i), (i, name)


def test_checkpoint_npz_roundtrip_logic(tmp_path):
    """Model.save_weights / load_weights without a device: the .npz is keyed by Keras layer names, kernel layout
    untouched, file name kept as given (ModelCheckpoint writes '...-epochNN-....h5')."""
    from fetal_net.model.unet3d import Model

    class Stub:
        layers = [dict(keras_name="conv3d_1", is_norm=False), dict(keras_name="instance_normalization_1", is_norm=True),
                  dict(keras_name="conv3d_2", is_norm=False)]
        input_shape, depth, n_base_filters, n_labels = (None, 1, 8, 8, 8), 2, 16, 1
        _weights_from_mapping = Model._weights_from_mapping
        _weight_arrays = Model._weight_arrays
        load_optimizer_state = Model.load_optimizer_state

        class optimizer:
            lr = 3e-4

        def __init__(self):
            rng = np.random.default_rng(0)
            shapes = [(3, 3, 3, 1, 16), (16,), (16,), (16,), (3, 3, 3, 16, 1), (1,)]
            self.w = [rng.standard_normal(s).astype(np.float32) for s in shapes]
            self.moments = [tuple(rng.standard_normal(s).astype(np.float32) for s in (k, b, k, b))
                            for k, b in zip(shapes[0::2], shapes[1::2])]
            self.restored = None

        def get_weights(self):
            return self.w

        def set_weights(self, ws):
            self.loaded = ws

        def get_optimizer_state(self):
            return 1234, 2.5e-5, self.moments

        def set_optimizer_state(self, iterations, lr, moments):
            self.restored = (iterations, lr, moments)

    a, b = Stub(), Stub()
    path = str(tmp_path / "fetal_net_model-epoch01-loss-0.500-acc0.900.h5")
    Model.save_weights(a, path)
    with np.load(path) as z:
        assert "conv3d_1/kernel:0" in z.files and "instance_normalization_1/gamma:0" in z.files
        assert "__iterations__" not in z.files                       # save_weights stays weights-only
    Model.load_weights(b, path)
    assert len(b.loaded) == 6 and all(np.array_equal(p, q) for p, q in zip(b.loaded, a.w))
    assert Model.load_optimizer_state(b, path) is False and b.restored is None
    # Keras' model.save(): weights + Adam moments + iteration count + the current (plateau-reduced) learning rate,
    # which load_old_model restores so a resumed run continues like the reference's load_model path
    full = str(tmp_path / "fetal_net_model-epoch02-loss-0.600-acc0.910.h5")
    Model.save(a, full)
    c = Stub()
    Model.load_weights(c, full)
    assert all(np.array_equal(p, q) for p, q in zip(c.loaded, a.w))
    assert Model.load_optimizer_state(c, full) is True
    it, lr, moments = c.restored
    assert it == 1234 and lr == 2.5e-5 and len(moments) == 3
    for got, ref in zip(moments, a.moments):
        assert all(np.array_equal(g, r) for g, r in zip(got, ref))


def test_train_model_driver_and_callbacks_on_a_stub(tmp_path):
    """train_model -> Model.fit_generator -> callbacks (training.py:26-42,89-124) without a device: checkpoint file
    pattern and save-best-only, CSV log, ReduceLROnPlateau, EarlyStopping, and load_old_model's builder dispatch."""
    import itertools
    from fetal_net import training
    from fetal_net.model.unet3d import Model

    class Opt:
        lr = 1e-3

    val_losses = [-0.50, -0.60, -0.55, -0.58, -0.57, -0.56, -0.40, -0.40]

    class Stub:
        metrics_names = ['loss', 'binary_accuracy', 'vod_coefficient']
        fit_generator = Model.fit_generator
        save = Model.save_weights
        _weight_arrays = Model._weight_arrays
        _weights_from_mapping = Model._weights_from_mapping
        layers = [dict(keras_name="conv3d_1", is_norm=False)]
        input_shape, depth, n_base_filters, n_labels, name, isensee_levels = (None, 1, 8, 8, 8), 2, 16, 1, 'unet_model_3d', None

        def __init__(self):
            self.optimizer, self.stop_training, self.steps, self.epoch = Opt(), False, 0, 0

        def get_weights(self):
            return [np.full((3, 3, 3, 1, 16), self.steps, np.float32), np.zeros(16, np.float32)]

        def train_on_batch(self, x, y):
            self.steps += 1
            return [-0.3, 0.9, 0.2]

        def test_on_batch(self, x, y):
            return [val_losses[min(self.epoch, len(val_losses) - 1)], 0.95, 0.3]

    model = Stub()

    def gen():
        for i in itertools.count():
            yield np.zeros((2, 1, 8, 8, 8), np.float32), np.zeros((2, 1, 8, 8, 8), np.float32)

    class EpochCounter(training.Callback):
        def on_epoch_begin(self, epoch, logs=None):
            model.epoch = epoch

    orig = training.get_callbacks
    training.get_callbacks = lambda *a, **k: [EpochCounter()] + orig(*a, **dict(k, verbosity=0))
    try:
        hist = training.train_model(model, str(tmp_path / "fetal_net_model"), gen(), gen(), steps_per_epoch=3,
                                    validation_steps=2, initial_learning_rate=1e-3, learning_rate_drop=0.5,
                                    learning_rate_patience=2, early_stopping_patience=4, n_epochs=20,
                                    output_folder=str(tmp_path))
    finally:
        training.get_callbacks = orig
    # best val_loss -0.60 at epoch 2; EarlyStopping(patience 4) stops after epoch 6
    assert len(hist["val_loss"]) == 6 and model.steps == 18
    files = sorted(os.path.basename(f) for f in os.listdir(tmp_path) if f.endswith(".h5"))
    assert files == ["fetal_net_model-epoch01-loss-0.500-acc0.950.h5", "fetal_net_model-epoch02-loss-0.600-acc0.950.h5"]
    assert training.get_last_model_path(str(tmp_path / "fetal_net_model")).endswith("epoch02-loss-0.600-acc0.950.h5")
    assert model.optimizer.lr == pytest.approx(2.5e-4)              # halved after epochs 4 and 6 (patience 2)
    rows = open(tmp_path / "training").read().strip().splitlines()
    assert rows[0].startswith("epoch,") and len(rows) == 7
    with np.load(tmp_path / files[-1]) as z:
        assert str(z["__builder__"]) == "unet_model_3d" and [int(v) for v in z["__config__"]] == [1, 8, 8, 8, 2, 16, 1]


def test_small_host_helpers_on_a_stub():
    """step_decay / LearningRateScheduler (training.py:22-23,34-37), Model.evaluate (Keras batch-size weighting),
    Model.summary / to_json - host logic only, no device."""
    import json
    from fetal_net import training
    from fetal_net.model.unet3d import Model
    assert training.step_decay(0, 1e-3, 0.5, 10) == pytest.approx(1e-3)
    assert training.step_decay(9, 1e-3, 0.5, 10) == pytest.approx(5e-4)          # floor((1 + 9) / 10) = 1
    assert training.step_decay(29, 1e-3, 0.5, 10) == pytest.approx(1.25e-4)

    class Opt:
        lr = 1.0

    class Stub:
        metrics_names = ['loss', 'binary_accuracy', 'vod_coefficient']
        optimizer = Opt()
        ndim, name, depth, n_base_filters, n_labels = 3, 'unet_model_3d', 2, 16, 1
        input_shape = (None, 1, 8, 8, 8)
        layers = [dict(name="enc0a", keras_name="conv3d_1", cin=1, cout=16, k=3, is_norm=False),
                  dict(name="l0_in_norm", keras_name="instance_normalization_1", cin=16, cout=16, k=0, is_norm=True)]

        def test_on_batch(self, x, y):
            return [float(len(x)), 1.0, 0.5]

        def count_params(self):
            return 27 * 16 + 16 + 32

    sched = training.LearningRateScheduler(lambda e: 0.1 * (e + 1))
    sched.set_model(Stub)
    sched.on_epoch_begin(2)
    assert Stub.optimizer.lr == pytest.approx(0.3)
    x = np.zeros((5, 1, 2, 2, 2), np.float32)
    ev = Model.evaluate(Stub(), x, x, batch_size=2)                  # batches of 2, 2, 1 -> weighted mean of 2, 2, 1
    assert ev[0] == pytest.approx((2 * 2 + 2 * 2 + 1 * 1) / 5) and ev[1] == pytest.approx(1.0)
    lines = []
    Model.summary(Stub(), print_fn=lines.append)
    assert lines[-1] == "Total params: %d" % (27 * 16 + 16 + 32) and "conv3d_1" in lines[1] and lines[2].split()[-1] == "32"
    cfg = json.loads(Model.to_json(Stub()))
    assert cfg["class_name"] == "unet_model_3d" and cfg["input_shape"] == [1, 8, 8, 8]


def test_dice_and_xent_host_helpers_and_device_loss_spec():
    """fetal_net.metrics.dice_and_xent / dice_and_xent_mask (reference metrics.py:68-95): host evaluation against torch's
    binary_cross_entropy, and the mapping of loss callables to fm_model_set_loss arguments."""
    import functools
    import torch
    import torch.nn.functional as F
    import fetal_net.metrics as fm
    from oracle import unet_oracle as uo
    rng = np.random.default_rng(0)
    p = rng.random((2, 1, 8, 8, 8)).astype(np.float32)
    p.ravel()[:3] = [0.0, 1.0, 1e-9]                                 # the Keras clip at 1e-7 is active
    t = (rng.random(p.shape) < 0.4).astype(np.float32)
    mask = (5 * rng.random(p.shape)).astype(np.float32)
    pc = torch.as_tensor(p, dtype=torch.float64).clamp(float(np.float32(1e-7)), float(np.float32(1) - np.float32(1e-7)))
    bce = F.binary_cross_entropy(pc, torch.as_tensor(t, dtype=torch.float64), reduction="none")
    want = fm.dice_coefficient_loss(t, p) + 0.5 * float(bce.mean())
    assert fm.dice_and_xent(t, p, xent_weight=0.5) == pytest.approx(want, rel=1e-12)
    wantm = fm.dice_coefficient_loss(t, p) + float((torch.exp(-torch.as_tensor(mask, dtype=torch.float64) / 3) * bce).mean())
    assert fm.dice_and_xent_mask(mask)(t, p) == pytest.approx(wantm, rel=1e-12)
    # the oracle's torch version agrees with the host helper
    tt, pp = torch.as_tensor(t, dtype=torch.float64), torch.as_tensor(p, dtype=torch.float64)
    assert float(uo.dice_and_xent(tt, pp, 1.0, torch.as_tensor(mask, dtype=torch.float64), 3.0)) == \
        pytest.approx(wantm, rel=1e-12)
    spec = fm.device_loss_spec
    assert spec(fm.dice_coefficient_loss) == (0, 0.0, 0.0)
    assert spec(fm.dice_and_xent) == (1, 1.0, 0.0)
    assert spec(functools.partial(fm.dice_and_xent, xent_weight=0.3)) == (1, 0.3, 0.0)
    assert spec(fm.dice_and_xent_mask) is None                       # needs the mask input (mask_shape)
    assert spec(fm.dice_and_xent_mask, has_mask_input=True) == (2, 1.0, 3.0)
    assert spec(fm.dice_and_xent_mask(None, xent_weight=2.0, dist_sigma=5), has_mask_input=True) == (2, 2.0, 5.0)
    assert spec(fm.vod_coefficient_loss) is None and spec(fm.focal_loss) is None


def test_keras_h5_bridge_orders_transposed_convs_and_batch_norm_statistics():
    """fetal_net.keras_h5: Deconvolution3D layers ('conv3d_transpose_<n>') and BatchNormalization layers (gamma, beta,
    moving_mean, moving_variance) of a deconvolution=True / batch_normalization=True reference checkpoint are numbered on
    their own, in creation order, under the names Model.load_weights looks up."""
    from fetal_net import keras_h5
    z = lambda *s: np.zeros(s, np.float32)
    # graph-depth order as a Keras file would list it: suffixes are the creation order
    entries = [("conv", "conv3d_12", z(3, 3, 3, 8, 8), z(8)), ("deconv", "conv3d_transpose_4", z(2, 2, 2, 8, 8) + 4, z(8)),
               ("norm", "batch_normalization_9", z(8) + 9, z(8)), ("moving", "batch_normalization_9", z(8) + 90, z(8)),
               ("conv", "conv3d_11", z(3, 3, 3, 1, 8), z(8)), ("deconv", "conv3d_transpose_3", z(2, 2, 2, 8, 8) + 3, z(8)),
               ("norm", "batch_normalization_8", z(8) + 8, z(8)), ("moving", "batch_normalization_8", z(8) + 80, z(8))]
    arrays = keras_h5.to_npz_arrays(keras_h5.creation_order(entries))
    assert arrays["conv3d_1/kernel:0"].shape == (3, 3, 3, 1, 8) and arrays["conv3d_2/kernel:0"].shape == (3, 3, 3, 8, 8)
    assert arrays["conv3d_transpose_1/kernel:0"].flat[0] == 3 and arrays["conv3d_transpose_2/kernel:0"].flat[0] == 4
    assert arrays["batch_normalization_1/gamma:0"][0] == 8 and arrays["batch_normalization_2/gamma:0"][0] == 9
    assert arrays["batch_normalization_1/moving_mean:0"][0] == 80
    assert arrays["batch_normalization_2/moving_mean:0"][0] == 90
    assert "batch_normalization_2/moving_variance:0" in arrays
    # instance-norm checkpoints keep their names
    arrays = keras_h5.to_npz_arrays([("norm", "instance_normalization_5", z(4) + 1, z(4))])
    assert list(arrays) == ["instance_normalization_1/gamma:0", "instance_normalization_1/beta:0"]
