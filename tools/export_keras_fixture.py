"""Pins the network oracle against REAL Keras output. Run this once where the reference's Keras/TensorFlow
environment lives, with the reference checkout on PYTHONPATH (no GPU needed):

    PYTHONPATH=/path/to/Fetal-MRI-Segmentation python tools/export_keras_fixture.py tests/golden/keras_fixture.npz

For each of the three builders on the hot path it builds the UNMODIFIED reference model at a small shape, draws
an input, and stores: every weight array keyed by its Keras variable name (re-sorted into layer creation order by
`fetal_net.keras_h5.creation_order`, the order `Model.set_weights` of this repository takes), the input, `model.predict(x)`, and the loss /
metrics of ONE `train_on_batch(x, t)` followed by the updated weights (soft Dice + Keras-Adam step).
`tests/test_oracle_pinning.py::test_oracle_matches_keras_fixture` (CPU) and
`tests/test_gpu_model.py::test_gpu_matches_keras_fixture` (GPU) pick the file up when it exists and compare
oracle / CUDA path against it with the tolerances stated there; until then they skip and the network oracle stays
"parity unpinned" (DESIGN.md §2)."""
import sys

import numpy as np


def record(out, tag, model, x, t):
    # keyed by the Keras variable name ('conv3d_7/kernel:0'): `model.layers` is sorted by graph depth, not by
    # creation, so the reader re-sorts by the numeric suffix (fetal_net.keras_h5.creation_order)
    from keras import backend as K
    names = [w.name for w in model.weights]
    out[tag + "/names"] = np.array(names)
    for n, w in zip(names, K.batch_get_value(model.weights)):
        out["%s/w/%s" % (tag, n)] = np.asarray(w, np.float32)
    out[tag + "/x"] = x
    out[tag + "/t"] = t
    out[tag + "/predict"] = np.asarray(model.predict(x), np.float32)
    res = model.train_on_batch(x, t)
    out[tag + "/train_metrics"] = np.asarray(res, np.float64)            # [loss, binary_accuracy, vod_coefficient]
    for n, w in zip(names, K.batch_get_value(model.weights)):
        out["%s/w_after/%s" % (tag, n)] = np.asarray(w, np.float32)
    out[tag + "/lr"] = np.float64(float(K.get_value(model.optimizer.lr)))


def main():
    if len(sys.argv) != 2:
        sys.exit(__doc__)
    from keras import backend as K
    K.set_image_dim_ordering('th')                                      # as fetal/train_fetal.py does
    from fetal_net.model import unet_model_3d, unet_model_2d, isensee2017_model_3d
    rng = np.random.RandomState(0)
    out = {}
    x = rng.standard_normal((2, 1, 32, 32, 32)).astype(np.float32)
    t = (rng.random_sample(x.shape) < 0.3).astype(np.float32)
    record(out, "unet3d_d4_nf16", unet_model_3d(input_shape=(1, 32, 32, 32), depth=4, n_base_filters=16,
                                               initial_learning_rate=1e-4), x, t)
    x = rng.standard_normal((2, 1, 32, 32, 16)).astype(np.float32)
    t = (rng.random_sample(x.shape) < 0.3).astype(np.float32)
    record(out, "isensee3d_d3_nf8_seg2", isensee2017_model_3d(input_shape=(1, 32, 32, 16), depth=3, n_base_filters=8,
                                                              n_segmentation_levels=2, dropout_rate=0.0,
                                                              initial_learning_rate=5e-4), x, t)
    K.set_image_dim_ordering('tf')                                      # the 2D builder is channels-last
    x = rng.standard_normal((2, 32, 32, 6)).astype(np.float32)
    x[..., 5] = rng.random_sample((2, 32, 32)) < 0.3
    t = (rng.random_sample((2, 32, 32, 1)) < 0.3).astype(np.float32)
    record(out, "unet2d_d3_nf16", unet_model_2d(input_shape=(32, 32, 6), depth=3, n_base_filters=16,
                                               initial_learning_rate=1e-4), x, t)
    np.savez_compressed(sys.argv[1], **out)
    print("wrote", sys.argv[1], "(%d arrays)" % len(out))


if __name__ == "__main__":
    main()
