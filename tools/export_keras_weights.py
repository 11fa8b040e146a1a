"""Run this where the reference's Keras environment lives (it only needs numpy + h5py, no GPU, no TensorFlow):

    python tools/export_keras_weights.py fetal_net_model-epoch37-loss-0.912-acc0.991.h5 weights.npz

Writes the checkpoint's convolution / normalisation weights, in Keras creation order, as the .npz that
`fetal_net.model.Model.load_weights` of this repository reads (keys conv3d_<n>/kernel:0, conv3d_<n>/bias:0,
instance_normalization_<n>/gamma:0, .../beta:0, renumbered from 1). Kernels stay in Keras layout."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fetal-mri-segmentation_b200",
                                "fetal_net"))
import keras_h5  # noqa: E402  (plain module import: works without the CUDA library)


def main():
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    entries = keras_h5.read_keras_h5_weights(sys.argv[1])
    arrays = keras_h5.to_npz_arrays(entries)
    with open(sys.argv[2], "wb") as f:
        np.savez(f, **arrays)
    print("wrote %s: %d convolutions, %d normalisation layers" %
          (sys.argv[2], sum(k == "conv" for k, *_ in entries), sum(k == "norm" for k, *_ in entries)))


if __name__ == "__main__":
    main()
