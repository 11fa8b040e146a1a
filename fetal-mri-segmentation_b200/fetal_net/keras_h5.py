"""Optional bridge to the reference's Keras checkpoints (fetal_net/training.py:30-32 ModelCheckpoint `.h5` files,
fetal/utils.py:42-43 get_last_model_path) — SURVEY.md §8f rank 2.

The build image has neither HDF5 nor Keras, so this module is import-guarded and NOT exercised by the test suite:
it needs `h5py` (always present where Keras is). Two ways to move weights:

  * in an environment with h5py:   Model.load_weights("fetal_net_model-epoch37-....h5") reads the file directly;
  * anywhere else:                 run `python tools/export_keras_weights.py model.h5 weights.npz` next to the Keras
                                   install once, then Model.load_weights("weights.npz") here.

Layout read (Keras 2.x): a weights file has the layer groups at the root, a full `model.save()` file under
`model_weights/`; the group attribute `layer_names` gives the layer order, each layer group's `weight_names` its
arrays (`<layer>/kernel:0`, `<layer>/bias:0`; `gamma:0` / `beta:0` for keras_contrib InstanceNormalization).
Keras numbers auto-named layers per session (`conv3d_7` ...), so layers are matched to ours BY ORDER, not by name:
convolutions (kernel + bias) and normalisations (gamma + beta) in the order Keras CREATED them, which is the order
of this package's layer table. The file's `layer_names` attribute is `model.layers`, which Keras sorts by graph depth,
not by creation (for isensee2017 with n_segmentation_levels >= 2 the coarse Conv3D(n_labels, 1) heads come after
later-created decoder convs) - so each kind is re-sorted by the numeric suffix of its auto-name (conv3d_7 < conv3d_12),
which is the true creation order."""
import re
import numpy as np

HDF5_MAGIC = b"\x89HDF\r\n\x1a\n"


def is_hdf5(path):
    with open(path, "rb") as f:
        return f.read(8) == HDF5_MAGIC


def _as_str(v):
    return v.decode("utf8") if isinstance(v, bytes) else str(v)


def _auto_index(layer_name):
    """'conv3d_12' -> 12, 'conv3d' (TF-Keras names the first one without a suffix) -> 0."""
    m = re.search(r"_(\d+)$", layer_name)
    return int(m.group(1)) if m else 0


def creation_order(entries):
    """Re-sorts (kind, layer_name, first, second) entries so that, within each kind, layers follow the numeric
    suffix of their Keras auto-name = the order the builder created them in. The relative interleaving of kinds is
    irrelevant (to_npz_arrays numbers convolutions and normalisations separately)."""
    convs = sorted((e for e in entries if e[0] == "conv"), key=lambda e: _auto_index(e[1]))
    deconvs = sorted((e for e in entries if e[0] == "deconv"), key=lambda e: _auto_index(e[1]))
    norms = sorted((e for e in entries if e[0] == "norm"), key=lambda e: _auto_index(e[1]))
    moving = sorted((e for e in entries if e[0] == "moving"), key=lambda e: _auto_index(e[1]))
    return convs + deconvs + norms + moving


def read_keras_h5_weights(path):
    """-> list of (kind, layer_name, first, second) in Keras CREATION order; kind 'conv' (kernel, bias) or 'norm'
    (gamma, beta). Kernels are returned in Keras layout (k, k[, k], Cin, Cout), exactly what fm_model_set_weights takes."""
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - h5py is absent in the build image
        raise ImportError("reading a Keras .h5 checkpoint needs h5py; convert it once with "
                          "tools/export_keras_weights.py where Keras is installed and load the .npz instead") from e
    out = []
    with h5py.File(path, "r") as f:
        g = f["model_weights"] if "model_weights" in f else f
        for lname in [_as_str(n) for n in g.attrs["layer_names"]]:
            names = [_as_str(n) for n in g[lname].attrs.get("weight_names", [])]
            arrays = {n.split("/")[-1].split(":")[0]: np.asarray(g[lname][n]) for n in names}
            if "kernel" in arrays:
                bias = arrays.get("bias", np.zeros(arrays["kernel"].shape[-1], np.float32))
                # Deconvolution3D/2D layers ("conv3d_transpose_<n>") are numbered on their own by Keras
                out.append(("deconv" if "transpose" in lname else "conv", lname, arrays["kernel"].astype(np.float32),
                            bias.astype(np.float32)))
            elif "gamma" in arrays and "beta" in arrays:
                out.append(("norm", lname, arrays["gamma"].astype(np.float32), arrays["beta"].astype(np.float32)))
                if "moving_mean" in arrays:       # BatchNormalization: the non-trainable moving statistics
                    out.append(("moving", lname, arrays["moving_mean"].astype(np.float32),
                                arrays["moving_variance"].astype(np.float32)))
    return creation_order(out)


def to_npz_arrays(entries):
    """The same weights keyed the way Model.save_weights / load_weights key their .npz (layers renumbered from 1 in
    creation order: conv3d_1/kernel:0, conv3d_1/bias:0, instance_normalization_1/gamma:0, ...)."""
    arrays, n_conv, n_norm, n_deconv, n_moving = {}, 0, 0, 0, 0
    for e in entries:
        kind, a, b = e[0], e[-2], e[-1]                     # (kind, a, b) or (kind, layer_name, a, b)
        stem = re.sub(r"_\d+$", "", e[1]) if len(e) == 4 and kind in ("norm", "moving") else "instance_normalization"
        if kind == "moving":
            n_moving += 1
            prefix = "%s_%d" % (stem, n_moving)
            arrays[prefix + "/moving_mean:0"], arrays[prefix + "/moving_variance:0"] = a, b
            continue
        if kind == "deconv":
            n_deconv += 1
            prefix = "conv%dd_transpose_%d" % (a.ndim - 2, n_deconv)
            arrays[prefix + "/kernel:0"], arrays[prefix + "/bias:0"] = a, b
        elif kind == "conv":
            n_conv += 1
            prefix = "conv%dd_%d" % (a.ndim - 2, n_conv)
            arrays[prefix + "/kernel:0"], arrays[prefix + "/bias:0"] = a, b
        else:
            n_norm += 1
            prefix = "%s_%d" % (stem, n_norm)
            arrays[prefix + "/gamma:0"], arrays[prefix + "/beta:0"] = a, b
    return arrays
