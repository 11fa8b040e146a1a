"""TEST INFRASTRUCTURE ONLY. Reader for `tests/golden/keras_fixture.npz`, the file `tools/export_keras_fixture.py`
writes where the reference's Keras/TensorFlow environment lives (real Keras weights, inputs, `predict` outputs and one
`train_on_batch` step of the three hot-path builders). The file does not exist in this image (no Keras): the tests that
consume it skip, and the network oracle stays PARITY UNPINNED until somebody drops it in."""
import os
import re

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "keras_fixture.npz")


def _suffix(var_name):
    m = re.search(r"_(\d+)/", var_name)
    return int(m.group(1)) if m else 0


def named_weights(z, tag, layers, which="w"):
    """{'<our layer>/kernel|bias|gamma|beta': array} from the fixture's Keras-named variables. Keras numbers its
    auto-named layers in creation order (conv3d_7 < conv3d_12), which is the order of the oracle's layer tables
    (`unet3d_layers`, `isensee3d_layers`); `model.weights` itself is sorted by graph depth, so sort by the suffix."""
    names = [str(n) for n in z[tag + "/names"]]
    convs = sorted({n.split("/")[0] for n in names if n.endswith("/kernel:0")}, key=lambda n: _suffix(n + "/"))
    norms = sorted({n.split("/")[0] for n in names if n.endswith("/gamma:0")}, key=lambda n: _suffix(n + "/"))
    assert len(convs) == len(layers), (len(convs), len(layers))
    out, ni = {}, 0
    for (lname, cin, cout, k), kname in zip(layers, convs):
        kern = np.asarray(z["%s/%s/%s/kernel:0" % (tag, which, kname)])
        assert kern.shape[-2:] == (cin, cout), (lname, kname, kern.shape)
        out[lname + "/kernel"] = kern
        out[lname + "/bias"] = np.asarray(z["%s/%s/%s/bias:0" % (tag, which, kname)])
        if norms and not lname.endswith("_seg"):
            out[lname + "/gamma"] = np.asarray(z["%s/%s/%s/gamma:0" % (tag, which, norms[ni])])
            out[lname + "/beta"] = np.asarray(z["%s/%s/%s/beta:0" % (tag, which, norms[ni])])
            ni += 1
    assert ni == len(norms)
    return out


def load():
    return np.load(PATH) if os.path.exists(PATH) else None
