"""Noise floor of the pipelining test: runs the 4-step training sequence of
tests/test_gpu_model.py::_PIPELINE_SCRIPT several times per route (pageable / pinned) and prints the losses, so a race
in the pipelined route can be told from the run-to-run spread of the fp32 red.add gradients."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fetal-mri-segmentation_b200"))
from fetal_net.model import unet_model_3d
from oracle import unet_oracle as uo
from tests.test_gpu_model import decisive_weights, blob_target
w = decisive_weights(uo.unet3d_layers(4, 16))
rng = np.random.default_rng(21)
xs = [rng.standard_normal((2, 1, 32, 32, 32)).astype(np.float32) for _ in range(4)]
ts = [blob_target(x.shape, rng) for x in xs]
for rep in range(3):
    for pinned in (False, True):
        model = unet_model_3d(input_shape=(1, 32, 32, 32), n_base_filters=16, depth=4, initial_learning_rate=1e-4)
        model.set_named_weights(w)
        ls = []
        for x, t in zip(xs, ts):
            if pinned:
                xp, tp = torch.as_tensor(x).pin_memory(), torch.as_tensor(t).pin_memory()
                ls.append(model.train_on_batch(xp.numpy(), tp.numpy())[0])
                xp.zero_(); tp.zero_()
            else:
                ls.append(model.train_on_batch(x, t)[0])
        print("pinned" if pinned else "pageable", ["%.7f" % l for l in ls], flush=True)
