"""Per-launch CUDA-event trace of one training step (batch 8 x 64^3): name, ms, TFLOP/s or GB/s, in launch order.
    python tools/step_trace.py [filter]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fetal-mri-segmentation_b200"))
from fetal_net import _lib
from fetal_net.model import unet_model_3d

flt = sys.argv[1] if len(sys.argv) > 1 else ""
ctx = _lib.get_context(0)
m = unet_model_3d(input_shape=(1, 64, 64, 64), n_base_filters=16, depth=4, initial_learning_rate=1e-4)
m.init_glorot_uniform(seed=0)
x = np.random.default_rng(1).standard_normal((8, 1, 64, 64, 64)).astype(np.float32)
t = (np.random.default_rng(2).random(x.shape) < 0.3).astype(np.float32)
for _ in range(3):
    m.train_on_batch(x, t)
best = None
for rep in range(3):
    ctx.profile(True)
    m.train_on_batch(x, t)
    recs = ctx.profile_records()
    ctx.profile(False)
    if best is None:
        best = [list(r) for r in recs]
    else:
        for b, r in zip(best, recs):
            b[1] = min(b[1], r[1])
tot = 0.0
for i, (name, ms, fl, by) in enumerate(best):
    tot += ms
    if flt in name:
        print("%3d %-22s %7.3f ms  %7.1f TF/s  %7.1f GB/s" % (i, name, ms, fl / ms / 1e9 if fl else 0.0, by / ms / 1e6 if by else 0.0))
print("sum %.3f ms over %d launches" % (tot, len(best)))
