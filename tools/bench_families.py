"""Throughput of the other two model families on the hot path (BASELINE configs[2] and [3]); one JSON line each.
    python tools/bench_families.py [steps]
Isensee-2017 (depth 5, nf 16, 3 segmentation levels) on 128x128x64 patches: train step (batch 2, dropout 0.3) and
predict; 2.5D U-Net (depth 4, nf 32) on 256x256 stacks of 5 slices + 1 previous-truth channel: train step (batch 8)
and predict. Device time by CUDA events on the library's stream, host buffers (the public Model API)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fetal-mri-segmentation_b200"))
from fetal_net import _lib  # noqa: E402
from fetal_net.model import isensee2017_model_3d, unet_model_2d  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ctx = _lib.get_context(0)


def timed(fn, n):
    for _ in range(3):
        fn()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    ctx.synchronize()
    return (time.perf_counter() - t0) / n


def kernels(fn):
    ctx.profile(True)
    fn()
    agg = {}
    for name, ms, fl, by in ctx.profile_records():
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ms
        a[2] += fl
    ctx.profile(False)
    tot = sum(a[1] for a in agg.values())
    top = sorted(agg.items(), key=lambda kv: -kv[1][1])[:6]
    return tot, {k: dict(launches=a[0], ms=round(a[1], 3), tflops=round(a[2] / a[1] / 1e9, 1) if a[2] else None)
                 for k, a in top}


rng = np.random.default_rng(0)
# ---- configs[2]: Isensee
shape = (1, 128, 128, 64)
m = isensee2017_model_3d(input_shape=shape, n_base_filters=16, depth=5, n_segmentation_levels=3, dropout_rate=0.3,
                         initial_learning_rate=5e-4)
m.init_glorot_uniform(seed=0)
B = 2
x = rng.standard_normal((B,) + shape).astype(np.float32)
t = (rng.random(x.shape) < 0.3).astype(np.float32)
vox = B * int(np.prod(shape[1:]))
s_train = timed(lambda: m.train_on_batch(x, t), steps)
s_pred = timed(lambda: m.predict(x), steps)
kt, top_t = kernels(lambda: m.train_on_batch(x, t))
kp, top_p = kernels(lambda: m.predict(x))
print(json.dumps(dict(workload="isensee2017_model_3d depth 5 nf 16 nseg 3, %d x 128x128x64 (configs[2])" % B,
                      train_ms=s_train * 1e3, train_voxels_per_s=vox / s_train, train_kernel_ms=kt,
                      predict_ms=s_pred * 1e3, predict_voxels_per_s=vox / s_pred, predict_kernel_ms=kp,
                      fwd_gflop_per_patch=173.638, predict_conv_tflops=B * 173.638 / kp,
                      top_train=top_t, top_predict=top_p)))
del m
# ---- configs[3]: 2.5D U-Net
m2 = unet_model_2d(input_shape=(256, 256, 6), n_base_filters=32, depth=4, initial_learning_rate=1e-4)
m2.init_glorot_uniform(seed=0)
B = 8
x2 = rng.standard_normal((B, 256, 256, 6)).astype(np.float32)
t2 = (rng.random((B, 256, 256, 1)) < 0.3).astype(np.float32)
pix = B * 256 * 256
s_train = timed(lambda: m2.train_on_batch(x2, t2), steps)
s_pred = timed(lambda: m2.predict(x2), steps)
kt, top_t = kernels(lambda: m2.train_on_batch(x2, t2))
kp, top_p = kernels(lambda: m2.predict(x2))
print(json.dumps(dict(workload="unet_model_2d depth 4 nf 32, %d x 256x256x(5 slices + 1 prev-truth) (configs[3])" % B,
                      train_ms=s_train * 1e3, train_pixels_per_s=pix / s_train, train_kernel_ms=kt,
                      predict_ms=s_pred * 1e3, predict_pixels_per_s=pix / s_pred, predict_kernel_ms=kp,
                      fwd_gflop_per_stack=71.504, predict_conv_tflops=B * 71.504 / kp,
                      top_train=top_t, top_predict=top_p)))
