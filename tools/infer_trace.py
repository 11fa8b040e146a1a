"""Host-side phase timing of one patch_wise_prediction call (configs[0]); run on the GPU box:
    FETAL_B200_TRACE=1 python tools/infer_trace.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fetal-mri-segmentation_b200"))
from fetal_net.model import unet_model_3d
from fetal_net import prediction as P

model = unet_model_3d(input_shape=(1, 64, 64, 64), n_base_filters=16, depth=4)
model.init_glorot_uniform(seed=0)
vol = np.random.default_rng(0).standard_normal((1, 256, 256, 64)).astype(np.float32)
for i in range(4):
    t0 = time.perf_counter()
    g = P._geometry(model, vol, (64, 64, 64), 0.5)
    t1 = time.perf_counter()
    idx = P.patch_plan(g["padded"], g["patch_shape"], g["prediction_shape"], 0.5)
    t2 = time.perf_counter()
    o = np.empty(g["out_dims"] + (1,), np.float64)
    t3 = time.perf_counter()
    print("geometry %.3f ms, plan %.3f ms, np.empty %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), file=sys.stderr)
    t0 = time.perf_counter()
    out = P.patch_wise_prediction(model, vol, (64, 64, 64), overlap_factor=0.5, batch_size=int(sys.argv[1]) if len(sys.argv) > 1 else 49)
    print("call %d total %.3f ms" % (i, (time.perf_counter() - t0) * 1e3), file=sys.stderr)
