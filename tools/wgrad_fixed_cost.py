"""Fixed cost of a marching-wgrad launch: time vs planes per CTA (N=8, Y=Z=64, Cin=Cout=32, X varied).
    python tools/wgrad_fixed_cost.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fetal-mri-segmentation_b200"))
from fetal_net import _lib  # noqa: E402


def main():
    ctx = _lib.get_context(0)
    lib = _lib.load()
    rng = np.random.default_rng(0)
    B, Y, Z, C1, Co = 8, 64, 64, 32, 32
    for X in (2, 4, 8, 16, 32, 64):
        vox = B * X * Y * Z
        x1 = rng.standard_normal((vox, C1), dtype=np.float32)
        y = rng.standard_normal((vox, Co), dtype=np.float32)
        dw = np.empty((3, 3, 3, C1, Co), np.float32)
        ctx.profile(True)
        for _ in range(4):
            _lib.check(lib.fm_op_conv3d_wgrad(ctx.handle, 2, _lib.fptr(x1), _lib.fptr(y), B, X, Y, Z, C1, Co,
                                              _lib.fptr(dw), None))
        recs = [r for r in ctx.profile_records() if r[0].startswith("conv3d_wgrad")]
        ctx.profile(False)
        ms = min(r[1] for r in recs)
        print("X=%2d planes/CTA=%6.1f  %.4f ms  %.1f TF/s" % (X, 256 * X / 148.0, ms, recs[0][2] / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
