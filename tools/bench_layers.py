"""Per-layer kernel timing of the 3D U-Net's conv layers (batch 8 x 64^3, depth 4, nf 16) through the
per-op C-ABI hooks with per-launch CUDA events (ctx.profile). Prints TFLOP/s per layer and implementation.
    python tools/bench_layers.py [fprop|dgrad|wgrad] [B]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fetal-mri-segmentation_b200"))
from fetal_net import _lib  # noqa: E402

LAYERS = [  # name, S (cube edge), C1, C2, Cout
    ("enc0b", 64, 16, 0, 32), ("enc1a", 32, 32, 0, 32), ("enc1b", 32, 32, 0, 64), ("enc2a", 16, 64, 0, 64),
    ("enc2b", 16, 64, 0, 128), ("enc3a", 8, 128, 0, 128), ("enc3b", 8, 128, 0, 256), ("dec2a", 16, 256, 128, 128),
    ("dec2b", 16, 128, 0, 128), ("dec1a", 32, 128, 64, 64), ("dec1b", 32, 64, 0, 64), ("dec0a", 64, 64, 32, 32),
    ("dec0b", 64, 32, 0, 32),
]


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "fprop"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    only = sys.argv[3] if len(sys.argv) > 3 else None
    ctx = _lib.get_context(0)
    lib = _lib.load()
    rng = np.random.default_rng(0)
    rows = []
    for name, S, C1, C2, Co in LAYERS:
        if only and only not in name:
            continue
        vox = B * S ** 3
        x1 = rng.standard_normal((vox, C1), dtype=np.float32)
        x2 = rng.standard_normal((vox, C2), dtype=np.float32) if C2 else None
        w = (rng.standard_normal((3, 3, 3, C1 + C2, Co), dtype=np.float32) * 0.05)
        b = np.zeros(Co, np.float32)
        y = np.empty((vox, Co), np.float32)
        res = {}
        for impl, tag in ((0, "tap"), (2, "march"), (3, "march-shared")):
            ctx.profile(True)
            try:
                for _ in range(3):
                    if what == "fprop":
                        _lib.check(lib.fm_op_conv3d_fprop(ctx.handle, impl, _lib.fptr(x1), _lib.fptr(x2), _lib.fptr(w),
                                                          _lib.fptr(b), B, S, S, S, C1, C2, Co, 3, 1, _lib.fptr(y)))
                    elif what == "dgrad" and C2 == 0:
                        dx = np.empty((vox, C1), np.float32)
                        _lib.check(lib.fm_op_conv3d_dgrad(ctx.handle, impl, _lib.fptr(y), _lib.fptr(w), None,
                                                          B, S, S, S, C1, Co, _lib.fptr(dx)))
                    elif what == "wgrad" and C2 == 0:
                        dw = np.empty_like(w)
                        _lib.check(lib.fm_op_conv3d_wgrad(ctx.handle, impl, _lib.fptr(x1), _lib.fptr(y), B, S, S, S, C1, Co,
                                                          _lib.fptr(dw), None))
                recs = [r for r in ctx.profile_records() if r[0].startswith("conv3d")]
                if recs:
                    ms = min(r[1] for r in recs)
                    res[tag] = (ms, recs[0][2] / ms / 1e9)
            except _lib.FetalB200Error as e:
                res[tag] = None
            ctx.profile(False)
        rows.append((name, S, C1, C2, Co, res))
        print("%-6s %3d^3 %3d+%-3d->%-3d  " % (name, S, C1, C2, Co) +
              "  ".join("%s: %s" % (k, "%.3f ms %7.1f TF" % v if v else "n/a") for k, v in res.items()), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump([(r[0], r[1], r[2], r[3], r[4], r[5]) for r in rows],
              open(os.path.join(ROOT, "gpurun_out", "layers_%s.json" % what), "w"))


if __name__ == "__main__":
    main()
