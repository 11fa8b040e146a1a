#!/bin/bash
for d in 4 5; do echo "== debug $d"; FETAL_B200_WGRAD_DEBUG=$d python tools/wgrad_fixed_cost.py 2>&1 | tail -6 | head -2; done
python - <<'PY'
# reference: an event-bracketed trivial kernel of the library (adam on 1k elements)
import sys, os, numpy as np
sys.path.insert(0, 'fetal-mri-segmentation_b200')
from fetal_net import _lib
ctx = _lib.get_context(0); lib = _lib.load()
p = np.zeros(1024, np.float32); g = np.ones(1024, np.float32); m = np.zeros(1024, np.float32); v = np.zeros(1024, np.float32)
ctx.profile(True)
for _ in range(5): _lib.check(lib.fm_op_adam(ctx.handle, _lib.fptr(p), _lib.fptr(g), _lib.fptr(m), _lib.fptr(v), 1024, 0, 1e-3))
print('tiny adam launches (ms):', [round(r[1], 4) for r in ctx.profile_records() if r[0] == 'adam'])
PY
