"""Runs every per-op parity case in its own subprocess (a trap in one kernel must not poison the rest),
prints a table and writes gpurun_out/diag.json. Usage on the GPU box:  python tools/gpu_diag.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fetal-mri-segmentation_b200"))

CHILD = r"""
import sys, json, time
sys.path.insert(0, %(root)r); sys.path.insert(0, %(pkg)r)
from tests import gpu_cases as gc
from fetal_net import _lib
ctx = _lib.get_context(0)
kind, impl, ci = %(kind)r, %(impl)d, %(ci)d
t0 = time.time()
if kind == 'mfprop': r = gc.conv_fprop_case(ctx, 2, gc.MARCH_CASES[ci])
elif kind == 'mdgrad': r = gc.conv_dgrad_case(ctx, 2, gc.MARCH_CASES[ci])
elif kind == 'mwgrad': r = gc.conv_wgrad_case(ctx, 2, gc.WGRAD_MARCH_CASES[ci])
elif kind == 'fprop': r = gc.conv_fprop_case(ctx, impl, gc.CONV_CASES[ci])
elif kind == 'dgrad': r = gc.conv_dgrad_case(ctx, impl, gc.CONV_CASES[ci])
elif kind == 'wgrad': r = gc.conv_wgrad_case(ctx, impl, gc.CONV_CASES[ci])
else: r = getattr(gc, kind + '_case')(ctx)
print('RESULT', json.dumps([bool(r[0]), float(r[1]), time.time() - t0]))
"""


def main():
    from tests import gpu_cases as gc
    jobs = []
    for kind in ("maxpool", "upsample", "dice", "adam"):
        jobs.append((kind, 0, 0, kind))
    for ci, c in enumerate(gc.CONV_CASES):
        for impl in (1, 0):
            jobs.append(("fprop", impl, ci, "fprop[%s] %s" % ("tc" if impl == 0 else "simt", c[0])))
    for ci, c in enumerate(gc.CONV_CASES):
        if c[6] == 0 and c[8] == 3:
            jobs.append(("dgrad", 0, ci, "dgrad[tc] %s" % c[0]))
            jobs.append(("wgrad", 1, ci, "wgrad[simt] %s" % c[0]))
            jobs.append(("wgrad", 0, ci, "wgrad[tc] %s" % c[0]))
    for ci, c in enumerate(gc.MARCH_CASES):
        jobs.append(("mfprop", 2, ci, "fprop[march] %s" % c[0]))
        if c[6] == 0:
            jobs.append(("mdgrad", 2, ci, "dgrad[march] %s" % c[0]))
    for ci, c in enumerate(gc.WGRAD_MARCH_CASES):
        jobs.append(("mwgrad", 2, ci, "wgrad[march] %s" % c[0]))
    only = sys.argv[1] if len(sys.argv) > 1 else None
    results = {}
    for kind, impl, ci, label in jobs:
        if only and only not in label:
            continue
        code = CHILD % dict(root=ROOT, pkg=os.path.join(ROOT, "fetal-mri-segmentation_b200"), kind=kind, impl=impl, ci=ci)
        try:
            p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
            if line:
                ok, worst, dt = json.loads(line[0][7:])
                results[label] = dict(ok=ok, worst=worst, secs=dt)
                print("%-44s %s worst/tol=%.4g  (%.1fs)" % (label, "PASS" if ok else "FAIL", worst, dt), flush=True)
            else:
                tail = (p.stdout + p.stderr).strip().splitlines()[-6:]
                results[label] = dict(ok=False, error=tail)
                print("%-44s ERROR rc=%d\n    %s" % (label, p.returncode, "\n    ".join(tail)), flush=True)
        except subprocess.TimeoutExpired:
            results[label] = dict(ok=False, error="timeout")
            print("%-44s TIMEOUT" % label, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w"), indent=1)
    nfail = sum(1 for r in results.values() if not r["ok"])
    print("%d cases, %d failed" % (len(results), nfail))


if __name__ == "__main__":
    main()
