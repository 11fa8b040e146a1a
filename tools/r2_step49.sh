#!/bin/bash
# issue modes of the marching fprop/dgrad in the training passes: 1 (one warp per dz) against the unordered plane owners
# (3 warps / 4 warps, FETAL_B200_TRAIN_MARCH_MODE=3|4): per-layer times, then the GPU suite and the training bench
# under the fastest mode
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s49_*
python - <<'PY'
import os, re, subprocess, sys, json
out = "gpurun_out/s49_modes.log"
tot = {}
with open(out, "w") as f:
    for mode in ("1", "3", "4"):
        tot[mode] = 0.0
        for what in ("fprop", "dgrad"):
            env = dict(os.environ, FETAL_B200_TRAIN_MARCH_MODE=mode)
            try:
                r = subprocess.run([sys.executable, "tools/bench_layers.py", what, "8", "0"], env=env, capture_output=True,
                                   text=True, timeout=150)
                txt = r.stdout + r.stderr[-500:]
            except subprocess.TimeoutExpired:
                txt = "TIMEOUT\n"
                tot[mode] += 1e9
            f.write("== mode %s %s\n%s" % (mode, what, txt))
            for m in re.finditer(r"march-shared:\s+([0-9.]+) ms", txt):
                tot[mode] += float(m.group(1))
            if "march-shared: n/a" in txt or "march-shared" not in txt:
                tot[mode] += 1e9
    f.write("totals %s\n" % json.dumps(tot))
best = min(tot, key=tot.get)
open("gpurun_out/s49_best", "w").write(best)
print(open(out).read()[-3500:])
print("totals", tot, "best", best)
PY
BEST=$(cat $OUT/s49_best)
export FETAL_B200_TRAIN_MARCH_MODE=$BEST
( timeout 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider ) > $OUT/s49_tests.log 2>&1
tail -3 $OUT/s49_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload train > $OUT/s49_bench.json 2> $OUT/s49_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s49_bench.json') if l.startswith('{')][-1])
print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'))
for k,v in d['kernel_breakdown'].items():
    if 'march' in k: print(k, v['ms_per_step'], v.get('tflops'))
PY
