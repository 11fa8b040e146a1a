#!/bin/bash
# deeper slab rings (marching fprop/dgrad: up to 12 slots per dz ring instead of 4; marching wgrad: 5 instead of 4):
# parity of the marching cases, then per-layer A/B against the old depths
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s46_*
( timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "march" -p no:cacheprovider ) > $OUT/s46_ops.log 2>&1
tail -3 $OUT/s46_ops.log
for what in fprop wgrad dgrad; do
  echo "== $what new" >> $OUT/s46_layers.log
  timeout 200 python tools/bench_layers.py $what >> $OUT/s46_layers.log 2>&1
  echo "== $what old" >> $OUT/s46_layers.log
  FETAL_B200_MARCH_STAGES=12 FETAL_B200_WGRAD_S3=4 timeout 200 python tools/bench_layers.py $what >> $OUT/s46_layers.log 2>&1
done
grep -E "==|march" $OUT/s46_layers.log | cut -c1-160
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload train > $OUT/s46_bench.json 2> $OUT/s46_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s46_bench.json') if l.startswith('{')][-1])
print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'))
for k,v in d['kernel_breakdown'].items():
    if 'march' in k: print(k, v['ms_per_step'], v.get('tflops'))
PY
