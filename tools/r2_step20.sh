#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s20_*
( timeout 1500 python -m pytest tests -q -m gpu -x ) > $OUT/s20_tests.log 2>&1
python bench.py --no-cpu-baseline --steps 12 > $OUT/s20_bench.json 2> $OUT/s20_bench.err
grep -E "passed|failed|FAILED|Error|assert" $OUT/s20_tests.log | head -20; tail -3 $OUT/s20_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s20_bench.json'))
print('train', d['ms_per_step'], d['e2e']['ms_per_step'], 'infer', d['infer']['ms_per_step'], d['infer']['e2e']['ms_per_step'])
for fam in ('isensee','unet2d'):
    f=d['families'][fam]; print(fam, f['ms_per_step'], f['e2e']['ms_per_step'], f['predict']['ms_per_call'])
    for k,v in f['kernel_breakdown'].items(): print('   %-24s n=%3d %.4f ms tf %s gbs %s'%(k,v['launches'],v['ms_per_step'],v['tflops'] and round(v['tflops']),v['gbs'] and round(v['gbs'])))
PY
