#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s41_*
( timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ) > $OUT/s41_tests.log 2>&1
grep -E "passed|failed|^FAILED|^E  " $OUT/s41_tests.log | head -8
python bench.py --no-cpu-baseline --steps 20 > $OUT/s41_bench.json 2> $OUT/s41_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s41_bench.json'))
print('train', d['ms_per_step'], d['e2e']['ms_per_step'])
for k,v in d['kernel_breakdown'].items():
    if 'head' in k: print('   %-24s %.4f ms gbs %s'%(k,v['ms_per_step'],v['gbs'] and round(v['gbs'])))
i=d['infer']; print('infer', i['ms_per_step'], i['e2e']['ms_per_step'])
for k,v in i['kernel_breakdown'].items():
    if 'head' in k: print('   %-24s %.4f ms gbs %s'%(k,v['ms'],v['gbs'] and round(v['gbs'])))
for fam in ('isensee','unet2d'):
    f=d['families'][fam]; print(fam, f['ms_per_step'], f['predict']['ms_per_call'])
PY
tail -2 $OUT/s41_bench.err
