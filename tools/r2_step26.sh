#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s26_*
( timeout 1500 python -m pytest tests -q -m gpu ) > $OUT/s26_tests.log 2>&1
grep -E "passed|failed|FAILED|Error" $OUT/s26_tests.log | head -20
python bench.py --no-cpu-baseline --workload train --steps 20 > $OUT/s26_bench.json 2> $OUT/s26_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s26_bench.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'])
for k,v in d['kernel_breakdown'].items():
    if v['ms_per_step']>0.06: print('   %-24s n=%3d %.4f ms tf %s'%(k,v['launches'],v['ms_per_step'],v['tflops'] and round(v['tflops'])))
PY
tail -3 $OUT/s26_bench.err
