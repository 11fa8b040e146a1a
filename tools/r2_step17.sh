#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s17_*
( time timeout 1500 python -m pytest tests -q -m gpu --durations=5 ) > $OUT/s17_tests.log 2>&1
python bench.py --no-cpu-baseline --workload train --steps 20 > $OUT/s17_bench.json 2> $OUT/s17_bench.err
grep -E "passed|failed|FAILED|Error|assert" $OUT/s17_tests.log | head -30; tail -3 $OUT/s17_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s17_bench.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'])
for k,v in d['kernel_breakdown'].items(): print('%-24s n=%3d %.4f ms tf %s gbs %s frac %s'%(k,v['launches'],v['ms_per_step'],v['tflops'] and round(v['tflops']),v['gbs'] and round(v['gbs']),v['frac_of_peak'] and round(v['frac_of_peak'],3)))
PY
