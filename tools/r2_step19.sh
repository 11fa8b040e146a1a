#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s19_*
python bench.py --no-cpu-baseline --workload train --steps 20 > $OUT/s19_a.json 2> $OUT/s19_a.err
FETAL_B200_UP_COARSE_ALL=1 python bench.py --no-cpu-baseline --workload train --steps 20 > $OUT/s19_b.json 2> $OUT/s19_b.err
FETAL_B200_UP_COARSE_ALL=1 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_model.py -q -m gpu -k "cfg2 or train_step_matches_oracle" 2>&1 | tail -3
python - <<'PY'
import json
for f in ('a','b'):
    d=json.load(open('gpurun_out/s19_%s.json'%f))
    print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'])
    for k,v in d['kernel_breakdown'].items():
        if v['ms_per_step']>0.04: print('   %-24s n=%3d %.4f ms tf %s'%(k,v['launches'],v['ms_per_step'],v['tflops'] and round(v['tflops'])))
PY
tail -3 $OUT/s19_b.err
