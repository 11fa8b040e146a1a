#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
FETAL_B200_WGRAD_GEN=3 timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_model.py -q -m gpu 2>&1 | tail -3
for g in 1 3 1 3; do
FETAL_B200_WGRAD_GEN=$g timeout 300 python bench.py --no-cpu-baseline --workload train --steps 30 > $OUT/s38_g$g.json 2> $OUT/s38_g$g.err
python - <<PY
import json
d=json.load(open('gpurun_out/s38_g$g.json'))
print('gen $g', d['ms_per_step'], d['e2e']['ms_per_step'], [ (k, round(v['ms_per_step'],4)) for k,v in d['kernel_breakdown'].items() if k=='conv3d_wgrad_march'])
PY
done
