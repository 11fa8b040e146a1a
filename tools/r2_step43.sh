#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s43_*
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "wgrad" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_model.py tests/test_gpu_distributed.py -q -m gpu 2>&1 | tail -3
python tools/wgrad_fixed_cost.py 2>&1 | tail -6
python bench.py --no-cpu-baseline --workload train --steps 30 > $OUT/s43_bench.json 2> $OUT/s43_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s43_bench.json'))
print('train', d['ms_per_step'], d['e2e']['ms_per_step'], [(k, round(v['ms_per_step'],4), round(v['tflops'])) for k,v in d['kernel_breakdown'].items() if k=='conv3d_wgrad_march'])
PY
