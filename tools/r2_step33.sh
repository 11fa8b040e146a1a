#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s33_*
timeout 900 python -m pytest tests/test_sampler.py tests/test_gpu_model.py -q -m gpu 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 20 > $OUT/s33_bench.json 2> $OUT/s33_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s33_bench.json'))
print('train', d['ms_per_step'], d['e2e']['ms_per_step'], 'sampled', d['train_sampled'])
PY
tail -3 $OUT/s33_bench.err
