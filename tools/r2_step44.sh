#!/bin/bash
# marching fprop/dgrad with alternate-plane epilogue groups: parity of every marching case + the network tests that run on
# it, then per-kernel times of one training step
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s44_*
( timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "march or first or up_" -p no:cacheprovider ) > $OUT/s44_ops.log 2>&1
tail -3 $OUT/s44_ops.log
( timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_prediction.py -q -m gpu -x -p no:cacheprovider ) > $OUT/s44_model.log 2>&1
tail -3 $OUT/s44_model.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload train > $OUT/s44_bench.json 2> $OUT/s44_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/s44_launches_train.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload train > $OUT/s44_ncu_train.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s44_bench.json') if l.startswith('{')][-1])
print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'))
for k,v in d['kernel_breakdown'].items():
    if 'march' in k: print(k, v['ms_per_step'], v.get('tflops'))
PY
