#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s22_*
FETAL_B200_WGRAD_GEN=2 timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "wgrad" -x 2>&1 | tail -15 > $OUT/s22_tests.log
cat $OUT/s22_tests.log
FETAL_B200_WGRAD_GEN=2 timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_model.py -q -m gpu -k "cfg2 or cfg5 or train_step_matches_oracle" 2>&1 | tail -5
for g in 1 2; do
FETAL_B200_WGRAD_GEN=$g timeout 300 python bench.py --no-cpu-baseline --workload train --steps 20 > $OUT/s22_g$g.json 2> $OUT/s22_g$g.err
done
python - <<'PY'
import json
for f in ('g1','g2'):
    try: d=json.load(open('gpurun_out/s22_%s.json'%f))
    except Exception as e: print(f, 'no json', e); continue
    print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'])
    for k,v in d['kernel_breakdown'].items():
        if 'wgrad' in k: print('   %-24s n=%3d %.4f ms tf %s'%(k,v['launches'],v['ms_per_step'],v['tflops'] and round(v['tflops'])))
PY
tail -3 $OUT/s22_g2.err
FETAL_B200_WGRAD_GEN=2 timeout 300 python tools/bench_layers.py wgrad 8 2>&1 | tail -14
