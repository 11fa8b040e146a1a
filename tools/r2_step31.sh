#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s31_*
for m in 0 1; do
FETAL_B200_SPLIT=$m timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "march" -x 2>&1 | tail -2
FETAL_B200_SPLIT=$m timeout 300 python bench.py --no-cpu-baseline --workload train --steps 20 > $OUT/s31_m$m.json 2> $OUT/s31_m$m.err
for what in wgrad fprop; do
FETAL_B200_SPLIT=$m timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"wgrad_march|march2" --launch-skip 1 -c 1 --csv --log-file $OUT/s31_${what}_m$m.csv \
   python tools/bench_layers.py $what 8 dec0b > /dev/null 2>&1
done
done
python - <<'PY'
import json,csv
for m in (0,1):
    d=json.load(open('gpurun_out/s31_m%d.json'%m))
    print('mode',m, d['ms_per_step'], d['e2e']['ms_per_step'])
    for k,v in d['kernel_breakdown'].items():
        if 'march' in k: print('   %-24s n=%3d %.4f ms tf %s'%(k,v['launches'],v['ms_per_step'],v['tflops'] and round(v['tflops'])))
    for what in ('wgrad','fprop'):
        rows=[r for r in csv.reader(open('gpurun_out/s31_%s_m%d.csv'%(what,m))) if len(r)>10 and r[0].isdigit()]
        print('   ncu',what,[(r[-3],r[-2],r[-1]) for r in rows])
PY
