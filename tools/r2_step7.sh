#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/s7_*
FETAL_B200_TRACE=1 python tools/infer_trace.py > gpurun_out/s7_trace1.log 2>&1
FETAL_B200_TRACE=1 python tools/infer_trace.py > gpurun_out/s7_trace2.log 2>&1
python bench.py --workload infer --steps 5 > gpurun_out/s7_infer1.json 2> gpurun_out/s7_infer1.err
python bench.py --workload infer --steps 5 > gpurun_out/s7_infer2.json 2> gpurun_out/s7_infer2.err
tail -30 gpurun_out/s7_trace1.log; tail -12 gpurun_out/s7_trace2.log
