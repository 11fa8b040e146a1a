#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/s8_*
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/s8_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err
FETAL_B200_SEPARATE_BIAS_GRAD=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload train > gpurun_out/s8_bench_sepbias.json 2> gpurun_out/s8_bench_sepbias.err
FETAL_B200_PW_SUB=0 python bench.py --steps 10 --workload infer --no-cpu-baseline > gpurun_out/s8_infer_nosub.json 2> gpurun_out/s8_infer_nosub.err
FETAL_B200_TRACE=1 python tools/infer_trace.py > gpurun_out/s8_trace.log 2>&1
python tools/bench_layers.py wgrad 8 > gpurun_out/s8_wgrad.log 2>&1
cat gpurun_out/s8_tests.log; tail -3 gpurun_out/s8_bench.err; tail -12 gpurun_out/s8_trace.log; cat gpurun_out/s8_wgrad.log
