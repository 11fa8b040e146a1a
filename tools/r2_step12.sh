#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/s12_*
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/s12_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload train > gpurun_out/s12_bench.json 2> gpurun_out/s12_bench.err
python bench.py --steps 10 --workload infer --no-cpu-baseline > gpurun_out/s12_infer.json 2> gpurun_out/s12_infer.err
FETAL_B200_TRACE=1 python tools/infer_trace.py > gpurun_out/s12_trace.log 2>&1
cat gpurun_out/s12_tests.log; tail -3 gpurun_out/s12_bench.err; tail -12 gpurun_out/s12_trace.log
