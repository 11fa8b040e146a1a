#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/s11_*
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "first or up_" 2>&1 | tail -25 > gpurun_out/s11_new_tests.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/s11_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload train > gpurun_out/s11_bench.json 2> gpurun_out/s11_bench.err
FETAL_B200_SIMT_FIRST=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload train > gpurun_out/s11_bench_simtfirst.json 2> gpurun_out/s11_bench_simtfirst.err
cat gpurun_out/s11_new_tests.log gpurun_out/s11_tests.log; tail -3 gpurun_out/s11_bench.err
