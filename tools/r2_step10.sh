#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/s10_*
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "up_" 2>&1 | tail -25 > gpurun_out/s10_up_tests.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/s10_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload train > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err
FETAL_B200_NO_UP_COARSE=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload train > gpurun_out/s10_bench_noup.json 2> gpurun_out/s10_bench_noup.err
python bench.py --steps 10 --workload infer --no-cpu-baseline > gpurun_out/s10_infer.json 2> gpurun_out/s10_infer.err
cat gpurun_out/s10_up_tests.log gpurun_out/s10_tests.log; tail -3 gpurun_out/s10_bench.err
