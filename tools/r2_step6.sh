#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/s6_*
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/s6_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
cat gpurun_out/s6_tests.log; tail -3 gpurun_out/s6_bench.err
