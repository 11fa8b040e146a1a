#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/s9_*
echo "== fold" > gpurun_out/s9_noise.log
FETAL_B200_DETERMINISTIC=1 python tools/pipeline_noise.py >> gpurun_out/s9_noise.log 2>&1
echo "== separate bias grad" >> gpurun_out/s9_noise.log
FETAL_B200_DETERMINISTIC=1 FETAL_B200_SEPARATE_BIAS_GRAD=1 python tools/pipeline_noise.py >> gpurun_out/s9_noise.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/s9_tests.log
cat gpurun_out/s9_noise.log gpurun_out/s9_tests.log
