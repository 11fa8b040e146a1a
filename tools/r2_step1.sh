#!/bin/bash
# round 2, step 1: march2 kernel parity + A/B against the round-1 kernels
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -x -q -k "march" 2>&1 | tail -15 > gpurun_out/s1_tests.log
echo "== v2 layers fprop" > gpurun_out/s1_layers.log
python tools/bench_layers.py fprop 8 >> gpurun_out/s1_layers.log 2>&1
echo "== v2 layers dgrad" >> gpurun_out/s1_layers.log
python tools/bench_layers.py dgrad 8 >> gpurun_out/s1_layers.log 2>&1
echo "== v1 layers fprop" >> gpurun_out/s1_layers.log
FETAL_B200_MARCH_V1=1 python tools/bench_layers.py fprop 8 >> gpurun_out/s1_layers.log 2>&1
echo "== v1 layers dgrad" >> gpurun_out/s1_layers.log
FETAL_B200_MARCH_V1=1 python tools/bench_layers.py dgrad 8 >> gpurun_out/s1_layers.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s1_bench_v2.json 2> gpurun_out/s1_bench_v2.err
FETAL_B200_MARCH_V1=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s1_bench_v1.json 2> gpurun_out/s1_bench_v1.err
python bench.py --workload infer --steps 5 --warmup 3 > gpurun_out/s1_infer_v2.json 2> gpurun_out/s1_infer_v2.err
FETAL_B200_MARCH_V1=1 python bench.py --workload infer --steps 5 --warmup 3 > gpurun_out/s1_infer_v1.json 2> gpurun_out/s1_infer_v1.err
cat gpurun_out/s1_tests.log
