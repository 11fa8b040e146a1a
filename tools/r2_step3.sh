#!/bin/bash
# round 2, step 3: 8-warp pipelined epilogue; MODE 2 deadlock fix; ablation floors
mkdir -p gpurun_out; rm -f gpurun_out/s3_*
python -m pytest tests/test_gpu_ops.py -x -q -k "march" 2>&1 | tail -15 > gpurun_out/s3_tests.log
for m in 0 1 2; do
  echo "== MARCH_MODE=$m fprop" >> gpurun_out/s3_layers.log
  FETAL_B200_MARCH_MODE=$m timeout 300 python tools/bench_layers.py fprop 8 2>&1 | grep -E "enc0b|enc1a|enc1b|dec0a|dec0b" >> gpurun_out/s3_layers.log
done
for d in 2 6 10 14 15; do
  echo "== MODE 1 ablation DEBUG=$d (1 no TMA, 2 no MMA, 4 no epilogue global ld/st, 8 no TMEM ld/st) fprop" >> gpurun_out/s3_layers.log
  FETAL_B200_MARCH_MODE=1 FETAL_B200_DEBUG=$d timeout 300 python tools/bench_layers.py fprop 8 dec0b 2>&1 | grep -E "dec0b" >> gpurun_out/s3_layers.log
done
echo "== MODE 1 dgrad" >> gpurun_out/s3_layers.log
FETAL_B200_MARCH_MODE=1 timeout 300 python tools/bench_layers.py dgrad 8 2>&1 | grep -E "enc0b|enc1a|enc1b|dec0a|dec0b" >> gpurun_out/s3_layers.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
python bench.py --workload infer --steps 5 --warmup 3 > gpurun_out/s3_infer.json 2> gpurun_out/s3_infer.err
python -m pytest tests/test_gpu_baseline_shapes.py -q 2>&1 | tail -25 > gpurun_out/s3_baseline_tests.log
cat gpurun_out/s3_tests.log gpurun_out/s3_baseline_tests.log
