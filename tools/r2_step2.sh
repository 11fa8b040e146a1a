#!/bin/bash
# round 2, step 2: token-ordered issue (MODE 2) parity / reproducibility / speed, ablations, baseline-shape tests
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -x -q -k "march" 2>&1 | tail -15 > gpurun_out/s2_tests.log
for m in 0 1 2; do
  echo "== MARCH_MODE=$m fprop" >> gpurun_out/s2_layers.log
  FETAL_B200_MARCH_MODE=$m python tools/bench_layers.py fprop 8 2>&1 | grep -E "enc0b|enc1a|enc1b|dec0a|dec0b" >> gpurun_out/s2_layers.log
done
for d in 1 2 4 5 6; do
  echo "== MODE 1 ablation DEBUG=$d (1 no TMA, 2 no MMA, 4 no epilogue stores) fprop" >> gpurun_out/s2_layers.log
  FETAL_B200_MARCH_MODE=1 FETAL_B200_DEBUG=$d python tools/bench_layers.py fprop 8 dec0b 2>&1 | grep -E "dec0b" >> gpurun_out/s2_layers.log
done
python bench.py --workload infer --steps 5 --warmup 3 > gpurun_out/s2_infer.json 2> gpurun_out/s2_infer.err
python -m pytest tests/test_gpu_baseline_shapes.py -x -q 2>&1 | tail -15 > gpurun_out/s2_baseline_tests.log
cat gpurun_out/s2_tests.log gpurun_out/s2_baseline_tests.log
