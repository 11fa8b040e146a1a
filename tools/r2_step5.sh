#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/s5_*
L=gpurun_out/s5_layers.log
for m in 0 1 2 3; do
  echo "== pytest march MODE=$m" >> gpurun_out/s5_tests.log
  FETAL_B200_MARCH_MODE=$m timeout 600 python -m pytest tests/test_gpu_ops.py -q -k "march" 2>&1 | tail -6 >> gpurun_out/s5_tests.log
done
for m in 0 1 2 3; do
  echo "== MARCH_MODE=$m fprop" >> $L
  FETAL_B200_MARCH_MODE=$m timeout 300 python tools/bench_layers.py fprop 8 2>&1 | grep -E "enc0b|enc1a|enc1b|dec0a|dec0b" >> $L
done
for m in 0 1 3; do
  echo "== MODE $m DEBUG=31 skeleton" >> $L
  FETAL_B200_MARCH_MODE=$m FETAL_B200_DEBUG=31 timeout 300 python tools/bench_layers.py fprop 8 dec0b 2>&1 | grep -E "dec0b" >> $L
done
echo "== MODE 3 dgrad" >> $L
FETAL_B200_MARCH_MODE=3 timeout 300 python tools/bench_layers.py dgrad 8 2>&1 | grep -E "enc0b|enc1a|enc1b|dec0a|dec0b" >> $L
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err
python bench.py --workload infer --steps 5 --warmup 3 > gpurun_out/s5_infer.json 2> gpurun_out/s5_infer.err
cat gpurun_out/s5_tests.log; cat $L
