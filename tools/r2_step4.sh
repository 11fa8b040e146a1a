#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/s4_*
L=gpurun_out/s4_layers.log
for m in 1 0; do
for d in 0 16 18 31 15; do
  echo "== MODE $m DEBUG=$d (1 no TMA, 2 no MMA, 4 no epi global, 8 no TMEM, 16 arrive instead of commit) fprop dec0b B=8" >> $L
  FETAL_B200_MARCH_MODE=$m FETAL_B200_DEBUG=$d timeout 300 python tools/bench_layers.py fprop 8 dec0b 2>&1 | grep -E "dec0b" >> $L
done
done
for b in 1 2 4 8; do
  echo "== MODE 1 DEBUG=15 B=$b" >> $L
  FETAL_B200_MARCH_MODE=1 FETAL_B200_DEBUG=15 timeout 300 python tools/bench_layers.py fprop $b dec0b 2>&1 | grep -E "dec0b" >> $L
  echo "== MODE 1 DEBUG=0 B=$b" >> $L
  FETAL_B200_MARCH_MODE=1 timeout 300 python tools/bench_layers.py fprop $b dec0b 2>&1 | grep -E "dec0b" >> $L
done
cat $L
