#!/bin/bash
# where does a plane of the marching fprop go? FETAL_B200_DEBUG ablations (1 no slab TMA, 2 no MMAs, 4 no epilogue global
# traffic, 8 no TMEM ld/st, 16 plain arrivals instead of tcgen05.commit) on the 16->32 and 32->32 layers at 8 x 64^3
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s45_*
for d in 0 1 2 4 8 16 3 12 14 15 18; do
  echo "== debug $d" >> $OUT/s45_ablate.log
  FETAL_B200_DEBUG=$d timeout 120 python tools/bench_layers.py fprop 8 0b >> $OUT/s45_ablate.log 2>&1
done
cat $OUT/s45_ablate.log
