#!/bin/bash
# tensor-pipe rate of the MMA shapes the marching kernels issue (tools/mma_rate.cu, built here with nvcc)
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s50_*
timeout 60 tools/mma_rate.bin > $OUT/s50_mma_rate.txt 2>&1
cat $OUT/s50_mma_rate.txt
