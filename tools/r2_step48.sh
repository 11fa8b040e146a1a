#!/bin/bash
# per-instruction stall samples (ncu source page) of ONE launch of the marching fprop (three-issuer mode, 32->32 at
# 8 x 64^3) and ONE of the marching wgrad (same layer)
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s48_*
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv3d_march2 -s 3 -c 1 -o $OUT/s48_fprop \
    python tools/bench_layers.py fprop 8 dec0b > $OUT/s48_fprop.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv3d_wgrad_march -s 1 -c 1 -o $OUT/s48_wgrad \
    python tools/bench_layers.py wgrad 8 dec0b > $OUT/s48_wgrad.log 2>&1
ls -la $OUT | grep s48; tail -2 $OUT/s48_fprop.log $OUT/s48_wgrad.log
