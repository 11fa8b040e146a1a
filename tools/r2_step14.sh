#!/bin/bash
# round 2, step 14: new loss / dropout / 2D-Isensee tests + full suite + family bench
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s14_*
( time timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 ) > $OUT/s14_tests.log 2>&1
python bench.py --no-cpu-baseline > $OUT/s14_bench.json 2> $OUT/s14_bench.err
tail -25 $OUT/s14_tests.log; tail -3 $OUT/s14_bench.err
