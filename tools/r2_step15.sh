#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s15_*
( time timeout 1500 python -m pytest tests -q -m gpu --durations=8 ) > $OUT/s15_tests.log 2>&1
grep -E "passed|failed|FAILED|Error|assert" $OUT/s15_tests.log | head -60
