#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s35_*
python __graft_entry__.py --smoke 2>&1 | tail -2
for i in 1 2 3; do
( timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ) > $OUT/s35_tests_$i.log 2>&1
grep -E "passed|failed|^FAILED|^E  " $OUT/s35_tests_$i.log | head -8
done
