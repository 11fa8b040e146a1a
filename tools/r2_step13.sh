#!/bin/bash
# round 2, step 13: full GPU test suite, the default bench line, and the round-2 ncu evidence
# (launch lists of the train and infer workloads + one `--set full` pass over a whole training step and
# a whole sliding-window call); only CSV text travels back.
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s13_*
( time timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 ) > $OUT/s13_tests.log 2>&1
python bench.py > $OUT/s13_bench.json 2> $OUT/s13_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/s13_bench_ref.json 2> $OUT/s13_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/s13_launches_train.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload train > $OUT/s13_ncu_train.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/s13_launches_infer.csv \
    python bench.py --steps 5 --no-cpu-baseline --workload infer > $OUT/s13_ncu_infer.log 2>&1
timeout 900 ncu --set full --clock-control none -c 170 -o $OUT/s13_full_train \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload train > $OUT/s13_ncu_full_train.log 2>&1
ncu -i $OUT/s13_full_train.ncu-rep --page raw --csv > $OUT/s13_full_train_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -c 110 -o $OUT/s13_full_infer \
    python bench.py --steps 5 --no-cpu-baseline --workload infer > $OUT/s13_ncu_full_infer.log 2>&1
ncu -i $OUT/s13_full_infer.ncu-rep --page raw --csv > $OUT/s13_full_infer_raw.csv 2>/dev/null
find $OUT -name "*.ncu-rep" -size +20M -delete
tail -30 $OUT/s13_tests.log; tail -3 $OUT/s13_bench.err; ls -la $OUT
