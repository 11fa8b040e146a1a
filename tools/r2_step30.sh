#!/bin/bash
# round 2, final evidence: full GPU suite, the default bench line, the reference arm, ncu launch lists (train + infer) and
# one `--set full` pass over one whole training step and one sliding-window call of the FINAL build
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s30_*
( time timeout 1500 python -m pytest tests -q -m gpu --durations=10 ) > $OUT/s30_tests.log 2>&1
python bench.py > $OUT/s30_bench.json 2> $OUT/s30_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/s30_bench_ref.json 2> $OUT/s30_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/s30_launches_train.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload train > $OUT/s30_ncu_train.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/s30_launches_infer.csv \
    python bench.py --steps 5 --no-cpu-baseline --workload infer > $OUT/s30_ncu_infer.log 2>&1
timeout 1200 ncu --set full --clock-control none -c 72 -o $OUT/s30_full_train \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload train > $OUT/s30_ncu_full_train.log 2>&1
ncu -i $OUT/s30_full_train.ncu-rep --page raw --csv > $OUT/s30_full_train_raw.csv 2>/dev/null
timeout 1200 ncu --set full --clock-control none -c 48 -o $OUT/s30_full_infer \
    python bench.py --steps 5 --no-cpu-baseline --workload infer > $OUT/s30_ncu_full_infer.log 2>&1
ncu -i $OUT/s30_full_infer.ncu-rep --page raw --csv > $OUT/s30_full_infer_raw.csv 2>/dev/null
find $OUT -name "*.ncu-rep" -size +20M -delete
tail -6 $OUT/s30_tests.log; tail -3 $OUT/s30_bench.err; ls -la $OUT | grep s30
