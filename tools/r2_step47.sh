#!/bin/bash
# final build of round 2: full GPU suite, smoke, default bench line, reference arm, ncu launch lists (train + infer)
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s47_*
python __graft_entry__.py --smoke 2>&1 | tail -2
( time timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider ) > $OUT/s47_tests.log 2>&1
tail -5 $OUT/s47_tests.log
python bench.py > $OUT/s47_bench.json 2> $OUT/s47_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/s47_bench_ref.json 2> $OUT/s47_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/s47_launches_train.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload train > $OUT/s47_ncu_train.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/s47_launches_infer.csv \
    python bench.py --steps 5 --no-cpu-baseline --workload infer > $OUT/s47_ncu_infer.log 2>&1
tail -c 600 $OUT/s47_bench.json; ls -la $OUT | grep s47
