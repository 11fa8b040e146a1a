#!/bin/bash
# round 2, step 16 (2 GPUs): NCCL equality tests + the data-parallel bench line
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s16_*
( time timeout 900 python -m pytest tests/test_gpu_distributed.py -q -m gpu --durations=5 ) > $OUT/s16_tests.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/s16_bench_n2.json 2> $OUT/s16_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $OUT/s16_bench_ref_n2.json 2> $OUT/s16_bench_ref_n2.err
grep -E "passed|failed|FAILED|Error|skipped" $OUT/s16_tests.log | head; tail -5 $OUT/s16_bench_n2.err; wc -c $OUT/s16_bench_n2.json
