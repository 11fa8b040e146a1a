#!/bin/bash
# round 2, step 28 (4 GPUs): the data-parallel bench line of the final build
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s28_*
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 4 --steps 10 --warmup 3 > $OUT/s28_bench_n4.json 2> $OUT/s28_bench_n4.err
tail -3 $OUT/s28_bench_n4.err; wc -c $OUT/s28_bench_n4.json
