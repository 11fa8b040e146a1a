#!/bin/bash
# round 2, step 34 (8 GPUs): the data-parallel bench line of the final build + the reference arm under torchrun
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s34_*
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/s34_bench_n8.json 2> $OUT/s34_bench_n8.err
tail -3 $OUT/s34_bench_n8.err; wc -c $OUT/s34_bench_n8.json
