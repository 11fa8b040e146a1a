#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "fast_inference" 2>&1 | grep -E "^E|passed|failed" | head -20
