#!/bin/bash
# decode-sized GEMM-1 / GEMM-2 device time: weight-streaming kernel (default) against the 128x256-tile kernel
# (DCMOE_FFN_STREAM=0); see tools/bench_decode_gemm.py
for s in 1 0; do
  DCMOE_FFN_STREAM=$s timeout 100 python tools/bench_decode_gemm.py 2 8 16 32 64 2>&1 | grep "stream=\|rror"
done
