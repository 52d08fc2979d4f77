#!/bin/bash
# decode-sized GEMMs: weight-streaming kernel (default) against the 128x256-tile kernel
for s in 1 0; do
  DCMOE_FFN_STREAM=$s timeout 100 python tools/bench_decode_gemm.py 2 8 16 32 64 2>&1 | grep "GEOM\|rror"
done
