#!/bin/bash
# Instrumented (globaltimer-stamped) build of the library for tools/attn_timing.py / tools/gemm_timing.py; never loaded by the product or the tests.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_timing
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
nvcc $F -DLC_ATTN_TIMING -c libcontinual_b200/csrc/lc_vit.cu -o tools/_timing/lc_vit_attn.o
nvcc -shared -o tools/_timing/liblc_attn_timing.so tools/_timing/lc_vit_attn.o libcontinual_b200/_C/lc_resnet.o libcontinual_b200/_C/lc_ops.o
echo tools/_timing/liblc_attn_timing.so
