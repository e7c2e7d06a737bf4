#!/bin/sh
# diagnostics build of the CUDA library with per-stage cycle counters (not the product library)
cd "$(dirname "$0")/../av_aloha_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DAVSIM_PROFILE -shared -Xcompiler -fPIC -o libavsim_prof.so avsim_api.cu
