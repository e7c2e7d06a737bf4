#!/bin/sh
# First GPU call of the next round: build the compile-time variants that were prepared but not measured, and time them with
# tools/variant_bench.py on the steady-state bench workload (tools/_steady/steady_B4096.npz; made by `python tools/steady.py make 4096 8`
# if missing).  Usage (on the GPU box):  sh tools/round2_sweep.sh > gpurun_out/round2_sweep.log 2>&1
#   - more helper warps at a lower register cap: AV_MAX_WARPS = 18 / 20 (cap 113 / 102 registers; the default 16 -> 128)
#   - queue sorted by non-pooled cycles only (AVSIM_KEY=0), barrier once per substep (sync 1)
cd "$(dirname "$0")/.."
CS=av_aloha_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
[ -f tools/_steady/steady_B4096.npz ] || { python tools/steady.py make 4096 8 && mkdir -p tools/_steady && cp gpurun_out/steady_B4096.npz tools/_steady/; }
for w in 18 20; do
    [ -f $CS/libavsim_w$w.so ] || (cd $CS && nvcc $F -DAV_MAX_WARPS=$w -o libavsim_w$w.so avsim_api.cu)
done
echo "== IK kernels"; python tools/ik_bench.py
echo "== default library"; python tools/variant_bench.py 4096 8 16:1:2:11,16:1:1:11
echo "== AVSIM_KEY=0"; AVSIM_KEY=0 python tools/variant_bench.py 4096 8 16:1:2:11
for w in 18 20; do
    echo "== AV_MAX_WARPS=$w"; AVSIM_LIB=$PWD/$CS/libavsim_w$w.so python tools/variant_bench.py 4096 8 $w:1:2:11,$w:1:2:10
done
