#!/bin/bash
# Tuning builds of the product library next to the shipped one: tools/ab_<name>.so with extra -D flags, e.g.
#   tools/ab_build.sh nohoist -DSCTL_CUDA_CONST_HOIST=0
#   tools/ab_build.sh nt320u3 -DSCTL_TUNE_CONST_NT=320 -DSCTL_TUNE_CONST_U=3
# A gpurun command then swaps one in (cp tools/ab_<name>.so sctl_b200/libsctl_cuda.so) between two runs of a rate tool.
set -e
name=$1; shift
cd "$(dirname "$0")/../sctl_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2 \
  -Xptxas -v "$@" -shared -o ../../tools/ab_$name.so sctl_cuda.cu -lcudart -ldl 2> ../../tools/ab_$name.ptxas.txt
grep -A2 "eval_pairs_const_kernelI[df]Li3E" ../../tools/ab_$name.ptxas.txt | grep -E "Used|spill" | head -6
