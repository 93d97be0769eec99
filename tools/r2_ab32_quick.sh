#!/bin/bash
# A/B of tuning builds under tools/ab_rate32.py (constant-bank rows) and a Stokes3D_FxU FP64 line per build
O=gpurun_out
mkdir -p $O
cp sctl_b200/libsctl_cuda.so /tmp/ship.so
for v in ship "$@"; do
  if [ $v = ship ]; then cp /tmp/ship.so sctl_b200/libsctl_cuda.so; else cp tools/ab_$v.so sctl_b200/libsctl_cuda.so; fi
  timeout 400 python tools/ab_rate32.py $v 2>$O/ab32_quick_$v.err | grep "const" | tee -a $O/r2i_ab_rate32.md | cut -d'|' -f2-10
  timeout 300 python tools/ab_rate.py $v 19 2>>$O/ab32_quick_$v.err | grep Stokes3D_FxU | tee -a $O/r2i_ab_rate.md | cut -d'|' -f2-8
done
cp /tmp/ship.so sctl_b200/libsctl_cuda.so
