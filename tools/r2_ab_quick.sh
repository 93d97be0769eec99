#!/bin/bash
# A/B of tuning builds under tools/ab_rate.py only (no tests, no bench): ship, the named builds, ship again
O=gpurun_out
mkdir -p $O
cp sctl_b200/libsctl_cuda.so /tmp/ship.so
for v in ship "$@" ship; do
  if [ $v = ship ]; then cp /tmp/ship.so sctl_b200/libsctl_cuda.so; else cp tools/ab_$v.so sctl_b200/libsctl_cuda.so; fi
  timeout 300 python tools/ab_rate.py $v ${LOG2N:-20} 2>$O/ab_quick_$v.err | tee -a $O/r2h_ab_rate.md
done
cp /tmp/ship.so sctl_b200/libsctl_cuda.so
