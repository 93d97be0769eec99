#!/bin/bash
# FP32 constant-bank shape variants (tools/ab_build.sh c32_*), swapped in one by one under tools/ab_rate32.py
O=gpurun_out
mkdir -p $O
cp sctl_b200/libsctl_cuda.so /tmp/ship.so
echo "| build | kernel | Nt | digits | path | grid | ms | Gpair/s | frac of FP32 peak (F_alg) | checksum | FFMA peak TF/s |" > $O/r2e_ab_rate32.md
echo "|---|---|---|---|---|---|---|---|---|---|---|" >> $O/r2e_ab_rate32.md
for v in ship "$@"; do
  if [ $v = ship ]; then cp /tmp/ship.so sctl_b200/libsctl_cuda.so; else cp tools/ab_$v.so sctl_b200/libsctl_cuda.so; fi
  timeout 400 python tools/ab_rate32.py $v 2>$O/r2e_ab32_$v.err | grep "const" >> $O/r2e_ab_rate32.md; echo "$v rc=$?"
done
cp /tmp/ship.so sctl_b200/libsctl_cuda.so
cut -d'|' -f2-10 $O/r2e_ab_rate32.md
