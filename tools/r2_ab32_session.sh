#!/bin/bash
# One single-GPU gpurun session: FP32 parity tests on the shipped library (packed FP32 loops), then the FP32 rate table
# for the shipped library and for the tuning builds named on the command line (tools/ab_build.sh), swapped in one by one.
O=gpurun_out
mkdir -p $O
cp sctl_b200/libsctl_cuda.so /tmp/ship.so
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c_gpu_tests.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/r2c_gpu_tests.txt
echo "| build | kernel | Nt | digits | path | grid | ms | Gpair/s | frac of FP32 peak (F_alg) | checksum | FFMA peak TF/s |" > $O/r2c_ab_rate32.md
echo "|---|---|---|---|---|---|---|---|---|---|---|" >> $O/r2c_ab_rate32.md
for v in ship "$@"; do
  if [ $v = ship ]; then cp /tmp/ship.so sctl_b200/libsctl_cuda.so; else cp tools/ab_$v.so sctl_b200/libsctl_cuda.so; fi
  timeout 400 python tools/ab_rate32.py $v >> $O/r2c_ab_rate32.md 2>$O/r2c_ab32_$v.err; echo "$v rc=$?"
done
cp /tmp/ship.so sctl_b200/libsctl_cuda.so
cat $O/r2c_ab_rate32.md
