#!/bin/bash
# One single-GPU gpurun session: the GPU suite on the shipped library, then the A/B rate runs of the tuning builds
# (tools/ab_build.sh) by swapping each library in, then bench.py on the shipped library again.
O=gpurun_out
mkdir -p $O
cp sctl_b200/libsctl_cuda.so /tmp/ship.so
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2b_gpu_tests.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/r2b_gpu_tests.txt
echo "| build | kernel | digits | best ms | median ms | Gpair/s | frac of peak (F_alg) | checksum | DFMA peak TF/s |" > $O/r2b_ab_rate.md
echo "|---|---|---|---|---|---|---|---|---|" >> $O/r2b_ab_rate.md
for v in ship "$@" ship; do
  if [ $v = ship ]; then cp /tmp/ship.so sctl_b200/libsctl_cuda.so; else cp tools/ab_$v.so sctl_b200/libsctl_cuda.so; fi
  timeout 300 python tools/ab_rate.py $v >> $O/r2b_ab_rate.md 2>$O/r2b_ab_$v.err; echo "$v rc=$?"
done
cp /tmp/ship.so sctl_b200/libsctl_cuda.so
cat $O/r2b_ab_rate.md
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2b_bench_n1.json 2>$O/r2b_bench_n1.err; echo "bench rc=$?"; cat $O/r2b_bench_n1.json
