#!/bin/bash
# One single-GPU gpurun session: GPU suite, the full rate table (16 kernels x {f64, f32} x three sizes + KernelMatrix),
# FP32 packed-vs-scalar for the Helmholtz kernels when tools/ab_f32scalar.so is there, and two FP32 bench lines.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2d_gpu_tests.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/r2d_gpu_tests.txt
timeout 900 python tools/rate_table.py > $O/r2d_rate_table.md 2>$O/r2d_rate_table.err; echo "rate table rc=$?"
cat $O/r2d_rate_table.md
timeout 400 python bench.py --steps 3 --warmup 3 --real f32 --kernel Laplace3D_FxdU > $O/r2d_bench_f32_laplace_fxdu_n1.json 2>$O/r2d_bench_f32_a.err; echo "bench f32 FxdU rc=$?"; cat $O/r2d_bench_f32_laplace_fxdu_n1.json | cut -c1-600
timeout 400 python bench.py --steps 3 --warmup 3 --real f32 --kernel Stokes3D_DxU > $O/r2d_bench_f32_stokes_dxu_n1.json 2>$O/r2d_bench_f32_b.err; echo "bench f32 DxU rc=$?"; cat $O/r2d_bench_f32_stokes_dxu_n1.json | cut -c1-600
