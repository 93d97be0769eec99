#!/bin/bash
# One single-GPU gpurun session for the packed FP32 loops: ncu --set full of the FP32 constant-bank kernel (Stokes3D_FxU,
# 2^20 x 2^20), compute-sanitizer memcheck over the FP32 parity tests and smoke().  Numbers printed under ncu are not
# bench values.
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --real f32"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_pairs_const -s 200 -c 2 -f -o $O/r2_prof_const_f32 $B > $O/r2_prof_const_f32.out 2>&1; echo "set full rc=$?"
ncu -i $O/r2_prof_const_f32.ncu-rep --page raw --csv > $O/r2_prof_const_f32_raw.csv 2>/dev/null
ncu -i $O/r2_prof_const_f32.ncu-rep --page source --csv > $O/r2_prof_const_f32_source.csv 2>/dev/null
ls -la $O/r2_prof_const_f32*
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "float32 or helmholtz_wavenumbers" > $O/r2_sanitizer_fp32.txt 2>&1; echo "memcheck rc=$?"; tail -8 $O/r2_sanitizer_fp32.txt
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_sanitizer_smoke.txt 2>&1; echo "memcheck smoke rc=$?"; tail -4 $O/r2_sanitizer_smoke.txt
