#!/bin/bash
# Round 2, last session (second run: rows-per-round staging, rolled copy-out): the rebuilt KernelMatrix kernel (csrc/kmat_kernel.cuh) against the library of the previous commit
# (tools/ab_kmat_old.so, built from HEAD~ before the change) on one box, its ncu capture, then the final single-GPU
# validation of the tree as the driver runs it.
O=gpurun_out
mkdir -p $O
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kernel_matrix" > $O/r2i_kmat_tests.txt 2>&1; echo "kmat tests rc=$?"; tail -5 $O/r2i_kmat_tests.txt
timeout 300 python tools/kmat_rate.py new > $O/r2i_kmat_rate_new.md 2>$O/r2i_kmat_rate_new.err; echo "rate new rc=$?"
cp sctl_b200/libsctl_cuda.so /tmp/new.so && cp tools/ab_kmat_old.so sctl_b200/libsctl_cuda.so
timeout 300 python tools/kmat_rate.py old > $O/r2i_kmat_rate_old.md 2>$O/r2i_kmat_rate_old.err; echo "rate old rc=$?"
cp /tmp/new.so sctl_b200/libsctl_cuda.so
paste -d'\n' $O/r2i_kmat_rate_old.md $O/r2i_kmat_rate_new.md | grep -E "8192 x 8192|2048 x 65536" | cut -c1-120
KMAT_KERNELS=Laplace3D_FxU,Stokes3D_FxU KMAT_SHAPES=1 KMAT_REPS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kernel_matrix_kernel -c 4 -f -o $O/r2i_kmat python tools/kmat_rate.py ncu > $O/r2i_kmat_ncu.out 2>&1; echo "ncu rc=$?"
ncu -i $O/r2i_kmat.ncu-rep --page raw --csv > $O/r2i_kmat_raw.csv 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2i_gpu_tests_final.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/r2i_gpu_tests_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2i_smoke_final.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2i_smoke_final.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2i_bench_reference_arm_final.json 2>$O/r2i_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/r2i_bench_reference_arm_final.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2i_bench_n1_final.json 2>$O/r2i_bench_n1.err; echo "bench rc=$?"; cut -c1-400 $O/r2i_bench_n1_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file $O/r2i_launches_bench_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2i_launches_bench_final.out 2>&1; echo "launch list rc=$?"
