#!/bin/bash
# Round 2, last session (fourth run: lean copy-out with an aligned fast path): KernelMatrix against the
# library of the commit before the rewrite (tools/ab_kmat_old.so) on one box, ncu capture, GPU suite and smoke.
O=gpurun_out
mkdir -p $O
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kernel_matrix" > $O/r2k_kmat_tests.txt 2>&1; echo "kmat tests rc=$?"; tail -5 $O/r2k_kmat_tests.txt
timeout 300 python tools/kmat_rate.py new > $O/r2k_kmat_rate_new.md 2>$O/r2k_kmat_rate_new.err; echo "rate new rc=$?"
cp sctl_b200/libsctl_cuda.so /tmp/new.so && cp tools/ab_kmat_old.so sctl_b200/libsctl_cuda.so
timeout 300 python tools/kmat_rate.py old > $O/r2k_kmat_rate_old.md 2>$O/r2k_kmat_rate_old.err; echo "rate old rc=$?"
cp /tmp/new.so sctl_b200/libsctl_cuda.so
KMAT_KERNELS=Laplace3D_FxU,Stokes3D_FxU,Helmholtz3D_FxU KMAT_SHAPES=2 KMAT_REPS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kernel_matrix_kernel -c 12 -f -o $O/r2k_kmat python tools/kmat_rate.py ncu > $O/r2k_kmat_ncu.out 2>&1; echo "ncu rc=$?"
ncu -i $O/r2k_kmat.ncu-rep --page raw --csv > $O/r2k_kmat_raw.csv 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2k_gpu_tests_final.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/r2k_gpu_tests_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2k_smoke_final.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2k_smoke_final.txt
