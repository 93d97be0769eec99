#!/bin/bash
# Round 2, last session, final run: KernelMatrix A/B of the shipped library (K1 == 1 kernels cut into 32 CTAs per SM) against
# the library of the commit before the rewrite, then the final single-GPU validation of the tree as the driver runs it: GPU
# suite, smoke(), the reference arm and our arm of bench.py (now with roofline.kernel_matrix_view), ncu launch list.
O=gpurun_out
mkdir -p $O
nvidia-smi -L
timeout 300 python tools/kmat_rate.py new > $O/r2m_kmat_rate_new.md 2>$O/r2m_kmat_rate_new.err; echo "rate new rc=$?"
cp sctl_b200/libsctl_cuda.so /tmp/new.so && cp tools/ab_kmat_old.so sctl_b200/libsctl_cuda.so
timeout 300 python tools/kmat_rate.py old > $O/r2m_kmat_rate_old.md 2>$O/r2m_kmat_rate_old.err; echo "rate old rc=$?"
cp /tmp/new.so sctl_b200/libsctl_cuda.so
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2m_gpu_tests_final.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/r2m_gpu_tests_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2m_smoke_final.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2m_smoke_final.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2m_bench_reference_arm_final.json 2>$O/r2m_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/r2m_bench_reference_arm_final.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2m_bench_n1_final.json 2>$O/r2m_bench_n1.err; echo "bench rc=$?"; cut -c1-300 $O/r2m_bench_n1_final.json; grep -o '"kernel_matrix_view": {[^}]*}' $O/r2m_bench_n1_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file $O/r2m_launches_bench_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-kmat-view > $O/r2m_launches_bench_final.out 2>&1; echo "launch list rc=$?"
