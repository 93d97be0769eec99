#!/bin/bash
# Round 2, last session: KernelMatrix grid size A/B (CTAs per SM a large matrix is cut into: 32 / 64 shipped / 128; tuning
# builds from tools/ab_build.sh), then the final single-GPU validation of the tree as the driver runs it: GPU suite, smoke(),
# the reference arm and our arm of bench.py, and the ncu launch list of the bench command.
O=gpurun_out
mkdir -p $O
nvidia-smi -L
export KMAT_KERNELS=Laplace3D_FxU,Laplace3D_DxU,Stokes3D_FxU,Helmholtz3D_FxU KMAT_SHAPES=4
cp sctl_b200/libsctl_cuda.so /tmp/shipped.so
for v in kmat_y32 kmat_y128; do
  if [ -f tools/ab_$v.so ]; then
    cp tools/ab_$v.so sctl_b200/libsctl_cuda.so
    timeout 200 python tools/kmat_rate.py $v > $O/r2l_kmat_rate_$v.md 2>$O/r2l_kmat_rate_$v.err; echo "rate $v rc=$?"
  fi
done
cp /tmp/shipped.so sctl_b200/libsctl_cuda.so
timeout 200 python tools/kmat_rate.py shipped > $O/r2l_kmat_rate_shipped.md 2>$O/r2l_kmat_rate_shipped.err; echo "rate shipped rc=$?"
unset KMAT_KERNELS KMAT_SHAPES
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2l_gpu_tests_final.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/r2l_gpu_tests_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2l_smoke_final.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2l_smoke_final.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2l_bench_reference_arm_final.json 2>$O/r2l_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/r2l_bench_reference_arm_final.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2l_bench_n1_final.json 2>$O/r2l_bench_n1.err; echo "bench rc=$?"; cut -c1-400 $O/r2l_bench_n1_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file $O/r2l_launches_bench_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2l_launches_bench_final.out 2>&1; echo "launch list rc=$?"
