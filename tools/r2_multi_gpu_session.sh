#!/bin/bash
# One multi-GPU gpurun session of round 2: multi-GPU tests, the BoundaryIntegralOp apply pipeline on 1..N GPUs, and the
# bench lines that need all GPUs (torus points at N GPUs; BASELINE config 4's remaining kernel/precision combinations).
#   gpurun --gpus 8 -- bash tools/r2_multi_gpu_session.sh 8
N=${1:-8}
L2=22; [ $N -lt 8 ] && L2=20   # the 2^22 problems need all eight GPUs to finish in a minute
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r2_multi_gpu_tests_n$N.log 2>&1; echo "multi tests rc=$?"; tail -3 $O/r2_multi_gpu_tests_n$N.log
GL=1; for g in 2 4 8; do [ $g -le $N ] && GL=$GL,$g; done
timeout 900 python tools/bio_apply_bench.py --gpus $GL --sizes 512x256 --kernels Laplace3D_FxU,Stokes3D_FxU > $O/r2_bio_apply_multi_gpu.md 2> $O/r2_bio_apply_multi_gpu.err; echo "bio bench rc=$?"
timeout 600 python tools/bio_apply_bench.py --gpus $GL --sizes 1024x512 --kernels Laplace3D_FxU >> $O/r2_bio_apply_multi_gpu.md 2>> $O/r2_bio_apply_multi_gpu.err; echo "bio bench 2 rc=$?"
cat $O/r2_bio_apply_multi_gpu.md
timeout 600 $TR --master-port 29701 bench.py --gpus $N --steps 10 --warmup 3 --points torus > $O/r2_bench_torus_n$N.json 2> $O/r2_bench_torus_n$N.err; echo "torus rc=$?"; cut -c1-400 $O/r2_bench_torus_n$N.json
timeout 600 $TR --master-port 29702 bench.py --gpus $N --steps 2 --warmup 1 --kernel Stokes3D_DxU --real f32 --log2n $L2 > $O/r2_bench_cfg4_stokes_dxu_f32_2p${L2}_n$N.json 2> $O/r2_bench_cfg4_a.err; echo "cfg4 a rc=$?"; cut -c1-300 $O/r2_bench_cfg4_stokes_dxu_f32_2p${L2}_n$N.json
timeout 600 $TR --master-port 29703 bench.py --gpus $N --steps 2 --warmup 1 --kernel Laplace3D_FxdU --real f64 --log2n $L2 > $O/r2_bench_cfg4_laplace_fxdu_f64_2p${L2}_n$N.json 2> $O/r2_bench_cfg4_b.err; echo "cfg4 b rc=$?"; cut -c1-300 $O/r2_bench_cfg4_laplace_fxdu_f64_2p${L2}_n$N.json
timeout 600 $TR --master-port 29704 bench.py --gpus $N --steps 10 --warmup 3 > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err; echo "cube rc=$?"; cut -c1-300 $O/r2_bench_n$N.json
