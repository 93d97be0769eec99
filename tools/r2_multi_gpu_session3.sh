#!/bin/bash
# Third 8-GPU session of round 2: BASELINE config 4's FP32 lines again, now with the packed FP32 loops
# (Stokes3D_DxU and Laplace3D_FxdU, 2^22 x 2^22, 8 GPUs).
N=${1:-8}
L2=${2:-22}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29721 bench.py --gpus $N --steps 2 --warmup 1 --kernel Stokes3D_DxU --real f32 --log2n $L2 > $O/r2f_bench_cfg4_stokes_dxu_f32_2p${L2}_n$N.json 2> $O/r2f_cfg4_a.err; echo "cfg4 a rc=$?"; cut -c1-260 $O/r2f_bench_cfg4_stokes_dxu_f32_2p${L2}_n$N.json
timeout 600 $TR --master-port 29722 bench.py --gpus $N --steps 2 --warmup 1 --kernel Laplace3D_FxdU --real f32 --log2n $L2 > $O/r2f_bench_cfg4_laplace_fxdu_f32_2p${L2}_n$N.json 2> $O/r2f_cfg4_b.err; echo "cfg4 b rc=$?"; cut -c1-260 $O/r2f_bench_cfg4_laplace_fxdu_f32_2p${L2}_n$N.json
