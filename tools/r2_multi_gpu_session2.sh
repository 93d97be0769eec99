#!/bin/bash
# Second 8-GPU session of round 2: the apply pipeline after the density-broadcast / worker-thread change, and the torus
# bench at 4 GPUs (the 8-GPU line is in the first session, 1 and 2 GPUs come from cheaper boxes).
O=gpurun_out
export SCTL_CUDA_BIO_TRACE=1
timeout 900 python tools/bio_apply_bench.py --gpus 1,2,4,8 --sizes 512x256 --kernels Laplace3D_FxU,Stokes3D_FxU > $O/r2_bio_apply_multi_gpu.md 2> $O/r2_bio_apply_multi_gpu.err; echo "bio bench rc=$?"
timeout 600 python tools/bio_apply_bench.py --gpus 1,2,4,8 --sizes 1024x512 --kernels Laplace3D_FxU >> $O/r2_bio_apply_multi_gpu.md 2>> $O/r2_bio_apply_multi_gpu.err; echo "bio bench 2 rc=$?"
cat $O/r2_bio_apply_multi_gpu.md; grep "sctl_cuda bio" $O/r2_bio_apply_multi_gpu.err | tail -12
unset SCTL_CUDA_BIO_TRACE
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 4 --steps 10 --warmup 3 --points torus > $O/r2_bench_torus_n4.json 2> $O/r2_bench_torus_n4.err; echo "torus n4 rc=$?"; cut -c1-200 $O/r2_bench_torus_n4.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r2_bench_n4.json 2> $O/r2_bench_n4.err; echo "cube n4 rc=$?"; cut -c1-200 $O/r2_bench_n4.json
