#!/usr/bin/env python
"""The C library's own single-process sharding (sctl_cuda_init(n): what a C++ caller of the drop-in gets when it asks
for all GPUs of the node): Stokes3D_FxU FP64 through the host-pointer entry, 1 GPU vs all GPUs, wall clock per call.

    python tools/inprocess_multi.py [--log2n 20]    -> markdown table on stdout (copied into profiles/)
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb
from sctl_b200 import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--kernel", default="Stokes3D_FxU")
    args = ap.parse_args()
    sb.load_library()
    K = sb.KERNELS[args.kernel]
    N = 1 << args.log2n
    r_src, n_src, v_src, r_trg = synth.cube(N, N, K.SrcDim(), K.NormalDim())
    print("| GPUs (sctl_cuda_init) | call | wall ms | Gpair/s | H2D / kernel / D2H ms (max over GPUs) | max diff vs 1 GPU / vinf |\n|---|---|---|---|---|---|")
    base = None
    for n in sorted({1, 2, sb.device_count()}):
        if n > sb.device_count():
            continue
        sb.init(n)
        out = np.zeros(N * K.TrgDim())
        K.Eval(out, r_trg, r_src, n_src, v_src)  # warm-up: streams, pools, NCCL channels
        best, t = 1e30, None
        for _ in range(3):
            out[:] = 0
            t0 = time.perf_counter()
            K.Eval(out, r_trg, r_src, n_src, v_src)
            dt = (time.perf_counter() - t0) * 1e3
            if dt < best:
                best, t = dt, sb.last_timing()
        if base is None:
            base = out.copy()
        print("| %d | sctl_cuda_eval, %s FP64 2^%d x 2^%d, pageable host arrays | %.1f | %.1f | %.1f / %.1f / %.1f | %.1e |" % (
            n, args.kernel, args.log2n, args.log2n, best, float(N) * N / (best * 1e-3) * 1e-9, t["h2d_ms"], t["kernel_ms"], t["d2h_ms"],
            float(np.abs(out - base).max() / np.abs(base).max())), flush=True)
    sb.init(1)


if __name__ == "__main__":
    main()
