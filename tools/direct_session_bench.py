#!/usr/bin/env python
"""Per-apply time of the device-resident direct evaluator (sctl_cuda_direct_*) against the host-pointer Eval entry
(sctl_cuda_eval: uploads coordinates every call), GMRES pattern: fixed coordinates, new density every apply.

    python tools/direct_session_bench.py [--gpus N]    -> markdown table on stdout (copied into profiles/)
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb
from sctl_b200 import synth


def median_ms(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    sb.load_library()
    print("| kernel | Real | Ns = Nt | host-pointer Eval ms/apply | resident session ms/apply (%d GPU) | speed-up | Gpair/s (session) | max diff / vinf |\n|---|---|---|---|---|---|---|---|" % args.gpus)
    cases = [("Laplace3D_FxU", np.float64, 4096), ("Laplace3D_FxU", np.float64, 20000), ("Laplace3D_FxU", np.float64, 65536), ("Laplace3D_FxU", np.float64, 524288),
             ("Stokes3D_FxU", np.float64, 20000), ("Stokes3D_FxU", np.float64, 131072), ("Stokes3D_DxU", np.float32, 131072)]
    for kernel, dtype, N in cases:
        K = sb.KERNELS[kernel]
        k0, k1, nd = K.SrcDim(), K.TrgDim(), K.NormalDim()
        if N == 524288:  # config 5's geometry: torus nodes, targets == sources, density * weights
            X, Nrm, w, sigma = synth.torus(1024, 512)
            r_src, r_trg, n_src, v_src = X, X, None, (sigma * w)
        else:
            r_src, n_src, v_src, r_trg = synth.cube(N, N, k0, nd, dtype=dtype)
        reps = 3 if N >= 131072 else 20
        dens = [(synth.splitmix_u01(40 + i, N * k0) - 0.5).astype(dtype) for i in range(2)]
        out = np.zeros(N * k1, dtype=dtype)

        def host_apply(i=[0]):
            out[:] = 0
            K.Eval(out, r_trg, r_src, n_src, dens[i[0] % 2], digits=11, dtype=dtype)
            i[0] += 1

        pd = sb.ParticleDirect(dtype, n_gpus=args.gpus)
        pd.SetAccuracy(11)
        pd.AddSrc("S"); pd.AddTrg("T")
        pd.SetKernelS2T("S", "T", K)
        pd.SetSrcCoord("S", r_src, n_src); pd.SetTrgCoord("T", r_trg)
        U = np.zeros(N * k1, dtype=dtype)

        def sess_apply(i=[0]):
            pd.SetSrcDensity("S", dens[i[0] % 2])
            pd.EvalDirect(U, "T", k1)
            i[0] += 1

        host_apply(); sess_apply()
        host_apply(); sess_apply()  # both have now seen density 1 last
        diff = float(np.abs(U - out).max() / np.abs(out).max())
        a, b = median_ms(host_apply, reps), median_ms(sess_apply, reps)
        print("| %s | %s | %d | %.3f | %.3f | %.2fx | %.1f | %.1e |" % (kernel, "f64" if dtype == np.float64 else "f32", N, a, b, a / b, float(N) * N / (b * 1e-3) * 1e-9, diff), flush=True)
        pd.close()


if __name__ == "__main__":
    main()
