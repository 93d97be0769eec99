#!/usr/bin/env python
"""Host-pointer call latency of sctl_cuda_eval / sctl_cuda_kernel_matrix for small problems (drop-in view)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb
from sctl_b200 import synth

sb.load_library()
print("| call | Ns x Nt | wall us/call (median of 50) | of which H2D / kernels / D2H (CUDA events, us) |\n|---|---|---|---|")
for kernel in ("Laplace3D_FxU", "Stokes3D_FxU"):
    K = sb.KERNELS[kernel]
    k0, k1, nd = K.SrcDim(), K.TrgDim(), K.NormalDim()
    for n in (16, 128, 1024, 4096, 20000):
        r_src, n_src, v_src, r_trg = synth.cube(n, n, k0, nd)
        out = np.zeros(n * k1)
        for _ in range(5):
            K.Eval(out, r_trg, r_src, n_src, v_src)
        ts = []
        for _ in range(50):
            t0 = time.perf_counter()
            K.Eval(out, r_trg, r_src, n_src, v_src)
            ts.append(time.perf_counter() - t0)
        t = sb.last_timing()
        print("| %s Eval | %d x %d | %.1f | %.1f / %.1f / %.1f |" % (kernel, n, n, np.median(ts) * 1e6, t["h2d_ms"] * 1e3, t["kernel_ms"] * 1e3, t["d2h_ms"] * 1e3))
    for n in (16, 128, 512):
        r_src, n_src, v_src, r_trg = synth.cube(n, n, k0, nd)
        M = None
        for _ in range(5):
            M = K.KernelMatrix(r_trg, r_src, n_src, M=M)
        ts = []
        for _ in range(50):
            t0 = time.perf_counter()
            M = K.KernelMatrix(r_trg, r_src, n_src, M=M)
            ts.append(time.perf_counter() - t0)
        print("| %s KernelMatrix | %d x %d | %.1f | |" % (kernel, n, n, np.median(ts) * 1e6))
