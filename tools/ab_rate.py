#!/usr/bin/env python
"""Pair rate of the FP64 constant-bank path for whichever libsctl_cuda.so is in place (tools/ab_build.sh builds tuning
variants, a gpurun command swaps them in between runs of this tool): one markdown row per (kernel, digits).

    python tools/ab_rate.py <label> [log2n]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb
from sctl_b200 import synth
from tools.rate_table import DIMS, F_ALG


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "shipped"
    n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 19)
    sb.load_library()
    peak = sb.fma_peak(8, 300.0)
    kernels = os.environ.get("AB_KERNELS", "Stokes3D_FxU,Laplace3D_FxU,Stokes3D_DxU").split(",")
    for kernel in kernels:
        k0, k1, nd = DIMS[kernel]
        r_src, n_src, v_src, r_trg = synth.cube(n, n, k0, nd)
        d = [None if a is None else torch.from_numpy(a).cuda() for a in (r_src, n_src, v_src, r_trg)]
        for digits in (-1, 11):
            v = torch.zeros(n * k1, dtype=torch.float64, device="cuda")
            ctx = 1.7 if "Helmholtz" in kernel else None
            sb.eval_device(kernel, d[0], d[1], d[2], d[3], v, digits=digits, ctx=ctx)
            torch.cuda.synchronize()
            chk = float(v.double().sum())
            ts = []
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sb.eval_device(kernel, d[0], d[1], d[2], d[3], v, digits=digits, ctx=ctx)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            rate = n * n / (min(ts) * 1e-3)
            print("| %s | %s | %d | %.2f | %.2f | %.1f | %.4f | %.17g | %.2f |" % (
                label, kernel, digits, min(ts), float(np.median(ts)), rate * 1e-9, rate * F_ALG[kernel] * 1e-12 / peak, chk, peak), flush=True)


if __name__ == "__main__":
    main()
