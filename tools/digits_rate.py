#!/usr/bin/env python
"""Pair rate and accuracy of the FP64 digits classes (csrc/sctl_cuda.cu rsq_mode_for_digits) on the current GPU:
seed only (digits <= 5), one first-order step (6..11: ParticleFMM's 10, BoundaryIntegralOp's 11), cubic step (>= 12, -1).

    python tools/digits_rate.py      -> markdown on stdout (copied into profiles/)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb
from sctl_b200 import synth
from tools.rate_table import DIMS, F_ALG

# FP64-pipe instructions per pair with the cubic step; the first-order step saves one, the seed five
INSTR = {"Laplace3D_FxU": 12, "Laplace3D_DxU": 17, "Laplace3D_FxdU": 17, "Stokes3D_FxU": 22, "Stokes3D_DxU": 25}


def run(kernel, Ns, Nt, digits, keep=None):
    k0, k1, nd = DIMS[kernel]
    r_src, n_src, v_src, r_trg = synth.cube(Ns, Nt, k0, nd)
    d = [None if a is None else torch.from_numpy(a).cuda() for a in (r_src, n_src, np.abs(v_src), r_trg)]
    v = torch.zeros(Nt * k1, dtype=torch.float64, device="cuda")
    sb.eval_device(kernel, d[0], d[1], d[2], d[3], v, digits=digits)
    torch.cuda.synchronize()
    res = v.clone()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sb.eval_device(kernel, d[0], d[1], d[2], d[3], v, digits=digits)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, res


def main():
    sb.load_library()
    peak = sb.fma_peak(8, 300.0)
    print("device: %s; measured DFMA peak %.2f TFLOP/s\n" % (torch.cuda.get_device_name(0), peak))
    print("| kernel | Ns x Nt | digits | rsqrt class | FP64 instr/pair | ms | Gpair/s | frac of peak (F_alg) | instr-count ceiling | max rel diff vs digits -1 |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for kernel in INSTR:
        for Ns, Nt in ((1 << 20, 1 << 17), (1 << 19, 1 << 19)):
            full = None
            for digits, cls, save in ((-1, "cubic", 0), (11, "first-order", 1), (5, "seed", 5)):
                ms, res = run(kernel, Ns, Nt, digits)
                if full is None:
                    full = res
                diff = float((res - full).abs().max() / full.abs().max())
                rate = Ns * Nt / (ms * 1e-3)
                n_i = INSTR[kernel] - save
                print("| %s | 2^%d x 2^%d | %d | %s | %d | %.2f | %.1f | %.3f | %.3f | %.1e |" % (
                    kernel, int(np.log2(Ns)), int(np.log2(Nt)), digits, cls, n_i, ms, rate * 1e-9, rate * F_ALG[kernel] * 1e-12 / peak, F_ALG[kernel] / (2.0 * n_i), diff))


if __name__ == "__main__":
    main()
