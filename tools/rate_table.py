#!/usr/bin/env python
"""Pair-rate table of sctl_cuda_eval_device / kernel_matrix_device on the current GPU (resident inputs, CUDA events).

    python tools/rate_table.py [--quick] [--paths]      -> markdown table on stdout (copied into profiles/)
"""
import argparse
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb
from sctl_b200 import synth

sb.load_library()
_INFO = {k: sb.kernel_info(k) for k in sb.KERNEL_IDS}
F_ALG = {k: v["f_alg"] for k, v in _INFO.items()}
DIMS = {k: (v["k0"], v["k1"], v["nd"]) for k, v in _INFO.items()}


def time_eval(kernel, dtype, Ns, Nt, reps=3, digits=-1):
    k0, k1, nd = DIMS[kernel]
    ctx = 1.7 if "Helmholtz" in kernel else None  # the wavenumber every fixture uses
    r_src, n_src, v_src, r_trg = synth.cube(Ns, Nt, k0, nd, dtype=dtype)
    d = [None if a is None else torch.from_numpy(a).cuda() for a in (r_src, n_src, v_src, r_trg)]
    v = torch.zeros(Nt * k1, dtype=d[0].dtype, device="cuda")
    sb.eval_device(kernel, d[0], d[1], d[2], d[3], v, digits=digits, ctx=ctx)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sb.eval_device(kernel, d[0], d[1], d[2], d[3], v, digits=digits, ctx=ctx)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, sb.last_plan()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--kernels", default="", help="comma-separated substring filter on kernel names")
    ap.add_argument("--paths", action="store_true", help="compare the Eval paths at shard sizes (Stokes3D_FxU f64)")
    args = ap.parse_args()
    sb.load_library()
    peak64, peak32 = sb.fma_peak(8, 300.0), sb.fma_peak(4, 300.0)
    print("device: %s; measured FMA peak: FP64 %.2f TFLOP/s, FP32 %.2f TFLOP/s" % (torch.cuda.get_device_name(0), peak64, peak32))
    if args.paths:
        print("\n| Nt | Ns | path | ms | Gpair/s | F_alg frac |\n|---|---|---|---|---|---|")
        for lg in (16, 17, 18, 19):
            Nt, Ns = 1 << lg, 1 << 18
            for name, opts in (("const", dict(const_min_targets=1, const_min_launch_flop=0, wide_min_targets=1 << 40)), ("wide(T=4)", dict(const_min_targets=1 << 40, wide_min_targets=1)),
                               ("main(T=2)", dict(const_min_targets=1 << 40, wide_min_targets=1 << 40))):
                for k, val in opts.items():
                    sb.set_option(k, val)
                ms, plan = time_eval("Stokes3D_FxU", np.float64, Ns, Nt)
                rate = Ns * Nt / (ms * 1e-3)
                print("| 2^%d | 2^18 | %s | %.3f | %.1f | %.3f |" % (lg, name, ms, rate * 1e-9, rate * 29e-12 / peak64))
        return
    print("\n| kernel | Real | Ns x Nt | path (grid, launches) | ms | Gpair/s | TFLOP/s (F_alg) | frac of measured peak |\n|---|---|---|---|---|---|---|---|")
    sizes = [(1 << 16, 1 << 16), (1 << 20, 1 << 17)] if args.quick else [(20000, 20000), (1 << 20, 1 << 17), (1 << 19, 1 << 19)]
    for kernel in DIMS:
        if args.kernels and not any(f in kernel for f in args.kernels.split(",")):
            continue
        for dtype in (np.float64, np.float32):
            for Ns, Nt in sizes:
                ms, plan = time_eval(kernel, dtype, Ns, Nt)
                rate = Ns * Nt / (ms * 1e-3)
                tf = rate * F_ALG[kernel] * 1e-12
                peak = peak64 if dtype == np.float64 else peak32
                print("| %s | %s | %d x %d | %dx%d CTAs x %d thr, %d launches | %.3f | %.1f | %.2f | %.3f |" % (
                    kernel, "f64" if dtype == np.float64 else "f32", Ns, Nt, plan["grid_x"], plan["grid_y"], plan["threads"], plan["n_launches"], ms, rate * 1e-9, tf, tf / peak))
    # KernelMatrix: store-bound, report achieved write bandwidth
    print("\n| KernelMatrix kernel | Real | Ns x Nt | ms | GB/s written |\n|---|---|---|---|---|")
    for kernel in ("Laplace3D_FxU", "Stokes3D_FxU", "Stokes3D_DxU"):
        for dtype in (np.float64, np.float32):
            k0, k1, nd = DIMS[kernel]
            Ns, Nt = 8192, 8192
            r_src, n_src, v_src, r_trg = synth.cube(Ns, Nt, k0, nd, dtype=dtype)
            d = [None if a is None else torch.from_numpy(a).cuda() for a in (r_src, n_src, v_src, r_trg)]
            M = torch.empty((Ns * k0, Nt * k1), dtype=d[0].dtype, device="cuda")
            sb.kernel_matrix_device(kernel, d[0], d[1], d[3], M)
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sb.kernel_matrix_device(kernel, d[0], d[1], d[3], M)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print("| %s | %s | %d x %d | %.3f | %.0f |" % (kernel, "f64" if dtype == np.float64 else "f32", Ns, Nt, best, M.numel() * M.element_size() / (best * 1e-3) * 1e-9))


if __name__ == "__main__":
    main()
