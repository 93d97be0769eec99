#!/usr/bin/env python
"""FP32 Eval: shared-memory ring vs the constant-bank path (overlapping launches), by kernel and target count."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb
from sctl_b200 import synth


def run(kernel, Ns, Nt, reps=3):
    info = sb.kernel_info(kernel)
    r_src, n_src, v_src, r_trg = synth.cube(Ns, Nt, info["k0"], info["nd"], dtype=np.float32)
    d = [None if a is None else torch.from_numpy(a).cuda() for a in (r_src, n_src, v_src, r_trg)]
    v = torch.zeros(Nt * info["k1"], dtype=torch.float32, device="cuda")
    sb.eval_device(kernel, d[0], d[1], d[2], d[3], v)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        v.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sb.eval_device(kernel, d[0], d[1], d[2], d[3], v)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, v, sb.last_plan()


def main():
    sb.load_library()
    peak32 = sb.fma_peak(4, 300.0)
    print("device: %s; measured FP32 FMA peak %.2f TFLOP/s" % (torch.cuda.get_device_name(0), peak32))
    print("\n| kernel | Nt | Ns | path | grid | launches | ms | Gpair/s | F_alg frac | diff vs smem / vinf |\n|---|---|---|---|---|---|---|---|---|---|")
    for kernel in ("Stokes3D_FxU", "Laplace3D_FxU", "Stokes3D_DxU", "Laplace3D_FxdU"):
        f_alg = sb.kernel_info(kernel)["f_alg"]
        for lg in (15, 17, 20):
            Nt, Ns = 1 << lg, 1 << 18
            base = None
            for name, opts in (("smem", dict(const_fp32=0)), ("const banks=2", dict(const_fp32=1, const_banks_fp32=2)), ("const banks=4", dict(const_fp32=1, const_banks_fp32=4))):
                sb.reset_options()
                for k, val in opts.items():
                    sb.set_option(k, val)
                ms, v, plan = run(kernel, Ns, Nt)
                if base is None:
                    base = v
                rate = Ns * Nt / (ms * 1e-3)
                print("| %s | 2^%d | 2^18 | %s | %dx%d x %d thr | %d | %.3f | %.1f | %.3f | %.1e |" % (
                    kernel, lg, name, plan["grid_x"], plan["grid_y"], plan["threads"], plan["n_launches"], ms, rate * 1e-9, rate * f_alg * 1e-12 / peak32,
                    float((v - base).abs().max() / base.abs().max())), flush=True)
    sb.reset_options()


if __name__ == "__main__":
    main()
