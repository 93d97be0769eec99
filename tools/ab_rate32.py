#!/usr/bin/env python
"""FP32 Eval rate by path (shared-memory ring / constant bank) and target count for whichever libsctl_cuda.so is in
place; rows carry a checksum so that two builds (tools/ab_build.sh, e.g. packed vs scalar FP32 loops) can be compared
bit for bit.

    python tools/ab_rate32.py <label>
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb
from sctl_b200 import synth


def run(kernel, Ns, Nt, digits, reps=4):
    info = sb.kernel_info(kernel)
    r_src, n_src, v_src, r_trg = synth.cube(Ns, Nt, info["k0"], info["nd"], dtype=np.float32)
    d = [None if a is None else torch.from_numpy(a).cuda() for a in (r_src, n_src, v_src, r_trg)]
    v = torch.zeros(Nt * info["k1"], dtype=torch.float32, device="cuda")
    sb.eval_device(kernel, d[0], d[1], d[2], d[3], v, digits=digits)
    torch.cuda.synchronize()
    chk = float(v.double().sum())
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sb.eval_device(kernel, d[0], d[1], d[2], d[3], v, digits=digits)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, chk, sb.last_plan()


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "shipped"
    sb.load_library()
    peak32 = sb.fma_peak(4, 300.0)
    for kernel in ("Stokes3D_FxU", "Laplace3D_FxU", "Stokes3D_DxU", "Laplace3D_FxdU"):
        f_alg = sb.kernel_info(kernel)["f_alg"]
        for lg, digits in ((13, -1), (15, -1), (17, -1), (20, -1), (20, 7)):
            Nt, Ns = 1 << lg, 1 << 18
            for name, opts in (("ring", dict(const_fp32=0)), ("const", dict(const_fp32=1))):
                if name == "const" and lg < 15:
                    continue
                sb.reset_options()
                for k, val in opts.items():
                    sb.set_option(k, val)
                ms, chk, plan = run(kernel, Ns, Nt, digits)
                rate = Ns * Nt / (ms * 1e-3)
                print("| %s | %s | 2^%d | %d | %s | %dx%d x %d thr | %.3f | %.1f | %.3f | %.17g | %.2f |" % (
                    label, kernel, lg, digits, name, plan["grid_x"], plan["grid_y"], plan["threads"], ms, rate * 1e-9, rate * f_alg * 1e-12 / peak32, chk, peak32), flush=True)
    sb.reset_options()


if __name__ == "__main__":
    main()
