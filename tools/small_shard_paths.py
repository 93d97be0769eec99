#!/usr/bin/env python
"""FP64 Eval at shard-sized target counts (2^13 .. 2^15 targets against 2^17 sources, the per-GPU far field of a
131 072-node surface on 4 .. 16 GPUs): shared-memory ring vs constant-bank path, by kernel and digits class."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb
from sctl_b200 import synth


def run(kernel, Ns, Nt, digits, reps=5):
    info = sb.kernel_info(kernel)
    r_src, n_src, v_src, r_trg = synth.cube(Ns, Nt, info["k0"], info["nd"])
    d = [None if a is None else torch.from_numpy(a).cuda() for a in (r_src, n_src, v_src, r_trg)]
    v = torch.zeros(Nt * info["k1"], dtype=torch.float64, device="cuda")
    sb.eval_device(kernel, d[0], d[1], d[2], d[3], v, digits=digits)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sb.eval_device(kernel, d[0], d[1], d[2], d[3], v, digits=digits)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, sb.last_plan()


def main():
    sb.load_library()
    print("| kernel | digits | Nt | Ns | path | grid | launches | ms | Gpair/s |\n|---|---|---|---|---|---|---|---|---|")
    kernels = sys.argv[1].split(",") if len(sys.argv) > 1 else ["Laplace3D_FxU", "Stokes3D_FxU"]
    sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [13, 14, 15]
    for kernel in kernels:
        for digits in (11, -1):
            for lg in sizes:
                Nt, Ns = 1 << lg, 1 << 17
                for name, opts in (("default", {}), ("ring", dict(const_min_targets=1 << 40)), ("const 2 slices", dict(const_min_targets=1, const_min_launch_flop=0)), ("const 4 slices", dict(const_min_targets=1, const_min_launch_flop=0, const_banks=4))):
                    sb.reset_options()
                    for k, val in opts.items():
                        sb.set_option(k, val)
                    ms, plan = run(kernel, Ns, Nt, digits)
                    print("| %s | %d | 2^%d | 2^17 | %s | %dx%d x %d thr | %d | %.3f | %.1f |" % (
                        kernel, digits, lg, name, plan["grid_x"], plan["grid_y"], plan["threads"], plan["n_launches"], ms, Ns * Nt / (ms * 1e-3) * 1e-9), flush=True)
    sb.reset_options()


if __name__ == "__main__":
    main()
