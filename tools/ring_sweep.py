#!/usr/bin/env python
"""Shared-memory ring path at small/medium sizes: shortest source chunk per CTA (ring_min_tiles) and path thresholds."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb
from sctl_b200 import synth


def run(kernel, dtype, Ns, Nt, reps=5):
    info = sb.kernel_info(kernel)
    r_src, n_src, v_src, r_trg = synth.cube(Ns, Nt, info["k0"], info["nd"], dtype=dtype)
    d = [None if a is None else torch.from_numpy(a).cuda() for a in (r_src, n_src, v_src, r_trg)]
    v = torch.zeros(Nt * info["k1"], dtype=d[0].dtype, device="cuda")
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
    return best, sb.last_plan()


def main():
    sb.load_library()
    print("| kernel | Real | Ns x Nt | config | grid | ms | Gpair/s |\n|---|---|---|---|---|---|---|")
    for kernel, dtype in (("Laplace3D_FxU", np.float64), ("Stokes3D_FxU", np.float64), ("Stokes3D_FxU", np.float32)):
        for Ns, Nt in ((20000, 20000), (10000, 10000), (5000, 5000), (2500, 2500), (30000, 30000), (65536, 8192), (1000, 30000)):
            for name, opts in (("min_tiles=8 (old)", dict(ring_min_tiles=8)), ("min_tiles=2", dict(ring_min_tiles=2)), ("min_tiles=1", dict(ring_min_tiles=1)),
                               ("min_tiles=1 split<=256", dict(ring_min_tiles=1, ring_max_split=256)), ("min_tiles=2 wide from 2^13", dict(ring_min_tiles=2, wide_min_targets=1 << 13)),
                               ("min_tiles=1 split<=256 wide from 2^13", dict(ring_min_tiles=1, ring_max_split=256, wide_min_targets=1 << 13))):
                sb.reset_options()
                for k, val in opts.items():
                    sb.set_option(k, val)
                ms, plan = run(kernel, dtype, Ns, Nt)
                print("| %s | %s | %d x %d | %s | %dx%d x %d thr, %d launches | %.3f | %.1f |" % (kernel, "f64" if dtype == np.float64 else "f32", Ns, Nt, name, plan["grid_x"], plan["grid_y"],
                                                                                            plan["threads"], plan["n_launches"], ms, Ns * Nt / (ms * 1e-3) * 1e-9), flush=True)
    sb.reset_options()


if __name__ == "__main__":
    main()
