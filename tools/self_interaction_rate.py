#!/usr/bin/env python
"""targets == sources (one r2 == 0 pair per target): Eval rate with the per-(chunk, target block) coincidence bits, with
the global flag only, and with the probe off (always masked); Stokes3D_FxU FP64, cube points and torus nodes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb
from sctl_b200 import synth


def run(kernel, r, v_src, reps=3):
    d_r, d_v = torch.from_numpy(r).cuda(), torch.from_numpy(v_src).cuda()
    N = r.size // 3
    out = torch.zeros(N * 3, dtype=torch.float64, device="cuda")
    sb.eval_device(kernel, d_r, None, d_v, d_r, out)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        out.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sb.eval_device(kernel, d_r, None, d_v, d_r, out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def main():
    sb.load_library()
    peak = sb.fma_peak(8, 300.0)
    print("measured FP64 FMA peak %.2f TFLOP/s\n" % peak)
    print("| points | N (targets == sources) | probe | ms | Gpair/s | F_alg frac | identical to always-masked |\n|---|---|---|---|---|---|---|")
    for name, lg in (("cube", 19), ("torus", 20)):
        N = 1 << lg
        r = synth.torus(1 << ((lg + 1) // 2), 1 << (lg // 2))[0] if name == "torus" else synth.cube(N, 0, 3, 0)[0]
        v = synth.splitmix_u01(4, N * 3) - 0.5
        base = None
        for label, opts in (("off (always masked)", dict(probe_min_pairs=0)), ("global flag only", dict(probe_per_block=0)), ("per (chunk, block) bits", dict())):
            sb.reset_options()
            for k, val in opts.items():
                sb.set_option(k, val)
            ms, out = run("Stokes3D_FxU", r, v)
            if base is None:
                base = out
            rate = float(N) * N / (ms * 1e-3)
            print("| %s | 2^%d | %s | %.2f | %.1f | %.3f | %s |" % (name, lg, label, ms, rate * 1e-9, rate * 29e-12 / peak, bool(torch.equal(out, base))), flush=True)
    sb.reset_options()


if __name__ == "__main__":
    main()
