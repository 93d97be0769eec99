#!/usr/bin/env python
"""Constant-bank path: banks in flight x cluster policy, by target count (Stokes3D_FxU f64, resident inputs).

    python tools/bank_sweep.py [--kernel K] -> markdown table on stdout (copied into profiles/)
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb
from sctl_b200 import synth


def run(kernel, Ns, Nt, reps=3):
    info = sb.kernel_info(kernel)
    r_src, n_src, v_src, r_trg = synth.cube(Ns, Nt, info["k0"], info["nd"])
    d = [None if a is None else torch.from_numpy(a).cuda() for a in (r_src, n_src, v_src, r_trg)]
    v = torch.zeros(Nt * info["k1"], dtype=torch.float64, device="cuda")
    sb.eval_device(kernel, d[0], d[1], d[2], d[3], v)
    torch.cuda.synchronize()
    first = v.clone()
    best = 1e30
    for _ in range(reps):
        v.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sb.eval_device(kernel, d[0], d[1], d[2], d[3], v)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    assert torch.equal(first, v), "run-to-run results differ"
    return best, v, sb.last_plan()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", default="Stokes3D_FxU")
    ap.add_argument("--sizes", default="16,17,18,20")
    ap.add_argument("--ns", type=int, default=18, help="log2 of the source count (0: Ns = Nt)")
    ap.add_argument("--all-banks", action="store_true")
    args = ap.parse_args()
    sb.load_library()
    peak64 = sb.fma_peak(8, 300.0)
    f_alg = sb.kernel_info(args.kernel)["f_alg"]
    print("device: %s; measured FP64 FMA peak %.2f TFLOP/s; kernel %s" % (torch.cuda.get_device_name(0), peak64, args.kernel))
    print("\n| Nt | Ns | path | grid | launches | ms | Gpair/s | F_alg frac | max diff vs 1 bank / vinf |\n|---|---|---|---|---|---|---|---|---|")
    for lg in [int(x) for x in args.sizes.split(",")]:
        Nt = 1 << lg if lg < 64 else lg  # small numbers are exponents, large ones literal target counts
        Ns = 1 << args.ns if args.ns else Nt
        base = None
        cfgs = [("const banks=1 cta/sm=6", dict(const_min_targets=1, const_min_launch_flop=0, const_banks=1, const_ctas_per_sm=6)),
                ("const banks=2 ov=1", dict(const_min_targets=1, const_min_launch_flop=0, const_banks=2, const_ctas_per_sm_overlap=1)),
                ("const banks=2 ov=2", dict(const_min_targets=1, const_min_launch_flop=0, const_banks=2, const_ctas_per_sm_overlap=2)),
                ("const banks=4 ov=1", dict(const_min_targets=1, const_min_launch_flop=0, const_banks=4, const_ctas_per_sm_overlap=1)),
                ("smem wide(T=4)", dict(const_min_targets=1 << 40, wide_min_targets=1, few_max_targets=0)),
                ("smem main(T=2)", dict(const_min_targets=1 << 40, wide_min_targets=1 << 40, few_max_targets=0)),
                ("smem few", dict(const_min_targets=1 << 40, wide_min_targets=1 << 40, few_max_targets=1 << 40))]
        if args.all_banks:
            cfgs[3:3] = [("const banks=2 ov=4", dict(const_min_targets=1, const_min_launch_flop=0, const_banks=2, const_ctas_per_sm_overlap=4)),
                         ("const banks=3 ov=1", dict(const_min_targets=1, const_min_launch_flop=0, const_banks=3, const_ctas_per_sm_overlap=1)),
                         ("const banks=4 ov=2", dict(const_min_targets=1, const_min_launch_flop=0, const_banks=4, const_ctas_per_sm_overlap=2))]
        for name, opts in cfgs:
            for k, val in opts.items():
                sb.set_option(k, val)
            ms, v, plan = run(args.kernel, Ns, Nt)
            if base is None:
                base = v
            diff = float((v - base).abs().max() / base.abs().max())
            rate = Ns * Nt / (ms * 1e-3)
            sb.reset_options()
            print("| %d | %d | %s | %dx%d x %d thr | %d | %.3f | %.1f | %.3f | %.1e |" % (
                Nt, Ns, name, plan["grid_x"], plan["grid_y"], plan["threads"], plan["n_launches"], ms, rate * 1e-9, rate * f_alg * 1e-12 / peak64, diff), flush=True)


if __name__ == "__main__":
    main()
