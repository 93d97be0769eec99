#!/usr/bin/env python
"""KernelMatrix write bandwidth for whichever libsctl_cuda.so is in place (a gpurun command swaps libraries between two
runs of this tool, like tools/ab_rate.py): one markdown row per (kernel, Real, shape, row pitch) with the GB/s written,
the fraction of the measured HBM copy bandwidth, and a checksum of the matrix so that two libraries can be compared.

    python tools/kmat_rate.py <label>

KMAT_KERNELS=a,b restricts the kernels, KMAT_SHAPES=n keeps the first n shapes, KMAT_REPS=0 launches every configuration
once (for an ncu capture), KMAT_BACK=n times n launches back to back per sample (default 8).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sctl_b200 as sb
from sctl_b200 import synth
from tools.rate_table import DIMS


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except (OSError, KeyError, ValueError):
        return 6400.0, "fallback (B200_PROFILING.md)"


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "shipped"
    sb.load_library()
    peak, src = hbm_peak()
    print("HBM copy bandwidth used as the roof: %.1f GB/s (%s)\n" % (peak, src))
    print("| library | kernel | Real | Ns x Nt | row pitch | ms per launch (8 back to back, best of 5) | GB/s written | frac of HBM copy peak | checksum |\n|---|---|---|---|---|---|---|---|---|")
    shapes = [(8192, 8192, 0), (8192, 8191, 0), (2048, 65536, 0), (512, 512, 0), (64, 200, 0)][: int(os.environ.get("KMAT_SHAPES", "99"))]
    reps = int(os.environ.get("KMAT_REPS", "5"))
    back = int(os.environ.get("KMAT_BACK", "8"))
    for kernel in os.environ.get("KMAT_KERNELS", "Laplace3D_FxU,Laplace3D_DxU,Stokes3D_FxU,Stokes3D_DxU,Stokes3D_FxT,Helmholtz3D_FxU").split(","):
        k0, k1, nd = DIMS[kernel]
        for dtype in (np.float64, np.float32):
            for Ns, Nt, pad in shapes:
                if k0 * k1 >= 9 and Ns * Nt > 8192 * 8192:
                    continue
                if k0 * k1 > 9 and Ns * Nt >= 8192 * 8191:
                    Ns //= 2
                r_src, n_src, v_src, r_trg = synth.cube(Ns, Nt, k0, nd, dtype=dtype)
                d = [None if a is None else torch.from_numpy(a).cuda() for a in (r_src, n_src, v_src, r_trg)]
                ld = Nt * k1 + pad
                M = torch.empty((Ns * k0, ld), dtype=d[0].dtype, device="cuda")
                ctx = 1.7 if "Helmholtz" in kernel else None
                sb.kernel_matrix_device(kernel, d[0], d[1], d[3], M, ldM=ld, ctx=ctx)
                torch.cuda.synchronize()
                chk = float(M[:, : Nt * k1].double().sum())
                best = 1e30
                for _ in range(reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(back):  # back to back: a 60 us kernel timed alone carries ~10 % of launch latency
                        sb.kernel_matrix_device(kernel, d[0], d[1], d[3], M, ldM=ld, ctx=ctx)
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / back)
                gbs = Ns * k0 * Nt * k1 * M.element_size() / (best * 1e-3) * 1e-9
                print("| %s | %s | %s | %d x %d | %d | %.4f | %.0f | %.3f | %.15g |" % (
                    label, kernel, "f64" if dtype == np.float64 else "f32", Ns, Nt, ld, best, gbs, gbs / peak, chk), flush=True)
                del M


if __name__ == "__main__":
    main()
