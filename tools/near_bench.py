#!/usr/bin/env python
"""Batched near-field blocks (sctl_cuda_near_*): device time of the subtract and apply phases on a large synthetic
block list against their HBM rooflines, plus what one apply costs end to end (F up, U down).

    python tools/near_bench.py      -> markdown table on stdout (copied into profiles/)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb
from sctl_b200 import synth


def block_list(n_elem, nodes, near, k0, k1, blk=256):
    nds_dsp = np.arange(n_elem, dtype=np.int64) * nodes
    near_dsp = np.arange(n_elem, dtype=np.int64) * near
    cols = [[] for _ in range(6)]
    for t0 in range(0, near, blk):
        n = min(blk, near - t0)
        cols[0].append(nds_dsp); cols[1].append(nds_dsp); cols[2].append(np.full(n_elem, nodes, dtype=np.int64))
        cols[3].append(near_dsp + t0); cols[4].append(np.full(n_elem, n, dtype=np.int64))
        cols[5].append(nds_dsp * near * k0 * k1 + t0 * k1 * nodes * k0)
    return [np.concatenate(c) for c in cols]


def main():
    sb.load_library()
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    hbm = peaks["hbm_gbs"]
    print("HBM copy peak (MEASURED_PEAKS.json): %.1f GB/s\n" % hbm)
    print("| kernel | Real | elements x nodes x near targets | K bytes | subtract kernel ms | GB/s (read+write of K) | frac of HBM peak | apply kernel ms | GB/s (read of K) | frac | apply end to end ms (F up + kernel + U down) |\n|---|---|---|---|---|---|---|---|---|---|---|")
    for kernel, dtype, n_elem, nodes, near in (("Laplace3D_FxU", np.float64, 8192, 64, 192), ("Stokes3D_FxU", np.float64, 4096, 64, 64), ("Laplace3D_DxU", np.float32, 8192, 64, 192)):
        K = sb.KERNELS[kernel]
        k0, k1, nd = K.SrcDim(), K.TrgDim(), K.NormalDim()
        tab = block_list(n_elem, nodes, near, k0, k1)
        nb = sb.NearBlocks(K, *tab, dtype=dtype)
        Xs, Xn, F, Xt = synth.cube(n_elem * nodes, n_elem * near, k0, nd, dtype=dtype)
        wts = np.full(n_elem * nodes, 1e-3, dtype=dtype)
        nb.set_matrix(np.zeros(nb.K_len, dtype=dtype))
        kb = nb.K_len * np.dtype(dtype).itemsize
        sub = []
        for _ in range(3):
            nb.subtract_direct(Xs, Xn, wts, Xt)
            sub.append(sb.last_timing()["kernel_ms"])
        U = np.zeros(nb.nU, dtype=dtype)
        app, wall = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            nb.apply(F, U)
            wall.append((time.perf_counter() - t0) * 1e3)
            app.append(sb.last_timing()["kernel_ms"])
        s_ms, a_ms = min(sub), min(app)
        print("| %s | %s | %d x %d x %d | %.2f GB | %.3f | %.0f | %.2f | %.3f | %.0f | %.2f | %.2f |" % (
            kernel, "f64" if dtype == np.float64 else "f32", n_elem, nodes, near, kb * 1e-9, s_ms, 2 * kb / (s_ms * 1e-3) * 1e-9, 2 * kb / (s_ms * 1e-3) * 1e-9 / hbm,
            a_ms, kb / (a_ms * 1e-3) * 1e-9, kb / (a_ms * 1e-3) * 1e-9 / hbm, float(np.median(wall))), flush=True)
        nb.close()


if __name__ == "__main__":
    main()
