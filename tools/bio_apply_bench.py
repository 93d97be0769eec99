#!/usr/bin/env python
"""Time one BoundaryIntegralOp apply (far field + near correction) on a torus discretisation, on 1..n GPUs of one box:
the device-side pipeline (sctl_cuda_bio_apply) against the same apply staged through the two sessions' host-array entries.

Geometry: Nu x Nv torus nodes, elements = 8 x 8-node patches; every element's near targets are the nodes of its
3 x 3 patch neighbourhood (576 near targets per 64-node element -- the near-list size BoundaryIntegralOp builds on this
surface at dist_far ~ 0.08, profiles/r1_bio_apply_sessions_vs_reference.txt: 1.25 M near pairs at 131 072 nodes), cut into
blocks of 256 targets.  K_near holds random values: the timings do not depend on them.

    python tools/bio_apply_bench.py [--gpus 1,2,4,8] [--sizes 512x256,1024x512]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sctl_b200 as sb  # noqa: E402
from sctl_b200 import synth  # noqa: E402


def near_structure(Nu, Nv, k0, k1, blk=256):
    pu, pv = Nu // 8, Nv // 8
    n_elem = pu * pv
    node = np.arange(Nu * Nv).reshape(Nu, Nv)
    # node order: element-major, so that every element owns a contiguous node range (the reference's layout)
    elem_nodes = node.reshape(pu, 8, pv, 8).transpose(0, 2, 1, 3).reshape(n_elem, 64)
    order = elem_nodes.reshape(-1)
    pos = np.empty_like(order)
    pos[order] = np.arange(order.size)  # torus grid index -> position in the element-major ordering
    cols = [[] for _ in range(6)]
    owners = []
    k_off = 0
    u_off = 0
    for e in range(n_elem):
        iu, iv = divmod(e, pv)
        nb = [((iu + du) % pu) * pv + (iv + dv) % pv for du in (-1, 0, 1) for dv in (-1, 0, 1)]
        trg = pos[elem_nodes[nb].reshape(-1)]
        owners.append(trg)
        for t0 in range(0, trg.size, blk):
            n = min(blk, trg.size - t0)
            for c, v in zip(cols, (e * 64, e * 64, 64, u_off + t0, n, k_off + t0 * k1 * 64 * k0)):
                c.append(v)
        k_off += trg.size * k1 * 64 * k0
        u_off += trg.size
    owner = np.concatenate(owners)
    scatter = np.argsort(owner, kind="stable").astype(np.int64)
    cnt = np.bincount(owner, minlength=Nu * Nv).astype(np.int64)
    dsp = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int64)
    return [np.array(c, dtype=np.int64) for c in cols], order, cnt, dsp, scatter, k_off


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--sizes", default="512x256")
    ap.add_argument("--kernels", default="Laplace3D_FxU,Stokes3D_FxU")
    ap.add_argument("--digits", type=int, default=11)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    sb.load_library()
    print("| kernel | nodes | near pairs | K_near GB | GPUs | pipeline ms/apply (host clock) | device: far (+near GEMV) / tail / total ms | staged ms/apply | far-field Gpair/s | max rel diff |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for size in args.sizes.split(","):
        Nu, Nv = [int(x) for x in size.split("x")]
        N = Nu * Nv
        Xg, Nrm, W, _ = synth.torus(Nu, Nv)
        for kernel in args.kernels.split(","):
            K = sb.KERNELS[kernel]
            k0, k1 = K.SrcDim(), K.TrgDim()
            tab, order, cnt, dsp, scatter, K_len = near_structure(Nu, Nv, k0, k1)
            X = Xg.reshape(-1, 3)[order].reshape(-1)
            Xn = Nrm.reshape(-1, 3)[order].reshape(-1)
            wts = W[order]
            rng = np.random.RandomState(1)
            K_near = rng.standard_normal(K_len) * 1e-4
            F = rng.standard_normal(N * k0)
            ref = None
            for n in [int(x) for x in args.gpus.split(",")]:
                if n > sb.device_count():
                    continue
                near = sb.NearBlocks(K, *tab)
                near.set_matrix(K_near)
                near.set_targets(cnt, dsp, scatter)
                bio = sb.BioApply(K, X, Xn, wts, X, None, digits=args.digits, n_gpus=n, near=near)
                U = bio.compute_potential(F)
                best = 1e30
                for _ in range(args.reps):
                    t0 = time.perf_counter()
                    U = bio.compute_potential(F, U=U)
                    best = min(best, time.perf_counter() - t0)
                tm = bio.last_timing()
                staged = 1e30
                for _ in range(max(2, args.reps // 2)):
                    t0 = time.perf_counter()
                    Us = bio.compute_potential_staged(F)
                    staged = min(staged, time.perf_counter() - t0)
                if ref is None:
                    ref = U.copy()
                diff = max(float(np.abs(U - ref).max() / np.abs(ref).max()), float(np.abs(U - Us).max() / np.abs(ref).max()))
                print("| %s | %d | %d | %.2f | %d | %.2f | %.2f / %.2f / %.2f | %.2f | %.0f | %.1e |" % (
                    kernel, N, scatter.size, K_len * 8e-9, n, best * 1e3, tm["far_ms"], tm["tail_ms"], tm["total_ms"], staged * 1e3, N * float(N) / (tm["far_ms"] * 1e-3) * 1e-9, diff))
                sys.stdout.flush()
                bio.close()
                near.close()


if __name__ == "__main__":
    main()
