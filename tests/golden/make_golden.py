#!/usr/bin/env python
"""Regenerates tests/golden/golden_eval.npz by running the UNMODIFIED reference (oracle/_ref, built from
/root/reference/include by oracle/Makefile) on the cases in cases.py.  Run in the build container:

    python tests/golden/make_golden.py

The .npz holds outputs only (v_trg / M / rsqrt samples); inputs are regenerated from SplitMix64 seeds."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cases  # noqa: E402
import oracle  # noqa: E402


def main():
    ref = oracle.RefOracle()
    out = {"meta/isa": np.array(ref.isa), "meta/veclen": np.array([ref.veclen(8), ref.veclen(4)])}
    for kernel in cases.KERNELS:
        for dtype in (np.float64, np.float32):
            for Ns, Nt, variant, digits in cases.EVAL_CASES:
                r_src, n_src, v_src, r_trg, v0 = cases.inputs(kernel, Ns, Nt, variant, dtype)
                out[cases.key(kernel, dtype, Ns, Nt, variant, digits)] = ref.eval(kernel, r_src, n_src, v_src, r_trg, v_trg=v0, dtype=dtype, omp=True, digits=digits, ctx=cases.CTX[kernel] or 0.0)
            for Ns, Nt in cases.KMAT_CASES:
                r_src, n_src, v_src, r_trg, _ = cases.inputs(kernel, Ns, Nt, "cube", dtype)
                out[cases.kmat_key(kernel, dtype, Ns, Nt)] = ref.kernel_matrix(kernel, r_src, n_src, r_trg, dtype=dtype, omp=False, ctx=cases.CTX[kernel] or 0.0)
    # approx_rsqrt<digits> samples (vec-test.hpp:431-465 style): x over 40 decades, plus masked lanes
    x = np.concatenate([10.0 ** np.linspace(-20, 20, 397), [0.0, 1.0, 4.0, 0.25]])
    out["rsqrt/x"] = x
    for digits in (-1, 4, 7, 10):
        out["rsqrt/f64/%d" % digits] = ref.rsqrt(x, digits)
    for digits in (-1, 4, 7):
        out["rsqrt/f32/%d" % digits] = ref.rsqrt(x.astype(np.float32), digits)
    # SCTL binary container fixtures written by the reference's own Vector::Write / Matrix::Write
    import ctypes
    here = os.path.dirname(os.path.abspath(__file__))
    vec = np.array([1.5, -2.25, 3.0, 0.0, 1e-300, -7.0, 42.0])
    mat = np.arange(6, dtype=np.float64).reshape(2, 3) - 2.5
    ref.lib.sctl_ref_vector_write.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p]
    ref.lib.sctl_ref_matrix_write.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
    ref.lib.sctl_ref_vector_write(os.path.join(here, "sctl_vector_f64.bin").encode(), vec.size, vec.ctypes.data_as(ctypes.c_void_p))
    ref.lib.sctl_ref_matrix_write(os.path.join(here, "sctl_matrix_f64.bin").encode(), 2, 3, mat.ctypes.data_as(ctypes.c_void_p))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_eval.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%d arrays, %.1f KB" % (len(out), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
