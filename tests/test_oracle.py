"""CPU tests: pin the oracle (oracle/sctl_oracle.c) against the reference's own fixtures and golden vectors."""
import numpy as np
import pytest

import cases
from conftest import TOL, rel_err

PI = np.pi
KERNELS = list(cases.KERNELS)


# ---- fixtures held by the reference's own tests ---------------------------------------------------------------

def test_dims_names_scale_flops(port):
    # test-generic-kernel.cpp:40-53, test-kernel_functions.cpp:39-43,55-59,73-77,97-101,121-125
    expect = {
        "Laplace3D_FxU": (1, 1, 0, 6, "Laplace3D-FxU", 1 / (4 * PI)),
        "Laplace3D_DxU": (1, 1, 3, 14, "Laplace3D-DxU", 1 / (4 * PI)),
        "Laplace3D_FxdU": (1, 3, 0, 11, "Laplace3D-FxdU", -1 / (4 * PI)),
        "Stokes3D_FxU": (3, 3, 0, 23, "Stokes3D-FxU", 1 / (8 * PI)),
        "Stokes3D_DxU": (3, 3, 3, 26, "Stokes3D-DxU", 3 / (4 * PI)),
    }
    for k, (k0, k1, nd, fl, name, sc) in expect.items():
        d = port.dims(k)
        assert (d["k0"], d["k1"], d["nd"], d["flops"], d["name"]) == (k0, k1, nd, fl, name)
        assert abs(d["scale64"] - sc) < 1e-15 and abs(d["scale32"] - sc) < 1e-7


def test_generic_kernel_fixtures(port):
    sf = 1 / (4 * PI)
    # 1x1 KernelMatrix = 1/(4 pi)  (test-generic-kernel.cpp:57-66)
    M = port.kernel_matrix("Laplace3D_FxU", [0, 0, 0], None, [1, 0, 0])
    assert M.shape == (1, 1) and abs(M[0, 0] - sf) < 1e-9 * sf
    # single source, targets at r = 1,2,3,5  (:69-87)
    v = port.eval("Laplace3D_FxU", [0, 0, 0], None, [1.0], [1, 0, 0, 2, 0, 0, 3, 0, 0, 5, 0, 0])
    np.testing.assert_allclose(v, sf / np.array([1, 2, 3, 5.0]), rtol=1e-9)
    # 2-source linearity (:90-103)
    v = port.eval("Laplace3D_FxU", [0, 0, 0, 4, 0, 0], None, [1.0, 1.0], [2, 0, 0])
    assert abs(v[0] - 2 * sf / 2) < 1e-9


def test_kernel_function_fixtures(port):
    # test-kernel_functions.cpp:45-48, 63-66, 81-90, 106-114
    assert abs(port.eval("Laplace3D_FxU", [0, 0, 0], None, [3.0], [2, 0, 0])[0] - 3 / (8 * PI)) < 1e-9
    assert abs(port.eval("Laplace3D_DxU", [0, 0, 0], [1, 0, 0], [1.0], [2, 0, 0])[0] - 1 / (16 * PI)) < 1e-9
    v = port.eval("Laplace3D_FxdU", [0, 0, 0], None, [1.0], [3, 0, 0])
    assert abs(np.linalg.norm(v) - 1 / (36 * PI)) < 1e-9 and abs(v[1]) < 1e-12 and abs(v[2]) < 1e-12
    v = port.eval("Stokes3D_FxU", [0, 0, 0], None, [1.0, 0, 0], [2, 0, 0])
    assert abs(v[0] - 1 / (8 * PI)) < 1e-9 and abs(v[1]) < 1e-12 and abs(v[2]) < 1e-12


@pytest.mark.parametrize("kernel", KERNELS)
def test_eval_equals_kernel_matrix_apply(port, kernel):
    # test-kernel_functions.cpp:488-526 (7 sources x 5 targets, deterministic inputs, tol 1e-10)
    k0, k1, nd = cases.KERNELS[kernel]
    Ns, Nt = 7, 5
    Xs = 0.1 * (np.arange(Ns * 3) + 1) - 0.35
    Vs = 0.3 * (np.arange(Ns * k0) % 5) - 0.4
    Xt = 1.7 + 0.2 * np.arange(Nt * 3)
    Xn = np.tile([0.0, 1.0, 0.0], Ns) if nd else None
    Vt = port.eval(kernel, Xs, Xn, Vs, Xt)
    M = port.kernel_matrix(kernel, Xs, Xn, Xt)
    np.testing.assert_allclose(Vt, M.T @ Vs, rtol=1e-10, atol=1e-10)


def test_rsqrt_contract(port):
    # vec-test.hpp:431-465: approx_rsqrt<4> < 2^-11, <7> < 2^-22, masked lanes exactly 0; the Newton recipe of
    # intrin-wrapper.hpp:540-559 with the AVX-512 iteration counts
    x = np.concatenate([10.0 ** np.linspace(-12, 12, 200), [0.0]])
    exact = np.where(x > 0, 1 / np.sqrt(np.where(x > 0, x, 1)), 0)
    assert [port.newton_iters(8, d) for d in (3, 4, 6, 8, 9, 13, 15, -1)] == [0, 0, 1, 1, 2, 2, 2, 2]
    assert [port.newton_iters(4, d) for d in (4, 6, -1)] == [0, 1, 1]
    for digits, bound in ((4, 2.0 ** -11), (7, 2.0 ** -22), (15, 1e-15)):
        y = port.rsqrt_newton(x, 14, port.newton_iters(8, digits))
        err = np.abs(y[:-1] - exact[:-1]) * np.sqrt(x[:-1])
        assert err.max() < bound, (digits, err.max())
        assert y[-1] == 0.0


# ---- golden vectors produced by the unmodified reference (tests/golden/make_golden.py) -------------------------

@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kernel", KERNELS)
def test_port_matches_golden_eval(port, golden, kernel, dtype):
    for Ns, Nt, variant, digits in cases.EVAL_CASES:
        r_src, n_src, v_src, r_trg, v0 = cases.inputs(kernel, Ns, Nt, variant, dtype)
        got = port.eval(kernel, r_src, n_src, v_src, r_trg, v_trg=v0, dtype=dtype)
        want = golden[cases.key(kernel, dtype, Ns, Nt, variant, digits)]
        assert got.dtype == want.dtype and np.all(np.isfinite(got))
        assert rel_err(got, want) <= TOL[dtype], (kernel, Ns, Nt, variant, rel_err(got, want))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kernel", KERNELS)
def test_port_matches_golden_kernel_matrix(port, golden, kernel, dtype):
    for Ns, Nt in cases.KMAT_CASES:
        r_src, n_src, _, r_trg, _ = cases.inputs(kernel, Ns, Nt, "cube", dtype)
        got = port.kernel_matrix(kernel, r_src, n_src, r_trg, dtype=dtype)
        want = golden[cases.kmat_key(kernel, dtype, Ns, Nt)]
        assert got.shape == want.shape
        assert rel_err(got, want) <= TOL[dtype] * (10 if dtype == np.float32 else 1)


def test_golden_rsqrt_contract(golden):
    x = golden["rsqrt/x"]
    exact = np.where(x > 0, 1 / np.sqrt(np.where(x > 0, x, 1)), 0)
    for digits, bound in ((4, 2.0 ** -11), (7, 2.0 ** -22), (10, 1e-10), (-1, 1e-15)):
        y = golden["rsqrt/f64/%d" % digits]
        assert np.all(y[x == 0] == 0)
        assert (np.abs(y - exact) * np.sqrt(x)).max() < bound
    for digits, bound in ((4, 2.0 ** -11), (7, 2.0 ** -22), (-1, 1e-6)):
        y = golden["rsqrt/f32/%d" % digits].astype(np.float64)
        xf = x.astype(np.float32).astype(np.float64)
        ex = np.where(xf > 0, 1 / np.sqrt(np.where(xf > 0, xf, 1)), 0)
        assert np.all(y[xf == 0] == 0)
        assert (np.abs(y - ex) * np.sqrt(xf)).max() < bound


# ---- the live reference, where oracle/_ref exists ---------------------------------------------------------------

@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kernel", KERNELS)
def test_port_matches_live_reference(port, ref, kernel, dtype):
    k0, k1, nd = cases.KERNELS[kernel]
    rng = np.random.default_rng(1234)
    for Ns, Nt in ((2000, 1500), (3, 17), (4096, 64)):
        xs, xt = rng.random(Ns * 3), rng.random(Nt * 3)
        n = rng.random(Ns * 3) - 0.5 if nd else None
        f = rng.random(Ns * k0) - 0.5
        a = port.eval(kernel, xs, n, f, xt, dtype=dtype)
        b = ref.eval(kernel, xs, n, f, xt, dtype=dtype)
        assert rel_err(a, b) <= TOL[dtype]
        c = ref.eval(kernel, xs, n, f, xt, dtype=dtype, omp=False)
        assert rel_err(c, b) <= TOL[dtype]


def test_reference_golden_is_current(ref, golden):
    """The committed golden file is what the reference produces today (guards against a stale fixture)."""
    for kernel in ("Stokes3D_FxU", "Laplace3D_DxU"):
        for dtype in (np.float64, np.float32):
            Ns, Nt, variant, digits = cases.EVAL_CASES[5]
            r_src, n_src, v_src, r_trg, v0 = cases.inputs(kernel, Ns, Nt, variant, dtype)
            got = ref.eval(kernel, r_src, n_src, v_src, r_trg, v_trg=v0, dtype=dtype, digits=digits)
            assert rel_err(got, golden[cases.key(kernel, dtype, Ns, Nt, variant, digits)]) <= TOL[dtype]


def test_splitmix_matches_c_generator(port):
    from sctl_b200 import synth
    for seed, n, first in ((1, 1000, 0), (4, 257, 12345)):
        assert np.array_equal(synth.splitmix_u01(seed, n, first), port.splitmix(seed, n, first=first))


def test_near_accumulate_restatement_small():
    """oracle/near_blocks.near_accumulate against a hand-worked case of ComputeNearInterac's tail
    (boundary_integral.txx:1366-1381; ScatterForward on one process is out[p] = in[scatter[p]], comm.txx:1499-1510)."""
    from oracle import near_blocks as nbo
    U_near = np.array([10.0, 20.0, 30.0, 40.0])          # one value per near pair, element-major
    scatter = np.array([2, 0, 3, 1])                      # scattered = [30, 10, 40, 20]
    cnt, dsp = np.array([1, 0, 3]), np.array([0, 0, 1])   # target 0 <- {30}; target 1 <- {}; target 2 <- {10, 40, 20}
    U = nbo.near_accumulate(1, np.array([1.0, 2.0, 3.0]), U_near, cnt, dsp, scatter)
    assert U.tolist() == [31.0, 2.0, 73.0]
    U = nbo.near_accumulate(2, np.zeros(4), np.arange(8.0), np.array([2, 0]), np.array([0, 0]), np.array([3, 1, 0, 2]))
    assert U.tolist() == [6 + 2, 7 + 3, 0, 0]
