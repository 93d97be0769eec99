"""GPU parity tests of the batched near-field blocks: sctl_cuda_near_* against the numpy restatement of
BoundaryIntegralOp's KNearSubtract / EvalNear loops (oracle/near_blocks.py, boundary_integral.txx:1165-1240,1324-1340)
on a synthetic block list with the reference's structure: elements own contiguous node ranges, each element has a
contiguous range of near targets, and the ranges are cut into blocks of at most BLK targets (:1009-1032)."""
import numpy as np
import pytest

import cases
from conftest import TOL, rel_err

pytestmark = pytest.mark.gpu


def block_list(n_elem, k0, k1, seed, blk=37, with_empty=True):
    """Per-element node counts / near-target counts -> per-block table in the reference's layout."""
    rng = np.random.RandomState(seed)
    nds = rng.randint(1, 70, size=n_elem)
    near = rng.randint(0, 120, size=n_elem)
    if with_empty:
        nds[3] = 0      # element without nodes: its blocks write zeros (:1327-1331)
        near[5] = 0     # element without near targets: no blocks
    nds_dsp = np.concatenate([[0], np.cumsum(nds)[:-1]])
    far_dsp = nds_dsp + 11  # far-field nodes live in their own arrays; shifted here to tell the two offsets apart
    near_dsp = np.concatenate([[0], np.cumsum(near)[:-1]])
    K_dsp = np.concatenate([[0], np.cumsum(nds * near)[:-1]])  # K_near_dsp (in units of K0*K1 entries)
    cols = [[] for _ in range(6)]
    for e in range(n_elem):
        for t0 in range(0, near[e], blk):
            n = min(blk, near[e] - t0)
            for c, v in zip(cols, (nds_dsp[e], far_dsp[e], nds[e], near_dsp[e] + t0, n, K_dsp[e] * k0 * k1 + t0 * k1 * nds[e] * k0)):
                c.append(v)
    tab = [np.array(c, dtype=np.int64) for c in cols]
    return tab, int(nds.sum()), int(far_dsp[-1] + nds[-1]), int(near.sum()), int((nds * near).sum()) * k0 * k1


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kernel", ["Laplace3D_FxU", "Laplace3D_DxU", "Stokes3D_FxU", "Stokes3D_DxU", "Laplace3D_FxdU", "Helmholtz3D_FxU"])
def test_near_blocks_match_the_reference_loops(sb, cpu_oracle, kernel, dtype):
    from oracle import near_blocks as nbo
    from sctl_b200 import synth
    k0, k1, nd = cases.KERNELS[kernel]
    tab, n_nodes, n_far, n_trg, K_len = block_list(40, k0, k1, seed=7)
    Xs, Xn, _, Xt = synth.cube(n_far, n_trg, k0, 3, dtype=dtype)
    Xn = Xn if nd else None
    wts = (synth.splitmix_u01(21, n_far) * 0.01 + 0.001).astype(dtype)
    K0_ = (synth.splitmix_u01(22, K_len) - 0.5).astype(dtype)
    nb = sb.NearBlocks(sb.KERNELS[kernel], *tab, dtype=dtype)
    assert (nb.K_len, nb.nU, nb.n_far) == (K_len, n_trg * k1, n_far) and nb.nF <= n_nodes * k0
    with pytest.raises(sb.SctlCudaError, match="no matrix"):
        nb.apply(np.zeros(nb.nF, dtype=dtype))
    nb.set_matrix(K0_)
    assert np.array_equal(nb.get_matrix(), K0_)
    # apply on the raw operator: pure batched GEMV
    F = (synth.splitmix_u01(23, n_nodes * k0) - 0.5).astype(dtype)
    want_U = nbo.near_apply(k0, k1, tab, K0_.astype(np.float64), F.astype(np.float64), n_trg * k1)
    U = nb.apply(F)
    tol = TOL[dtype]
    assert rel_err(U, want_U) <= (1e-13 if dtype == np.float64 else tol)
    # subtract the direct part on the device, compare the operator itself and the apply that uses it
    nb.subtract_direct(Xs, Xn, wts, Xt)
    want_K = nbo.near_subtract(cpu_oracle, kernel, k0, k1, nd, tab, K0_, Xs, Xn, wts, Xt, dtype=dtype)
    got_K = nb.get_matrix()
    assert rel_err(got_K, want_K) <= tol, rel_err(got_K, want_K)
    U2 = nb.apply(F, U)
    assert U2 is U
    assert rel_err(U, nbo.near_apply(k0, k1, tab, want_K.astype(np.float64), F.astype(np.float64), n_trg * k1)) <= tol
    # rows of an element without nodes are exactly zero; same inputs again give bit-identical results
    b = int(np.where(tab[2] == 0)[0][0])
    assert np.all(U[tab[3][b] * k1:(tab[3][b] + tab[4][b]) * k1] == 0)
    assert np.array_equal(nb.apply(F), U)
    nb.close()


def test_near_blocks_errors_and_empty(sb):
    e = np.zeros(0, dtype=np.int64)
    nb = sb.NearBlocks(sb.Laplace3D_FxU, e, e, e, e, e, e)
    assert (nb.K_len, nb.nF, nb.nU) == (0, 0, 0)
    nb.set_matrix(np.zeros(0))
    assert nb.apply(np.zeros(0)).size == 0
    nb.close()
    one, i64 = np.array([0], dtype=np.int64), lambda v: np.array([v], dtype=np.int64)
    nb = sb.NearBlocks(sb.Stokes3D_FxU, one, one, i64(4), one, i64(5), one)
    assert (nb.K_len, nb.nF, nb.nU, nb.n_far) == (4 * 5 * 9, 12, 15, 4)
    nb.set_matrix(np.zeros(nb.K_len))
    with pytest.raises(sb.SctlCudaError, match="F has"):
        nb.apply(np.zeros(3))
    with pytest.raises(sb.SctlCudaError, match="Xs has"):
        nb.subtract_direct(np.zeros(3), None, np.zeros(1), np.zeros(15))
    with pytest.raises(sb.SctlCudaError):
        sb.NearBlocks(sb.Stokes3D_FxU, i64(-1), one, one, one, one, one)
    nb.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kernel", ["Laplace3D_FxdU", "Stokes3D_FxT", "Laplace3D_DxdU"])
def test_near_blocks_trg_normal_dot_prod(sb, cpu_oracle, kernel, dtype):
    """trg_normal_dot_prod (boundary_integral.txx:1186-1200): the kernel's 3*K1_ target components are contracted with the
    target normals before the weighted transpose-subtract; the operator block then has K1_ rows per target."""
    from oracle import near_blocks as nbo
    from sctl_b200 import synth
    k0, k1, nd = cases.KERNELS[kernel]
    k1r = k1 // 3
    tab, n_nodes, n_far, n_trg, K_len = block_list(25, k0, k1r, seed=11)
    Xs, Xn, _, Xt = synth.cube(n_far, n_trg, k0, 3, dtype=dtype)
    Xn = Xn if nd else None
    Xn_trg = (synth.splitmix_u01(31, n_trg * 3) - 0.5).astype(dtype)
    wts = (synth.splitmix_u01(21, n_far) * 0.01 + 0.001).astype(dtype)
    K0_ = (synth.splitmix_u01(22, K_len) - 0.5).astype(dtype)
    nb = sb.NearBlocks(sb.KERNELS[kernel], *tab, dtype=dtype, trg_normal_dot_prod=True)
    assert (nb.K_len, nb.nU) == (K_len, n_trg * k1r)
    nb.set_matrix(K0_)
    with pytest.raises(sb.SctlCudaError, match="Xn_trg"):
        nb.subtract_direct(Xs, Xn, wts, Xt)
    nb.subtract_direct(Xs, Xn, wts, Xt, Xn_trg)
    want_K = nbo.near_subtract(cpu_oracle, kernel, k0, k1, nd, tab, K0_, Xs, Xn, wts, Xt, dtype=dtype, Xn_trg=Xn_trg)
    assert rel_err(nb.get_matrix(), want_K) <= TOL[dtype]
    F = (synth.splitmix_u01(23, n_nodes * k0) - 0.5).astype(dtype)
    U = nb.apply(F)
    assert rel_err(U, nbo.near_apply(k0, k1r, tab, want_K.astype(np.float64), F.astype(np.float64), n_trg * k1r)) <= TOL[dtype]
    nb.close()
    with pytest.raises(sb.SctlCudaError, match="multiple of 3"):
        sb.NearBlocks(sb.Laplace3D_FxU, *tab, dtype=dtype, trg_normal_dot_prod=True)


def near_target_tables(n_trg_op, n_near, seed):
    """near_trg_cnt / near_trg_dsp / near_scatter_index with the reference's structure (BuildNearList,
    boundary_integral.txx:490-520): the scatter index permutes the near pairs so that each target's contributions become
    contiguous; targets without near elements have count 0 and offset 0 (:507-509)."""
    rng = np.random.RandomState(seed)
    owner = rng.randint(0, n_trg_op, size=n_near)          # operator target of every near pair (element-major order)
    owner[rng.rand(n_near) < 0.1] = 3                      # one target with many contributions
    scatter = np.argsort(owner, kind="stable").astype(np.int64)
    cnt = np.bincount(owner, minlength=n_trg_op).astype(np.int64)
    dsp = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int64)
    dsp[cnt == 0] = 0
    return cnt, dsp, scatter


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kernel", ["Laplace3D_FxU", "Stokes3D_FxU"])
def test_near_apply_accumulate_matches_the_reference_tail(sb, kernel, dtype):
    """apply_accumulate = EvalNear + ScatterForward + per-target accumulation (boundary_integral.txx:1324-1340,1366-1381):
    bit-identical to the host tail applied to the session's own U_near (same order of additions)."""
    from oracle import near_blocks as nbo
    from sctl_b200 import synth
    k0, k1, nd = cases.KERNELS[kernel]
    tab, n_nodes, n_far, n_near, K_len = block_list(40, k0, k1, seed=5)
    n_trg_op = 517
    cnt, dsp, scatter = near_target_tables(n_trg_op, n_near, seed=9)
    nb = sb.NearBlocks(sb.KERNELS[kernel], *tab, dtype=dtype)
    K0_ = (synth.splitmix_u01(22, K_len) - 0.5).astype(dtype)
    F = (synth.splitmix_u01(23, n_nodes * k0) - 0.5).astype(dtype)
    U0 = (synth.splitmix_u01(24, n_trg_op * k1) - 0.5).astype(dtype)   # e.g. the far field
    with pytest.raises(sb.SctlCudaError, match="no matrix"):
        nb.apply_accumulate(F, U0.copy())
    nb.set_matrix(K0_)
    with pytest.raises(sb.SctlCudaError, match="no target tables"):
        nb.apply_accumulate(F, U0.copy())
    nb.set_targets(cnt, dsp, scatter)
    U_near = nb.apply(F)
    want = nbo.near_accumulate(k1, U0.copy(), U_near, cnt, dsp, scatter)
    got = nb.apply_accumulate(F, U0.copy())
    assert np.array_equal(got, want)
    # and against the numpy restatement of the whole thing
    want64 = nbo.near_accumulate(k1, U0.astype(np.float64), nbo.near_apply(k0, k1, tab, K0_.astype(np.float64), F.astype(np.float64), n_near * k1), cnt, dsp, scatter)
    assert rel_err(got, want64) <= (1e-13 if dtype == np.float64 else TOL[dtype])
    # accumulates: a second call adds the same correction again; wrong sizes and bad tables are refused
    twice = nb.apply_accumulate(F, got.copy())
    assert rel_err(twice - got, want64 - U0.astype(np.float64)) <= (1e-12 if dtype == np.float64 else 1e-4)
    with pytest.raises(sb.SctlCudaError, match="U has"):
        nb.apply_accumulate(F, np.zeros(n_trg_op * k1 + 1, dtype=dtype))
    bad = scatter.copy()
    bad[0] = n_near
    with pytest.raises(sb.SctlCudaError, match="scatter index"):
        nb.set_targets(cnt, dsp, bad)
    with pytest.raises(sb.SctlCudaError, match="near range"):
        nb.set_targets(cnt + 1, dsp, scatter)
    # plain apply still works after the tables were (re)set
    assert np.array_equal(nb.apply(F), U_near)
    nb.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kernel,ndot", [("Stokes3D_FxU", False), ("Laplace3D_DxU", False), ("Laplace3D_FxdU", True), ("Stokes3D_FxT", True)])
def test_bio_apply_pipeline_matches_the_staged_apply_and_the_oracle(sb, cpu_oracle, kernel, ndot, dtype):
    """sctl_cuda_bio_apply (csrc/bio_session.inc): BoundaryIntegralOp::ComputePotential (boundary_integral.txx:716-722) as one
    device-side pipeline -- F * wts, far field, normal contraction, near GEMV beside the far field, per-target accumulation,
    one upload and one download -- against (a) the same apply staged through the two sessions' host-array entries
    (bit-identical without trg_normal_dot_prod) and (b) the CPU oracle's Eval plus the numpy restatement of the near loops."""
    from oracle import near_blocks as nbo
    from sctl_b200 import synth
    k0, k1, nd = cases.KERNELS[kernel]
    k1r = k1 // 3 if ndot else k1
    tab, n_nodes, n_far, n_near, K_len = block_list(40, k0, k1r, seed=5)
    Ntrg = 517
    cnt, dsp, scatter = near_target_tables(Ntrg, n_near, seed=9)
    X_far, Xn_far, _, Xtrg = synth.cube(n_nodes, Ntrg, k0, 3, dtype=dtype)
    Xn_trg = (synth.splitmix_u01(31, Ntrg * 3) - 0.5).astype(dtype) if ndot else None
    wts = (0.01 + 0.001 * synth.splitmix_u01(32, n_nodes)).astype(dtype)
    K_near = (synth.splitmix_u01(22, K_len) - 0.5).astype(dtype)
    near = sb.NearBlocks(sb.KERNELS[kernel], *tab, dtype=dtype, trg_normal_dot_prod=ndot)
    near.set_matrix(K_near)
    near.set_targets(cnt, dsp, scatter)
    bio = sb.BioApply(sb.KERNELS[kernel], X_far, Xn_far, wts, Xtrg, Xn_trg, digits=-1, dtype=dtype, near=near)
    tol = 1e-12 if dtype == np.float64 else 2e-5
    U = None
    for it in range(3):  # a GMRES loop: new density, same geometry, the result array reused
        F = (synth.splitmix_u01(40 + it, n_nodes * k0) - 0.5).astype(dtype)
        U = bio.compute_potential(F, U=U)
        staged = bio.compute_potential_staged(F)
        if ndot:
            assert rel_err(U, staged) <= (1e-14 if dtype == np.float64 else 1e-6)
        else:
            assert np.array_equal(U, staged)
        f64 = lambda a: None if a is None else a.astype(np.float64)
        far = cpu_oracle.eval(kernel, f64(X_far), f64(Xn_far) if nd else None, (f64(F).reshape(-1, k0) * f64(wts)[:, None]).reshape(-1), f64(Xtrg))
        if ndot:
            far = (far.reshape(Ntrg, k1r, 3) * f64(Xn_trg).reshape(Ntrg, 1, 3)).sum(axis=2).reshape(-1)
        want = nbo.near_accumulate(k1r, far, nbo.near_apply(k0, k1r, tab, f64(K_near), f64(F), n_near * k1r), cnt, dsp, scatter)
        assert rel_err(U, want) <= tol, (kernel, ndot, it, rel_err(U, want))
    t = bio.last_timing()
    assert t["total_ms"] > 0 and t["far_ms"] > 0
    # a far-field density that differs from the node density (GetFarFieldDensity, :1267)
    Ffar = (synth.splitmix_u01(77, n_nodes * k0) - 0.5).astype(dtype)
    got = bio.compute_potential(F, F_far_nodes=Ffar)
    staged = bio.compute_potential_staged(F, F_far_nodes=Ffar)
    assert np.array_equal(got, staged) if not ndot else rel_err(got, staged) <= (1e-14 if dtype == np.float64 else 1e-6)
    # far field alone (element lists with dist_far = 0: no near list)
    bio.set_near(None)
    got = bio.compute_potential(F)
    staged = bio.compute_potential_staged(F)
    assert np.array_equal(got, staged) if not ndot else rel_err(got, staged) <= (1e-14 if dtype == np.float64 else 1e-6)
    # refusals: sizes that do not fit, a near-field session for another target count
    with pytest.raises(sb.SctlCudaError, match="density"):
        bio.compute_potential(F[:-1])
    near2 = sb.NearBlocks(sb.KERNELS[kernel], *tab, dtype=dtype, trg_normal_dot_prod=ndot)
    near2.set_matrix(K_near)
    c2, d2, s2 = near_target_tables(Ntrg + 1, n_near, seed=9)
    near2.set_targets(c2, d2, s2)
    bio.set_near(near2)
    with pytest.raises(sb.SctlCudaError, match="targets"):
        bio.compute_potential(F)
    bio.close()
    near.close()
    near2.close()
