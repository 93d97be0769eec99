"""TEST INFRASTRUCTURE ONLY -- numpy restatement of BoundaryIntegralOp's near-field block phases.

Follows the reference loops in include/sctl/boundary_integral.txx:
  * near_subtract  "KNearSubtract" (:1165-1240), both trg_normal_dot_prod branches, with an empty K_direct
    (element lists whose FarFieldDensityOperatorTranspose is the identity): Mker = KernelMatrix(Xtrg_near_, X, Xn)
    (:1196-1197), rows scaled by wts[s] (:1198-1205), K_near_[i][j] -= Mker[j][i] (:1219-1229);
  * near_apply     "EvalNear" (:1324-1340): U_ = K_near_ * F_ per block; blocks without sources give zeros (:1327-1331).
  * near_accumulate  the tail of ComputeNearInterac (:1366-1381): ScatterForward on one process (comm.txx:1499-1510,
    out[p] = in[scatter[p]]) and U[i] += the near_trg_cnt[i] entries from near_trg_dsp[i] on, j ascending.
KernelMatrix itself comes from the oracle object passed in (the C port or the unmodified reference), which is pinned.
PINNING: the block loops act on private state of BoundaryIntegralOp (K_near, near_blk_*), so this numpy restatement is
not run against the reference directly; the same device code that it checks is pinned against the reference's own
BoundaryIntegralOp by tests/cpp/test_near_blocks_ref.cpp (private members opened in that test TU, K_near and
ComputeNearInterac compared on a torus with non-empty near lists, with and without trg_normal_dot_prod).  near_accumulate
is checked on a hand-worked case (tests/test_oracle.py) and the device code it checks is compared entry by entry with the
reference's own ScatterForward + accumulation loops in that same C++ test (bit-identical), and with ComputeNearInterac.
Block table columns as in include/sctl_cuda.h (f_dsp, far_dsp, ns, u_dsp, nt, k_dsp)."""
import numpy as np


def near_subtract(orc, kernel, k0, k1, nd, tab, K, Xs, Xn, wts, Xt, dtype=np.float64, Xn_trg=None):
    """k1 is the kernel's target dimension; with Xn_trg (trg_normal_dot_prod, :1186-1200) the operator has k1/3 rows per target."""
    f_dsp, far_dsp, ns, u_dsp, nt, k_dsp = tab
    K = np.array(K, dtype=dtype).reshape(-1)
    Xs, Xt = np.asarray(Xs, dtype=dtype).reshape(-1, 3), np.asarray(Xt, dtype=dtype).reshape(-1, 3)
    Xn = None if Xn is None else np.asarray(Xn, dtype=dtype).reshape(-1, 3)
    wts = np.asarray(wts, dtype=dtype).reshape(-1)
    for b in range(len(ns)):
        if ns[b] == 0 or nt[b] == 0:
            continue
        s = slice(far_dsp[b], far_dsp[b] + ns[b])
        t = slice(u_dsp[b], u_dsp[b] + nt[b])
        M = orc.kernel_matrix(kernel, Xs[s].reshape(-1), None if Xn is None else Xn[s].reshape(-1), Xt[t].reshape(-1), dtype=dtype)  # (ns*k0) x (nt*k1)
        M = M * np.repeat(wts[s], k0)[:, None]
        k1r = k1
        if Xn_trg is not None:  # Mker[s*K0+k0][t*K1_+k1] = sum_l Mker_[s*K0+k0][(t*K1_+k1)*3+l] * wts[s] * Xn_trg[t*3+l]   (:1190-1199)
            k1r = k1 // 3
            nrm = np.asarray(Xn_trg, dtype=dtype).reshape(-1, 3)[t]
            M = np.einsum("atkl,tl->atk", M.reshape(ns[b] * k0, nt[b], k1r, 3), nrm).reshape(ns[b] * k0, nt[b] * k1r)
        rows, cols = nt[b] * k1r, ns[b] * k0
        blk = K[k_dsp[b]:k_dsp[b] + rows * cols].reshape(rows, cols)
        blk -= M.T.astype(dtype)
    return K


def near_apply(k0, k1, tab, K, F, nU, dtype=np.float64):
    f_dsp, far_dsp, ns, u_dsp, nt, k_dsp = tab
    K = np.asarray(K, dtype=dtype).reshape(-1)
    F = np.asarray(F, dtype=dtype).reshape(-1)
    U = np.zeros(nU, dtype=dtype)
    for b in range(len(ns)):
        rows, cols = nt[b] * k1, ns[b] * k0
        if rows == 0:
            continue
        if cols == 0:
            U[u_dsp[b] * k1:u_dsp[b] * k1 + rows] = 0
            continue
        blk = K[k_dsp[b]:k_dsp[b] + rows * cols].reshape(rows, cols)
        U[u_dsp[b] * k1:u_dsp[b] * k1 + rows] = blk @ F[f_dsp[b] * k0:f_dsp[b] * k0 + cols]
    return U


def near_accumulate(k1, U, U_near, near_trg_cnt, near_trg_dsp, near_scatter_index):
    """U (n_trg*k1, changed in place and returned) += scattered U_near, in the reference's order (sequential adds per target)."""
    U_near = np.asarray(U_near).reshape(-1, k1)
    scattered = U_near[np.asarray(near_scatter_index, dtype=np.int64)]  # ScatterForward, npes == 1
    Uv = U.reshape(-1, k1)
    for i in range(Uv.shape[0]):
        for j in range(int(near_trg_cnt[i])):
            Uv[i] += scattered[int(near_trg_dsp[i]) + j]
    return U
