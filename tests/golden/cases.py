"""Definition of the golden cases (shared by make_golden.py and the tests).  Inputs are regenerated from
sctl_b200.synth (SplitMix64 streams); only the reference's OUTPUTS are stored in golden_eval.npz."""
import numpy as np

from sctl_b200 import synth

KERNELS = {"Laplace3D_FxU": (1, 1, 0), "Laplace3D_DxU": (1, 1, 3), "Laplace3D_FxdU": (1, 3, 0), "Stokes3D_FxU": (3, 3, 0), "Stokes3D_DxU": (3, 3, 3),
           # beyond the north-star five
           "Stokes3D_FxT": (3, 9, 0), "Stokes3D_FSxU": (4, 3, 0), "Stokes3D_FxUP": (3, 4, 0), "Laplace3D_Fxd2U": (1, 9, 0), "Laplace3D_DxdU": (1, 3, 3),
           "BiotSavart3D_FxU": (3, 3, 0), "BiotSavart3D_FxdU": (3, 9, 0),
           "Helmholtz3D_FxU": (2, 2, 0), "Helmholtz3D_DxU": (2, 2, 3), "Helmholtz3D_FxdU": (2, 6, 0), "HelmholtzDiff3D_FxdU": (2, 6, 0)}
# kernel context (the Helmholtz wavenumber) used by every fixture; None for the kernels without one
CTX = {k: (1.7 if "Helmholtz" in k else None) for k in KERNELS}  # == oracle.DEFAULT_HELMHOLTZ_CTX
NORTH_STAR = ["Laplace3D_FxU", "Laplace3D_DxU", "Laplace3D_FxdU", "Stokes3D_FxU", "Stokes3D_DxU"]

# (Ns, Nt, variant, digits): variants -- "cube": distinct sets; "self": targets == sources (r2 == 0 pairs);
# "acc": v_trg pre-filled (accumulate semantics); shapes cover Nt in {1,7,8,9,1000} and Ns in {1, odd, 1000}
EVAL_CASES = [
    (1, 1, "cube", -1), (7, 5, "cube", -1), (33, 8, "cube", -1), (129, 9, "cube", -1), (1000, 7, "cube", -1),
    (1000, 1000, "cube", -1), (500, 500, "self", -1), (301, 64, "acc", -1), (257, 33, "cube", 10), (257, 33, "cube", 11),
    (64, 1000, "cube", -1),
]
KMAT_CASES = [(1, 1), (7, 5), (33, 40), (20, 1), (1, 20)]


def inputs(kernel, Ns, Nt, variant, dtype):
    k0, k1, nd = KERNELS[kernel]
    r_src, n_src, v_src, r_trg = synth.cube(Ns, Nt, k0, nd, dtype=np.float64)
    if variant == "self":
        r_trg = r_src[: Nt * 3].copy()
    v0 = None
    if variant == "acc":
        v0 = synth.splitmix_u01(9, Nt * k1) - 0.5
    cast = lambda a: None if a is None else a.astype(dtype)
    return cast(r_src), cast(n_src), cast(v_src), cast(r_trg), cast(v0)


def key(kernel, dtype, Ns, Nt, variant, digits):
    return "%s/%s/%d/%d/%s/%d" % (kernel, np.dtype(dtype).name, Ns, Nt, variant, digits)


def kmat_key(kernel, dtype, Ns, Nt):
    return "kmat/%s/%s/%d/%d" % (kernel, np.dtype(dtype).name, Ns, Nt)
