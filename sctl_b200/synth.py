"""Synthetic inputs of SURVEY.md section 8(d): counter-based SplitMix64 streams, unit cube and torus surface."""
import numpy as np

SEED_R_SRC, SEED_R_TRG, SEED_N_SRC, SEED_V_SRC = 1, 2, 3, 4


def splitmix_u01(seed, n, first=0):
    """u(seed, i) in [0, 1) for i = first .. first+n-1 (uint64 wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        i = np.arange(first, first + n, dtype=np.uint64)
        z = (np.uint64(seed) << np.uint64(32)) + i + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def cube(Ns, Nt, k0, nd, dtype=np.float64, first_src=0, first_trg=0):
    """Uniform-in-cube sources/targets; un-normalised normals and densities in [-0.5, 0.5) (fmm-wrapper.txx:51 style).
    Generated in f64 then cast so both precisions see the same points.  `first_*` selects a shard's slice."""
    r_src = splitmix_u01(SEED_R_SRC, Ns * 3, first_src * 3)
    r_trg = splitmix_u01(SEED_R_TRG, Nt * 3, first_trg * 3)
    n_src = splitmix_u01(SEED_N_SRC, Ns * nd, first_src * nd) - 0.5 if nd else None
    v_src = splitmix_u01(SEED_V_SRC, Ns * k0, first_src * k0) - 0.5
    cast = lambda a: None if a is None else a.astype(dtype)
    return cast(r_src), cast(n_src), cast(v_src), cast(r_trg)


def torus(Nu, Nv, R=1.0, a=0.25, dtype=np.float64):
    """Nu x Nv torus grid: nodes, outward unit normals, trapezoid weights, density 1 + 0.5 sin(3u) cos(2v)."""
    u = (np.arange(Nu) * (2 * np.pi / Nu))[:, None]
    v = (np.arange(Nv) * (2 * np.pi / Nv))[None, :]
    rho = R + a * np.cos(v)
    X = np.stack([rho * np.cos(u), rho * np.sin(u), a * np.sin(v) + 0 * u], axis=-1).reshape(-1, 3)
    Nrm = np.stack([np.cos(v) * np.cos(u), np.cos(v) * np.sin(u), np.sin(v) + 0 * u], axis=-1).reshape(-1, 3)
    w = (a * rho * (2 * np.pi / Nu) * (2 * np.pi / Nv) + 0 * u).reshape(-1)
    sigma = (1 + 0.5 * np.sin(3 * u) * np.cos(2 * v)).reshape(-1)
    return X.reshape(-1).astype(dtype), Nrm.reshape(-1).astype(dtype), w.astype(dtype), sigma.astype(dtype)
