"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the parity checkers.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  The product (``sctl_b200``, ``include/``, ``libsctl_cuda.so``) never does.

* :class:`PortOracle`  -- ``oracle/liboracle.so``: the plain-C restatement in ``oracle/sctl_oracle.c``.
* :class:`RefOracle`   -- ``oracle/_ref/libsctl_ref_<isa>.so``: the UNMODIFIED reference (SCTL) compiled from
  ``/root/reference/include`` by ``oracle/Makefile`` (``ref_shim.cpp`` only instantiates the reference templates).

Both expose ``eval`` (GenericKernel::Eval, generic-kernel.txx:102-211: accumulates into ``v_trg``) and
``kernel_matrix`` (generic-kernel.txx:213-329: overwrites ``M``) on numpy arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

KERNEL_IDS = {"Laplace3D_FxU": 0, "Laplace3D_DxU": 1, "Laplace3D_FxdU": 2, "Stokes3D_FxU": 3, "Stokes3D_DxU": 4, "Stokes3D_FxT": 5, "Stokes3D_FSxU": 6,
              "Stokes3D_FxUP": 7, "Laplace3D_Fxd2U": 8, "Laplace3D_DxdU": 9, "BiotSavart3D_FxU": 10, "BiotSavart3D_FxdU": 11,
              "Helmholtz3D_FxU": 12, "Helmholtz3D_DxU": 13, "Helmholtz3D_FxdU": 14, "HelmholtzDiff3D_FxdU": 15}
# (KDIM0, KDIM1, N_DIM): generic-kernel.hpp:22-41 applied to kernel_functions.hpp:43,72,92,136,191
KERNEL_DIMS = {0: (1, 1, 0), 1: (1, 1, 3), 2: (1, 3, 0), 3: (3, 3, 0), 4: (3, 3, 3), 5: (3, 9, 0), 6: (4, 3, 0), 7: (3, 4, 0), 8: (1, 9, 0), 9: (1, 3, 3),
               10: (3, 3, 0), 11: (3, 9, 0), 12: (2, 2, 0), 13: (2, 2, 3), 14: (2, 6, 0), 15: (2, 6, 0)}


# Wavenumber used by every Helmholtz fixture and test unless a call passes ctx explicitly (tests/golden/cases.py CTX)
DEFAULT_HELMHOLTZ_CTX = 1.7


def _kid(kernel):
    return KERNEL_IDS[kernel] if isinstance(kernel, str) else int(kernel)


def _ctx(kid, ctx):
    if ctx is None:
        return DEFAULT_HELMHOLTZ_CTX if kid >= 12 else 0.0
    return float(ctx)


def build(verbose=False):
    """Compile the checkers (``make -C oracle``).  ``_ref`` is rebuilt only where /root/reference is mounted."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


class _Base:
    def _prep(self, kernel, r_src, n_src, v_src, r_trg, dtype):
        kid = _kid(kernel)
        k0, k1, nd = KERNEL_DIMS[kid]
        r_src = np.ascontiguousarray(r_src, dtype=dtype).reshape(-1)
        r_trg = np.ascontiguousarray(r_trg, dtype=dtype).reshape(-1)
        ns, nt = r_src.size // 3, r_trg.size // 3
        assert r_src.size == ns * 3 and r_trg.size == nt * 3
        if v_src is not None:
            v_src = np.ascontiguousarray(v_src, dtype=dtype).reshape(-1)
            assert v_src.size == ns * k0
        if nd:
            n_src = np.ascontiguousarray(n_src, dtype=dtype).reshape(-1)
            assert n_src.size == ns * nd
        else:
            n_src = None
        return kid, k0, k1, nd, ns, nt, r_src, n_src, v_src, r_trg

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class PortOracle(_Base):
    """The plain-C port (oracle/sctl_oracle.c)."""
    kind = "port"

    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        L = self.lib
        L.oracle_eval.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64] + [C.c_void_p] * 5 + [C.c_double]
        L.oracle_kernel_matrix.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64] + [C.c_void_p] * 4 + [C.c_int64, C.c_double]
        L.oracle_rsqrt_newton.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.oracle_newton_iters.argtypes = [C.c_int, C.c_int]
        L.oracle_splitmix_fill.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_double, C.c_void_p]
        L.oracle_dims.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_int]

    def threads(self):
        return int(self.lib.oracle_threads())

    def eval(self, kernel, r_src, n_src, v_src, r_trg, v_trg=None, dtype=np.float64, omp=True, digits=-1, ctx=None):
        kid, k0, k1, nd, ns, nt, r_src, n_src, v_src, r_trg = self._prep(kernel, r_src, n_src, v_src, r_trg, dtype)
        out = np.zeros(nt * k1, dtype=dtype) if v_trg is None else np.ascontiguousarray(v_trg, dtype=dtype).reshape(-1).copy()
        assert out.size == nt * k1
        rc = self.lib.oracle_eval(kid, out.itemsize, int(omp), ns, nt, self._p(r_src), self._p(n_src), self._p(v_src), self._p(r_trg), self._p(out), _ctx(kid, ctx))
        assert rc == 0, rc
        return out

    def kernel_matrix(self, kernel, r_src, n_src, r_trg, dtype=np.float64, omp=True, digits=-1, ctx=None):
        kid, k0, k1, nd, ns, nt, r_src, n_src, _, r_trg = self._prep(kernel, r_src, n_src, None, r_trg, dtype)
        M = np.full((ns * k0, nt * k1), np.nan, dtype=dtype)
        rc = self.lib.oracle_kernel_matrix(kid, M.itemsize, int(omp), ns, nt, self._p(r_src), self._p(n_src), self._p(r_trg), self._p(M), nt * k1, _ctx(kid, ctx))
        assert rc == 0, rc
        return M

    def rsqrt_newton(self, x, seed_bits, max_iter):
        x = np.ascontiguousarray(x)
        y = np.empty_like(x)
        assert self.lib.oracle_rsqrt_newton(x.itemsize, seed_bits, max_iter, x.size, self._p(x), self._p(y)) == 0
        return y

    def newton_iters(self, real_bytes, digits):
        return int(self.lib.oracle_newton_iters(real_bytes, digits))

    def splitmix(self, seed, n, offset=0.0, first=0):
        out = np.empty(n, dtype=np.float64)
        self.lib.oracle_splitmix_fill(seed, first, n, offset, self._p(out))
        return out

    def dims(self, kernel):
        return _dims(self.lib.oracle_dims, _kid(kernel))


def _dims(fn, kid):
    k0, k1, nd, fl = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    s64, s32 = C.c_double(), C.c_float()
    name = C.create_string_buffer(64)
    rc = fn(kid, C.byref(k0), C.byref(k1), C.byref(nd), C.byref(fl), C.byref(s64), C.byref(s32), name, 64)
    assert rc == 0
    return dict(k0=k0.value, k1=k1.value, nd=nd.value, flops=fl.value, scale64=s64.value, scale32=s32.value, name=name.value.decode())


class RefOracle(_Base):
    """The unmodified reference, compiled where it lies (oracle/_ref).  Raises FileNotFoundError if never built."""
    kind = "reference"

    def __init__(self, isa=None):
        flags = _cpu_flags()
        if isa is None:
            isa = "avx512" if {"avx512f", "avx512dq", "avx512bw", "avx512vl", "avx512cd"} <= flags else "avx2"
        path = os.path.join(_HERE, "_ref", "libsctl_ref_%s.so" % isa)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle` where /root/reference is mounted)")
        self.isa = isa
        self.lib = C.CDLL(path)
        L = self.lib
        L.sctl_ref_eval.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64] + [C.c_void_p] * 6
        L.sctl_ref_kernel_matrix.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64] + [C.c_void_p] * 5
        L.sctl_ref_rsqrt.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.sctl_ref_dims.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_int]
        L.sctl_ref_veclen.argtypes = [C.c_int]

    def threads(self):
        return int(self.lib.sctl_ref_threads())

    def veclen(self, real_bytes):
        return int(self.lib.sctl_ref_veclen(real_bytes))

    def eval(self, kernel, r_src, n_src, v_src, r_trg, v_trg=None, dtype=np.float64, omp=True, digits=-1, ctx=None):
        kid, k0, k1, nd, ns, nt, r_src, n_src, v_src, r_trg = self._prep(kernel, r_src, n_src, v_src, r_trg, dtype)
        out = np.zeros(nt * k1, dtype=dtype) if v_trg is None else np.ascontiguousarray(v_trg, dtype=dtype).reshape(-1).copy()
        assert out.size == nt * k1
        if nt == 0 or out.size == 0:
            return out
        mu = np.array([_ctx(kid, ctx)], dtype=dtype)  # the context is a pointer to one Real of the evaluation type
        rc = self.lib.sctl_ref_eval(kid, out.itemsize, int(omp), digits, ns, nt, self._p(r_src), self._p(n_src), self._p(v_src), self._p(r_trg), self._p(out), self._p(mu))
        assert rc == 0, rc
        return out

    def kernel_matrix(self, kernel, r_src, n_src, r_trg, dtype=np.float64, omp=True, digits=-1, ctx=None):
        kid, k0, k1, nd, ns, nt, r_src, n_src, _, r_trg = self._prep(kernel, r_src, n_src, None, r_trg, dtype)
        M = np.full((ns * k0, nt * k1), np.nan, dtype=dtype)
        if M.size:
            mu = np.array([_ctx(kid, ctx)], dtype=dtype)
            rc = self.lib.sctl_ref_kernel_matrix(kid, M.itemsize, int(omp), digits, ns, nt, self._p(r_src), self._p(n_src), self._p(r_trg), self._p(M), self._p(mu))
            assert rc == 0, rc
        return M

    def rsqrt(self, x, digits=-1):
        x = np.ascontiguousarray(x)
        y = np.empty_like(x)
        assert self.lib.sctl_ref_rsqrt(x.itemsize, digits, x.size, self._p(x), self._p(y)) == 0
        return y

    def dims(self, kernel):
        return _dims(self.lib.sctl_ref_dims, _kid(kernel))


def best_cpu_oracle():
    """RefOracle when oracle/_ref was built, else the C port."""
    try:
        return RefOracle()
    except (FileNotFoundError, OSError):
        return PortOracle()
