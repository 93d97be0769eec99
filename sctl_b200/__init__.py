"""sctl_b200 -- B200-native dense kernel summation behind SCTL's GenericKernel API.

Python host-side mirror of ``sctl::GenericKernel<uKernel>`` (reference: include/sctl/generic-kernel.hpp:48-136)
for the five in-scope micro-kernels (include/sctl/kernel_functions.hpp:32-218).  Everything computes through the
C ABI in ``include/sctl_cuda.h`` (``libsctl_cuda.so``, hand-written sm_100a CUDA); there is no CPU fallback and
importing the compute entries without the built library raises.

    import sctl_b200 as sb
    K = sb.Stokes3D_FxU
    v = K.Eval(None, r_trg, r_src, None, v_src)         # numpy in / numpy out (host path, H2D + kernel + D2H)
    sb.eval_device("Stokes3D_FxU", d_r_src, None, d_v_src, d_r_trg, d_v_trg)   # torch CUDA tensors, in place
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsctl_cuda.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "sctl_cuda.h")

KERNEL_IDS = {"Laplace3D_FxU": 0, "Laplace3D_DxU": 1, "Laplace3D_FxdU": 2, "Stokes3D_FxU": 3, "Stokes3D_DxU": 4, "Stokes3D_FxT": 5, "Stokes3D_FSxU": 6,
              "Stokes3D_FxUP": 7, "Laplace3D_Fxd2U": 8, "Laplace3D_DxdU": 9, "BiotSavart3D_FxU": 10, "BiotSavart3D_FxdU": 11,
              "Helmholtz3D_FxU": 12, "Helmholtz3D_DxU": 13, "Helmholtz3D_FxdU": 14, "HelmholtzDiff3D_FxdU": 15}

_lib = None


class SctlCudaError(RuntimeError):
    pass


def load_library():
    """Load libsctl_cuda.so (built in-tree by ``__graft_entry__.build()`` / ``make -C sctl_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SctlCudaError("%s is missing: build it with `make -C sctl_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    i, i64, vp, dp = C.c_int, C.c_int64, C.c_void_p, C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    L.sctl_cuda_device_count.restype = i
    L.sctl_cuda_init.argtypes = [i]
    L.sctl_cuda_num_gpus.restype = i
    L.sctl_cuda_last_error.restype = C.c_char_p
    L.sctl_cuda_kernel_info.argtypes = [i, ip, ip, ip, ip, ip, dp, C.POINTER(C.c_char_p)]
    L.sctl_cuda_eval.argtypes = [i, i, i, i64, i64, vp, vp, vp, vp, vp, vp, i64]
    L.sctl_cuda_kernel_matrix.argtypes = [i, i, i, i64, i64, vp, vp, vp, vp, i64, vp, i64]
    L.sctl_cuda_eval_device.argtypes = [i, i, i, i64, i64, vp, vp, vp, vp, vp, vp, i64, vp]
    L.sctl_cuda_kernel_matrix_device.argtypes = [i, i, i, i64, i64, vp, vp, vp, vp, i64, vp, i64, vp]
    L.sctl_cuda_rsqrt_device.argtypes = [i, i, i64, vp, vp, vp]
    L.sctl_cuda_last_timing.argtypes = [dp, dp, dp]
    L.sctl_cuda_idle_thread_resources.argtypes = [i]
    L.sctl_cuda_idle_thread_resources.restype = i
    L.sctl_cuda_last_plan.argtypes = [ip, ip, ip, ip, ip, ip, dp]
    L.sctl_cuda_fma_peak.argtypes = [i, C.c_double, dp]
    L.sctl_cuda_set_option.argtypes = [C.c_char_p, i64]
    L.sctl_cuda_const_geometry.argtypes = [i, i, i64, i64, i, ip, ip, ip, ip, ip, C.POINTER(i64), C.POINTER(i64)]
    L.sctl_cuda_direct_create.argtypes = [i, i, C.POINTER(vp)]
    L.sctl_cuda_direct_create_on.argtypes = [i, i, ip, C.POINTER(vp)]
    L.sctl_cuda_direct_devices.argtypes = [vp, ip, ip, i]
    L.sctl_cuda_direct_destroy.argtypes = [vp]
    L.sctl_cuda_direct_set_src_coord.argtypes = [vp, i, i64, vp, vp, i]
    L.sctl_cuda_direct_set_src_density.argtypes = [vp, i, i64, vp]
    L.sctl_cuda_direct_set_trg_coord.argtypes = [vp, i, i64, vp]
    L.sctl_cuda_direct_set_kernel_s2t.argtypes = [vp, i, i, i, vp, i64]
    L.sctl_cuda_direct_eval.argtypes = [vp, i, i, vp, i64]
    p64 = C.POINTER(i64)
    L.sctl_cuda_near_create.argtypes = [i, i, i, i64, p64, p64, p64, p64, p64, p64, C.POINTER(vp)]
    L.sctl_cuda_near_destroy.argtypes = [vp]
    L.sctl_cuda_near_sizes.argtypes = [vp, p64, p64, p64, p64]
    L.sctl_cuda_near_set_matrix.argtypes = [vp, vp, i64]
    L.sctl_cuda_near_get_matrix.argtypes = [vp, vp, i64]
    L.sctl_cuda_near_subtract_direct.argtypes = [vp, i, i, i64, vp, vp, vp, i64, vp, vp, vp, i64]
    L.sctl_cuda_near_apply.argtypes = [vp, vp, i64, vp, i64]
    L.sctl_cuda_near_set_targets.argtypes = [vp, i64, p64, p64, i64, p64]
    L.sctl_cuda_near_apply_accumulate.argtypes = [vp, vp, i64, vp, i64]
    L.sctl_cuda_plan_near_slices.argtypes = [i, i, i, i64, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), i, C.POINTER(i64), C.POINTER(i64)]
    L.sctl_cuda_plan_eval.argtypes = [i64, i64, i, i, i, i, C.POINTER(i64), ip, C.POINTER(i64), dp]
    L.sctl_cuda_bio_create.argtypes = [vp, i, i, vp, i, vp, i64, vp, i64, C.POINTER(vp)]
    L.sctl_cuda_bio_destroy.argtypes = [vp]
    L.sctl_cuda_bio_apply.argtypes = [vp, i, vp, i64, vp, i64, vp, i64]
    L.sctl_cuda_bio_last_timing.argtypes = [vp, dp, dp, dp]
    L.sctl_cuda_bio_near_shards.argtypes = [vp, C.POINTER(i64), i]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise SctlCudaError(load_library().sctl_cuda_last_error().decode() or "sctl_cuda call failed (rc=%d)" % rc)


def _kid(kernel):
    if isinstance(kernel, GenericKernel):
        return kernel.kernel_id
    if isinstance(kernel, str):
        return KERNEL_IDS[kernel]
    return int(kernel)


def device_count():
    return int(load_library().sctl_cuda_device_count())


def _prefer_newest_nccl():
    """The C library dlopen()s ``libnccl.so.2`` when it is asked to shard over several GPUs.  A process can only hold ONE
    library of that soname: if the system's NCCL is loaded first and torch is imported afterwards, torch binds to it and may
    miss symbols of the newer NCCL it ships (``undefined symbol: ncclDevCommCreate``).  So before the C library loads NCCL,
    the copy that the Python environment ships (the ``nvidia-nccl`` wheel torch depends on) is loaded, if there is one and
    nothing has loaded NCCL yet; both users then share it."""
    import sys
    if "torch" in sys.modules:
        return  # torch has loaded its own NCCL already; the C library's dlopen finds that one by soname
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return
    except (ImportError, OSError, ValueError):
        pass


def init(n_gpus=0):
    """sctl_cuda_init: shard host-pointer calls over ``n_gpus`` GPUs of this process (0 = all visible)."""
    if int(n_gpus) != 1:
        _prefer_newest_nccl()
    _check(load_library().sctl_cuda_init(int(n_gpus)))
    return int(load_library().sctl_cuda_num_gpus())


def finalize():
    _check(load_library().sctl_cuda_finalize())


def kernel_info(kernel):
    L = load_library()
    k0, k1, nd, fl, fa = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
    sc, nm = C.c_double(), C.c_char_p()
    _check(L.sctl_cuda_kernel_info(_kid(kernel), C.byref(k0), C.byref(k1), C.byref(nd), C.byref(fl), C.byref(fa), C.byref(sc), C.byref(nm)))
    return dict(k0=k0.value, k1=k1.value, nd=nd.value, flops=fl.value, f_alg=fa.value, scale=sc.value, name=nm.value.decode())


def plan_near_slices(real_bytes, k0, k1, ns, u_dsp, nt, k_dsp, n_gpus):
    """The slices the multi-GPU apply pipeline cuts the near-field GEMV into (host logic, no GPU needed): a list of dicts
    (i0, i1, klo, khi, ulo, uhi) per GPU, the length of the work list, and whether the block list can be sharded at all."""
    cols = [np.ascontiguousarray(a, dtype=np.int64) for a in (ns, u_dsp, nt, k_dsp)]
    n_blk = cols[0].size
    out = np.zeros(6 * n_gpus, dtype=np.int64)
    n_items = C.c_int64()
    p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    rc = load_library().sctl_cuda_plan_near_slices(real_bytes, k0, k1, n_blk, *[p64(a) for a in cols], n_gpus, p64(out), C.byref(n_items))
    if rc > 0:
        _check(rc)
    keys = ("i0", "i1", "klo", "khi", "ulo", "uhi")
    return [dict(zip(keys, (int(x) for x in out[6 * p:6 * p + 6]))) for p in range(n_gpus)], n_items.value, rc == 0


def last_timing():
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    load_library().sctl_cuda_last_timing(C.byref(a), C.byref(b), C.byref(c))
    return dict(h2d_ms=a.value, kernel_ms=b.value, d2h_ms=c.value)


def last_plan():
    v = [C.c_int() for _ in range(6)]
    w = C.c_double()
    load_library().sctl_cuda_last_plan(*[C.byref(x) for x in v], C.byref(w))
    keys = ["grid_x", "grid_y", "threads", "smem_bytes", "ctas_per_sm", "n_launches"]
    out = {k: x.value for k, x in zip(keys, v)}
    out["wave_eff"] = w.value
    return out


def set_option(name, value):
    """sctl_cuda_set_option: Eval path selection and tuning knobs (include/sctl_cuda.h lists them)."""
    _check(load_library().sctl_cuda_set_option(name.encode(), int(value)))


def const_geometry(kernel, real_bytes, Ns, Nt, banks_wanted=2):
    """Layout of the constant-bank path for one call (csrc/eval_const_kernel.cuh ConstGeom): pure host logic."""
    v = [C.c_int() for _ in range(5)]
    w = [C.c_int64() for _ in range(2)]
    _check(load_library().sctl_cuda_const_geometry(_kid(kernel), real_bytes, Ns, Nt, banks_wanted, *[C.byref(x) for x in v], *[C.byref(x) for x in w]))
    return dict(banks=v[0].value, chunk=v[1].value, row_words=v[2].value, bank_stride=v[3].value, rec_bytes=v[4].value, seg_bytes=w[0].value, n_chunks=w[1].value)


def reset_options():
    """sctl_cuda_reset_options: every tuning knob back to its built-in default."""
    _check(load_library().sctl_cuda_reset_options())


def plan_eval(Ns, Nt, tpb, slots, min_chunk, max_split):
    """The grid planner (csrc/eval_kernel.cuh plan_eval): pure host logic, no GPU needed."""
    nb, ch, sp, eff = C.c_int64(), C.c_int64(), C.c_int(), C.c_double()
    _check(load_library().sctl_cuda_plan_eval(Ns, Nt, tpb, slots, min_chunk, max_split, C.byref(nb), C.byref(sp), C.byref(ch), C.byref(eff)))
    return dict(n_tblocks=nb.value, n_split=sp.value, chunk=ch.value, wave_eff=eff.value)


def fma_peak(real_bytes=8, ms=200.0):
    """Measured FMA throughput (TFLOP/s) of the current device: the FP64/FP32 CUDA-core roofline denominator."""
    t = C.c_double()
    _check(load_library().sctl_cuda_fma_peak(int(real_bytes), float(ms), C.byref(t)))
    return t.value


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def _dev_ptr(t):
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


def _real_bytes_of(t):
    return 8 if "64" in str(t.dtype) else 4


def _ctx_buf(ctx, real_bytes):
    if ctx is None:
        return None, 0
    buf = np.array([ctx], dtype=np.float64 if real_bytes == 8 else np.float32)
    return buf, buf.itemsize


def eval_device(kernel, d_r_src, d_n_src, d_v_src, d_r_trg, d_v_trg, digits=-1, stream=None, Ns=None, Nt=None, real_bytes=None, ctx=None):
    """sctl_cuda_eval_device on torch CUDA tensors (or raw device addresses): d_v_trg += K * d_v_src, asynchronous."""
    L = load_library()
    kid = _kid(kernel)
    if ctx is None and isinstance(kernel, GenericKernel):
        ctx = kernel.GetCtxPtr()
    if Ns is None:
        Ns = d_r_src.numel() // 3
    if Nt is None:
        Nt = d_r_trg.numel() // 3
    if real_bytes is None:
        real_bytes = _real_bytes_of(d_r_src)
    if stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    cbuf, nb = _ctx_buf(ctx, real_bytes)
    _check(L.sctl_cuda_eval_device(kid, real_bytes, int(digits), int(Ns), int(Nt), _dev_ptr(d_r_src), _dev_ptr(d_n_src), _dev_ptr(d_v_src),
                                   _dev_ptr(d_r_trg), _dev_ptr(d_v_trg), _np_ptr(cbuf), nb, C.c_void_p(stream)))


def kernel_matrix_device(kernel, d_r_src, d_n_src, d_r_trg, d_M, digits=-1, stream=None, ldM=None, ctx=None):
    L = load_library()
    kid = _kid(kernel)
    if ctx is None and isinstance(kernel, GenericKernel):
        ctx = kernel.GetCtxPtr()
    Ns, Nt = d_r_src.numel() // 3, d_r_trg.numel() // 3
    info = kernel_info(kid)
    if ldM is None:
        ldM = Nt * info["k1"]
    if stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    cbuf, nb = _ctx_buf(ctx, _real_bytes_of(d_r_src))
    _check(L.sctl_cuda_kernel_matrix_device(kid, _real_bytes_of(d_r_src), int(digits), Ns, Nt, _dev_ptr(d_r_src), _dev_ptr(d_n_src), _dev_ptr(d_r_trg),
                                            _dev_ptr(d_M), int(ldM), _np_ptr(cbuf), nb, C.c_void_p(stream)))


def rsqrt_device(d_x, d_y, digits=-1, stream=None):
    if stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    _check(load_library().sctl_cuda_rsqrt_device(_real_bytes_of(d_x), int(digits), d_x.numel(), _dev_ptr(d_x), _dev_ptr(d_y), C.c_void_p(stream)))


class GenericKernel:
    """Host-side mirror of ``sctl::GenericKernel<uKernel>`` (generic-kernel.hpp:48-136) for one micro-kernel.

    Static metadata (dims, name, FLOPS, scale factor) needs no GPU.  ``Eval`` / ``KernelMatrix`` take the
    reference's AoS arrays (numpy, any shape that flattens to ``[N*dim]``) and go through the host-pointer C ABI.
    """

    # (KDIM0, KDIM1, N_DIM), FLOPS(), Name(): kernel_functions.hpp:33-42, 62-71, 82-91, 126-135, 181-190
    _META = {
        0: (1, 1, 0, 6, "Laplace3D-FxU"), 1: (1, 1, 3, 14, "Laplace3D-DxU"), 2: (1, 3, 0, 11, "Laplace3D-FxdU"),
        3: (3, 3, 0, 23, "Stokes3D-FxU"), 4: (3, 3, 3, 26, "Stokes3D-DxU"),
        # kernel_functions.hpp:220-483
        5: (3, 9, 0, 39, "Stokes3D-FxT"), 6: (4, 3, 0, 26, "Stokes3D-FSxU"), 7: (3, 4, 0, 26, "Stokes3D-FxUP"), 8: (1, 9, 0, 25, "Laplace3D-Fxd2U"),
        9: (1, 3, 3, 22, "Laplace3D-DxdU"), 10: (3, 3, 0, 15, "BiotSavart3D-FxU"), 11: (3, 9, 0, 51, "BiotSavart3D-FxdU"),
        # Helmholtz (:490-659): complex values as (re, im) pairs; SetCtxPtr(wavenumber)
        12: (2, 2, 0, 20, "Helmholtz3D-FxU"), 13: (2, 2, 3, 29, "Helmholtz3D-DxU"), 14: (2, 6, 0, 35, "Helmholtz3D-FxdU"), 15: (2, 6, 0, 41, "HelmholtzDiff3D-FxdU"),
    }

    def __init__(self, kernel):
        self.kernel_id = KERNEL_IDS[kernel] if isinstance(kernel, str) else int(kernel)
        self._k0, self._k1, self._nd, self._flops, self._name = self._META[self.kernel_id]
        self._ctx = None

    def CoordDim(self):
        return 3

    def NormalDim(self):
        return self._nd

    def SrcDim(self):
        return self._k0

    def TrgDim(self):
        return self._k1

    def Name(self):
        return self._name

    def FLOPS(self):
        return self._flops

    def SetCtxPtr(self, ctx):
        """Kernel context (generic-kernel.hpp:79).  For the Helmholtz kernels: the wavenumber (a Python float; the
        reference takes a pointer to one Real of the evaluation type, kernel_functions.hpp:484-488)."""
        self._ctx = ctx

    def _ctx_arg(self, dtype):
        if self._ctx is None:
            return None, 0
        buf = np.array([self._ctx], dtype=dtype)
        return buf, buf.itemsize

    def GetCtxPtr(self):
        return self._ctx

    def uKerScaleFactor(self, dtype=np.float64):
        pi = dtype(np.pi)
        one = dtype(1)
        return {0: one / (4 * pi), 1: one / (4 * pi), 2: -one / (4 * pi), 3: one / (8 * pi), 4: dtype(3) / (4 * pi), 5: dtype(-3) / (4 * pi), 6: one / (8 * pi),
                7: one / (8 * pi), 8: one / (4 * pi), 9: one / (4 * pi), 10: one / (4 * pi), 11: -one / (4 * pi), 12: one / (4 * pi), 13: -one / (4 * pi),
                14: one / (4 * pi), 15: one / (4 * pi)}[self.kernel_id]

    def _arrays(self, r_trg, r_src, n_src, v_src, dtype, need_v=True):
        r_trg = np.ascontiguousarray(r_trg, dtype=dtype).reshape(-1)
        r_src = np.ascontiguousarray(r_src, dtype=dtype).reshape(-1)
        Ns, Nt = r_src.size // 3, r_trg.size // 3
        # SCTL_ASSERTs of generic-kernel.txx:125-128 (the reference aborts; here they raise)
        if r_trg.size != Nt * 3:
            raise SctlCudaError("r_trg.Dim() == Nt*DIM violated")
        if r_src.size != Ns * 3:
            raise SctlCudaError("r_src.Dim() == Ns*DIM violated")
        if need_v:
            v_src = np.ascontiguousarray(v_src, dtype=dtype).reshape(-1)
            if v_src.size != Ns * self._k0:
                raise SctlCudaError("v_src.Dim() == Ns*KDIM0 violated")
        if self._nd:
            n_src = np.ascontiguousarray(n_src if n_src is not None else [], dtype=dtype).reshape(-1)
            if n_src.size != Ns * self._nd:
                raise SctlCudaError("n_src.Dim() == Ns*N_DIM violated")
        else:
            n_src = None  # tolerated and ignored for single-layer kernels (:128)
        return Ns, Nt, r_trg, r_src, n_src, v_src

    def Eval(self, v_trg, r_trg, r_src, n_src, v_src, digits=-1, dtype=None):
        """``Eval<Real,omp,digits>`` (generic-kernel.txx:102-211).  Returns v_trg: accumulated into (a copy-free view of)
        the given array when it has Nt*K1 elements, else a fresh zero-initialised array (:129-132)."""
        if dtype is None:
            dtype = np.asarray(r_src).dtype.type if np.asarray(r_src).dtype in (np.float32, np.float64) else np.float64
        Ns, Nt, r_trg, r_src, n_src, v_src = self._arrays(r_trg, r_src, n_src, v_src, dtype)
        if v_trg is None or np.asarray(v_trg).size != Nt * self._k1:
            out = np.zeros(Nt * self._k1, dtype=dtype)
        else:
            out = v_trg if (isinstance(v_trg, np.ndarray) and v_trg.dtype == dtype and v_trg.flags.c_contiguous) else np.ascontiguousarray(v_trg, dtype=dtype)
        flat = out.reshape(-1)
        if digits >= 16:
            digits = -1  # generic-kernel.txx:97
        ctx, nb = self._ctx_arg(dtype)
        _check(load_library().sctl_cuda_eval(self.kernel_id, flat.itemsize, int(digits), Ns, Nt, _np_ptr(r_src), _np_ptr(n_src), _np_ptr(v_src),
                                             _np_ptr(r_trg), _np_ptr(flat), _np_ptr(ctx), nb))
        return out

    def KernelMatrix(self, Xt, Xs, Xn, digits=-1, dtype=None, M=None):
        """``KernelMatrix<Real,omp,digits>`` (generic-kernel.txx:213-329): (Ns*K0) x (Nt*K1), overwritten."""
        if dtype is None:
            dtype = np.asarray(Xs).dtype.type if np.asarray(Xs).dtype in (np.float32, np.float64) else np.float64
        Ns, Nt, Xt, Xs, Xn, _ = self._arrays(Xt, Xs, Xn, None, dtype, need_v=False)
        rows, cols = Ns * self._k0, Nt * self._k1
        if M is None or M.shape != (rows, cols) or M.dtype != dtype or not M.flags.c_contiguous:
            M = np.zeros((rows, cols), dtype=dtype)  # ReInit + SetZero (:220-223)
        ctx, nb = self._ctx_arg(dtype)
        _check(load_library().sctl_cuda_kernel_matrix(self.kernel_id, M.itemsize, int(digits), Ns, Nt, _np_ptr(Xs), _np_ptr(Xn), _np_ptr(Xt), _np_ptr(M), cols, _np_ptr(ctx), nb))
        return M


Laplace3D_FxU = GenericKernel("Laplace3D_FxU")
Laplace3D_DxU = GenericKernel("Laplace3D_DxU")
Laplace3D_FxdU = GenericKernel("Laplace3D_FxdU")
Stokes3D_FxU = GenericKernel("Stokes3D_FxU")
Stokes3D_DxU = GenericKernel("Stokes3D_DxU")
Stokes3D_FxT = GenericKernel("Stokes3D_FxT")
Stokes3D_FSxU = GenericKernel("Stokes3D_FSxU")
Stokes3D_FxUP = GenericKernel("Stokes3D_FxUP")
Laplace3D_Fxd2U = GenericKernel("Laplace3D_Fxd2U")
Laplace3D_DxdU = GenericKernel("Laplace3D_DxdU")
BiotSavart3D_FxU = GenericKernel("BiotSavart3D_FxU")
BiotSavart3D_FxdU = GenericKernel("BiotSavart3D_FxdU")
Helmholtz3D_FxU = GenericKernel("Helmholtz3D_FxU")  # SetCtxPtr(wavenumber) before Eval
Helmholtz3D_DxU = GenericKernel("Helmholtz3D_DxU")
Helmholtz3D_FxdU = GenericKernel("Helmholtz3D_FxdU")
HelmholtzDiff3D_FxdU = GenericKernel("HelmholtzDiff3D_FxdU")
KERNELS = {name: GenericKernel(name) for name in KERNEL_IDS}


class ParticleDirect:
    """Device-resident direct evaluator: the ``ParticleFMM::EvalDirect`` analogue (fmm-wrapper.hpp:35-200,
    fmm-wrapper.txx:551-623) over ``sctl_cuda_direct_*``.  Same method names and argument meaning as ``ParticleFMM``
    for the subset EvalDirect uses; coordinates are uploaded once and stay in HBM, an apply ships only the density in
    and the potential out.  ``n_gpus > 1`` shards the targets over GPUs 0..n_gpus-1 (results agree to rounding for any
    n_gpus); ``n_gpus == 1`` uses the device that is current on the creating thread (one rank per GPU: the rank's own
    GPU); ``devices=[...]`` names the CUDA device ordinals explicitly.  A kernel's context (the Helmholtz wavenumber
    set with ``SetCtxPtr``) is read again at every ``EvalDirect``, as the reference dereferences ``ctx_ptr`` at every
    Eval."""

    def __init__(self, dtype=np.float64, n_gpus=1, devices=None):
        self._dtype = np.dtype(dtype).type
        self._h = C.c_void_p()
        self._src, self._trg = {}, {}
        self._s2t = {}  # (src slot, trg slot) -> kernel object, for kernels with a context
        self._digits = 10  # ParticleFMM's default accuracy (fmm-wrapper.txx:205)
        if devices is not None:
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            _check(load_library().sctl_cuda_direct_create_on(np.dtype(dtype).itemsize, len(devices), devs, C.byref(self._h)))
        else:
            _check(load_library().sctl_cuda_direct_create(np.dtype(dtype).itemsize, int(n_gpus), C.byref(self._h)))

    def devices(self):
        """CUDA device ordinals of the session's GPUs."""
        n = C.c_int(0)
        devs = (C.c_int * 16)()
        _check(load_library().sctl_cuda_direct_devices(self._h, C.byref(n), devs, 16))
        return [int(devs[i]) for i in range(n.value)]

    def close(self):
        if self._h:
            load_library().sctl_cuda_direct_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def SetAccuracy(self, digits):
        self._digits = int(digits)

    def AddSrc(self, name, ker_s2m=None, ker_s2l=None):
        """``AddSrc(name, ker_s2m, ker_s2l)`` (fmm-wrapper.hpp:124); the FMM kernels are irrelevant to direct evaluation."""
        if name in self._src:
            raise SctlCudaError("Source name already exists.")
        self._src[name] = max(self._src.values(), default=-1) + 1

    def AddTrg(self, name, ker_m2t=None, ker_l2t=None):
        if name in self._trg:
            raise SctlCudaError("Target name already exists.")
        self._trg[name] = max(self._trg.values(), default=-1) + 1

    def _slot(self, table, name, what):
        if name not in table:
            raise SctlCudaError("%s name does not exist." % what)  # SCTL_ASSERT_MSG, fmm-wrapper.txx:555,574
        return table[name]

    def SetKernelS2T(self, src_name, trg_name, ker_s2t):
        k = ker_s2t if isinstance(ker_s2t, GenericKernel) else GenericKernel(ker_s2t)
        st = (self._slot(self._src, src_name, "Source"), self._slot(self._trg, trg_name, "Target"))
        self._push_s2t(st, k)
        self._s2t.pop(st, None)
        if k.GetCtxPtr() is not None:
            self._s2t[st] = k

    def _push_s2t(self, st, k):
        ctx, nb = k._ctx_arg(self._dtype)
        _check(load_library().sctl_cuda_direct_set_kernel_s2t(self._h, st[0], st[1], k.kernel_id, _np_ptr(ctx), nb))

    def DeleteSrc(self, name):
        s = self._slot(self._src, name, "Source")
        for t in self._trg.values():
            _check(load_library().sctl_cuda_direct_set_kernel_s2t(self._h, s, t, -1, None, 0))
            self._s2t.pop((s, t), None)
        _check(load_library().sctl_cuda_direct_set_src_coord(self._h, s, 0, None, None, 0))
        del self._src[name]

    def DeleteTrg(self, name):
        t = self._slot(self._trg, name, "Target")
        for s in self._src.values():
            _check(load_library().sctl_cuda_direct_set_kernel_s2t(self._h, s, t, -1, None, 0))
            self._s2t.pop((s, t), None)
        _check(load_library().sctl_cuda_direct_set_trg_coord(self._h, t, 0, None))
        del self._trg[name]

    def SetSrcCoord(self, name, src_coord, src_normal=None):
        X = np.ascontiguousarray(src_coord, dtype=self._dtype).reshape(-1)
        Ns = X.size // 3
        if X.size != Ns * 3:
            raise SctlCudaError("src_coord.Dim() == Ns*DIM violated")
        Xn, nd = None, 0
        if src_normal is not None and np.asarray(src_normal).size:
            Xn = np.ascontiguousarray(src_normal, dtype=self._dtype).reshape(-1)
            nd = 3
            if Xn.size != Ns * 3:
                raise SctlCudaError("src_normal.Dim() == Ns*DIM violated")
        _check(load_library().sctl_cuda_direct_set_src_coord(self._h, self._slot(self._src, name, "Source"), Ns, _np_ptr(X), _np_ptr(Xn), nd))

    def SetSrcDensity(self, name, src_density):
        F = np.ascontiguousarray(src_density, dtype=self._dtype).reshape(-1)
        _check(load_library().sctl_cuda_direct_set_src_density(self._h, self._slot(self._src, name, "Source"), F.size, _np_ptr(F)))

    def SetTrgCoord(self, name, trg_coord):
        X = np.ascontiguousarray(trg_coord, dtype=self._dtype).reshape(-1)
        Nt = X.size // 3
        if X.size != Nt * 3:
            raise SctlCudaError("trg_coord.Dim() == Nt*DIM violated")
        self._nt = getattr(self, "_nt", {})
        self._nt[name] = Nt
        _check(load_library().sctl_cuda_direct_set_trg_coord(self._h, self._slot(self._trg, name, "Target"), Nt, _np_ptr(X)))

    def EvalDirect(self, U, trg_name, trg_dim):
        """``EvalDirect(U, trg_name)`` (fmm-wrapper.txx:551-623): U is resized if needed and OVERWRITTEN.  ``trg_dim`` is
        the target dof of the kernels into ``trg_name`` (ParticleFMM reads it from the kernel given to AddTrg)."""
        t = self._slot(self._trg, trg_name, "Target")
        n = self._nt[trg_name] * int(trg_dim)
        if U is None or not isinstance(U, np.ndarray) or U.size != n or U.dtype != self._dtype or not U.flags.c_contiguous:
            U = np.zeros(n, dtype=self._dtype)
        for st, k in self._s2t.items():  # a context changed since SetKernelS2T takes effect, as in the reference
            if st[1] == t:
                self._push_s2t(st, k)
        _check(load_library().sctl_cuda_direct_eval(self._h, t, self._digits, _np_ptr(U.reshape(-1)), n))
        return U

    Eval = EvalDirect


class NearBlocks:
    """Batched near-field blocks of ``BoundaryIntegralOp`` on the device (``sctl_cuda_near_*``): the "KNearSubtract"
    setup phase (boundary_integral.txx:1165-1240) and the per-iteration "EvalNear" apply (:1324-1340) over one block
    list, with the operator resident in HBM.  Table columns are per block, in the units of include/sctl_cuda.h."""

    def __init__(self, kernel, f_dsp, far_dsp, ns, u_dsp, nt, k_dsp, dtype=np.float64, trg_normal_dot_prod=False):
        self.kernel = kernel if isinstance(kernel, GenericKernel) else GenericKernel(kernel)
        self.trg_normal_dot_prod = bool(trg_normal_dot_prod)
        if self.trg_normal_dot_prod and self.kernel.TrgDim() % 3:
            raise SctlCudaError("trg_normal_dot_prod needs a kernel whose target dimension is a multiple of 3")
        self.k1 = self.kernel.TrgDim() // 3 if self.trg_normal_dot_prod else self.kernel.TrgDim()  # KDIM1_ (boundary_integral.txx:1312)
        self._dtype = np.dtype(dtype).type
        cols = [np.ascontiguousarray(c, dtype=np.int64).reshape(-1) for c in (f_dsp, far_dsp, ns, u_dsp, nt, k_dsp)]
        n = cols[0].size
        if any(c.size != n for c in cols):
            raise SctlCudaError("block table columns differ in length")
        self._h = C.c_void_p()
        L = load_library()
        p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        _check(L.sctl_cuda_near_create(np.dtype(dtype).itemsize, self.kernel.SrcDim(), self.k1, n, *[p64(c) for c in cols], C.byref(self._h)))
        v = [C.c_int64() for _ in range(4)]
        _check(L.sctl_cuda_near_sizes(self._h, *[C.byref(x) for x in v]))
        self.K_len, self.nF, self.nU, self.n_far = [x.value for x in v]

    def close(self):
        if self._h:
            load_library().sctl_cuda_near_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_matrix(self, K):
        K = np.ascontiguousarray(K, dtype=self._dtype).reshape(-1)
        _check(load_library().sctl_cuda_near_set_matrix(self._h, _np_ptr(K), K.size))

    def get_matrix(self):
        K = np.zeros(self.K_len, dtype=self._dtype)
        _check(load_library().sctl_cuda_near_get_matrix(self._h, _np_ptr(K), K.size))
        return K

    def subtract_direct(self, Xs, Xn, wts, Xt, Xn_trg=None, digits=-1):
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=self._dtype).reshape(-1)
        Xs, Xn, wts, Xt, Xn_trg = c(Xs), c(Xn), c(wts), c(Xt), c(Xn_trg)
        if self.trg_normal_dot_prod != (Xn_trg is not None):
            raise SctlCudaError("Xn_trg must be given exactly when the session was created with trg_normal_dot_prod")
        ctx, nb = self.kernel._ctx_arg(self._dtype)
        _check(load_library().sctl_cuda_near_subtract_direct(self._h, self.kernel.kernel_id, int(digits), Xs.size // 3, _np_ptr(Xs), _np_ptr(Xn), _np_ptr(wts),
                                                             Xt.size // 3, _np_ptr(Xt), _np_ptr(Xn_trg), _np_ptr(ctx), nb))

    def apply(self, F, U=None):
        F = np.ascontiguousarray(F, dtype=self._dtype).reshape(-1)
        if U is None or U.size < self.nU or U.dtype != self._dtype or not U.flags.c_contiguous:
            U = np.zeros(self.nU, dtype=self._dtype)
        _check(load_library().sctl_cuda_near_apply(self._h, _np_ptr(F), F.size, _np_ptr(U.reshape(-1)), U.size))
        return U

    def set_targets(self, near_trg_cnt, near_trg_dsp, near_scatter_index):
        """Tables of ComputeNearInterac's tail (boundary_integral.txx:1366-1381), one-process ``Comm``."""
        cnt, dsp, sc = [np.ascontiguousarray(a, dtype=np.int64).reshape(-1) for a in (near_trg_cnt, near_trg_dsp, near_scatter_index)]
        if cnt.size != dsp.size:
            raise SctlCudaError("near_trg_cnt and near_trg_dsp differ in length")
        p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        _check(load_library().sctl_cuda_near_set_targets(self._h, cnt.size, p64(cnt), p64(dsp), sc.size, p64(sc)))
        self.n_trg = int(cnt.size)

    def apply_accumulate(self, F, U):
        """``U += `` the near-field correction ("EvalNear" + scatter + per-target accumulation); ``U`` is changed in place."""
        F = np.ascontiguousarray(F, dtype=self._dtype).reshape(-1)
        if not isinstance(U, np.ndarray) or U.dtype != self._dtype or not U.flags.c_contiguous:
            raise SctlCudaError("U must be a contiguous %s array (it is accumulated into)" % np.dtype(self._dtype).name)
        _check(load_library().sctl_cuda_near_apply_accumulate(self._h, _np_ptr(F), F.size, _np_ptr(U.reshape(-1)), U.size))
        return U


class BioApply:
    """One ``BoundaryIntegralOp::ComputePotential`` (boundary_integral.txx:716-722: ``ComputeFarField`` :1248-1309 +
    ``ComputeNearInterac`` :1311-1383) as one device-side pipeline over ``sctl_cuda_bio_*``: the density goes up once per
    GPU, ``F * wts_far`` and the target-normal contraction are device kernels, the near-field GEMV runs beside the
    far-field launches, the shards of a multi-GPU far field are collected on the first GPU, ``U`` comes down once.
    Mirrors ``sctl_cuda::BioApply<Real>`` (include/sctl_cuda/bio_apply.hpp); ``compute_potential_staged`` is the same apply
    phase after phase through the two sessions' host-array entries (the cross-check)."""

    def __init__(self, kernel, X_far, Xn_far, wts_far, Xtrg, Xn_trg=None, digits=11, dtype=np.float64, n_gpus=1, devices=None, near=None):
        self.kernel = kernel if isinstance(kernel, GenericKernel) else GenericKernel(kernel)
        self._dtype = np.dtype(dtype).type
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=self._dtype).reshape(-1)
        self._wts, self._nrm = c(wts_far), c(Xn_trg)
        self.ndot = Xn_trg is not None
        self.k0, self.k1 = self.kernel.SrcDim(), self.kernel.TrgDim()
        if self.ndot and self.k1 % 3:
            raise SctlCudaError("trg_normal_dot_prod needs a kernel whose target dimension is a multiple of 3")
        self.k1r = self.k1 // 3 if self.ndot else self.k1
        self.Nsrc, self.Ntrg = c(X_far).size // 3, c(Xtrg).size // 3
        self.far = ParticleDirect(dtype, n_gpus=n_gpus, devices=devices)
        self.far.SetAccuracy(digits)
        self.far.AddSrc("Src"); self.far.AddTrg("Trg")  # the names BoundaryIntegralOp uses (:573-575)
        self.far.SetKernelS2T("Src", "Trg", self.kernel)
        self.far.SetSrcCoord("Src", X_far, Xn_far)
        self.far.SetTrgCoord("Trg", Xtrg)
        self.near = near
        self._h = C.c_void_p()

    def set_near(self, near):
        self.close_pipeline()
        self.near = near

    def close_pipeline(self):
        if self._h:
            load_library().sctl_cuda_bio_destroy(self._h)
            self._h = C.c_void_p()

    def close(self):
        self.close_pipeline()
        self.far.close()

    def __del__(self):
        try:
            self.close_pipeline()
        except Exception:
            pass

    def compute_potential(self, F, F_far_nodes=None, U=None):
        L = load_library()
        F = np.ascontiguousarray(F, dtype=self._dtype).reshape(-1)
        Ff = None if F_far_nodes is None else np.ascontiguousarray(F_far_nodes, dtype=self._dtype).reshape(-1)
        if not self._h:
            _check(L.sctl_cuda_bio_create(self.far._h, self.far._src["Src"], self.far._trg["Trg"], self.near._h if self.near is not None else None, int(self.ndot),
                                          _np_ptr(self._wts), self._wts.size, _np_ptr(self._nrm), 0 if self._nrm is None else self._nrm.size, C.byref(self._h)))
        for st, k in self.far._s2t.items():
            self.far._push_s2t(st, k)
        n = self.Ntrg * self.k1r
        if U is None or not isinstance(U, np.ndarray) or U.size != n or U.dtype != self._dtype or not U.flags.c_contiguous:
            U = np.zeros(n, dtype=self._dtype)
        d = self.far._digits
        _check(L.sctl_cuda_bio_apply(self._h, -1 if d >= 16 else d, _np_ptr(F), F.size, _np_ptr(Ff), 0 if Ff is None else Ff.size, _np_ptr(U.reshape(-1)), n))
        return U

    def last_timing(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _check(load_library().sctl_cuda_bio_last_timing(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"far_ms": a.value, "tail_ms": b.value, "total_ms": c.value}

    def near_shards(self):
        """Work items of the near-field GEMV per GPU of the far-field session (all zero: not sharded)."""
        v = (C.c_int64 * 16)()
        _check(load_library().sctl_cuda_bio_near_shards(self._h, v, 16))
        return [int(x) for x in v][:len(self.far.devices())]

    def compute_potential_staged(self, F, F_far_nodes=None):
        F = np.ascontiguousarray(F, dtype=self._dtype).reshape(-1)
        Fs = F if F_far_nodes is None else np.ascontiguousarray(F_far_nodes, dtype=self._dtype).reshape(-1)
        self.far.SetSrcDensity("Src", (Fs.reshape(-1, self.k0) * self._wts[:, None]).reshape(-1))  # F_far <-- F * wts_far (:1270-1283)
        U = self.far.EvalDirect(None, "Trg", self.k1)
        if self.ndot:  # (:1293-1301)
            U = np.ascontiguousarray((U.reshape(self.Ntrg, self.k1r, 3) * self._nrm.reshape(self.Ntrg, 1, 3)).sum(axis=2).reshape(-1))
        if self.near is not None:
            self.near.apply_accumulate(F, U)
        return U
