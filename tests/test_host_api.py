"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/sctl_cuda.h declares, reports the
reference's kernel metadata, plans grids sanely, and FAILS LOUDLY (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sbm():
    import sctl_b200
    sctl_b200.load_library()
    return sctl_b200


def test_library_exports_every_declared_symbol(sbm):
    header = open(sbm.HEADER_PATH).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(sctl_cuda_\w+)\s*\(", header)))
    assert len(declared) >= 14, declared
    lib = ctypes.CDLL(sbm.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libsctl_cuda.so does not export %s" % name
    # signatures use plain C types only: no torch / C++ types in the header
    assert "torch" not in header and "std::" not in header and "at::" not in header


def test_product_does_not_link_or_import_the_oracle(sbm):
    out = subprocess.run(["ldd", sbm.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "sctl_ref" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "sctl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), f
                assert "liboracle" not in src and "sctl_ref" not in src, f
    for f in os.listdir(os.path.join(ROOT, "include", "sctl_cuda")) + ["../sctl_cuda.h"]:
        src = open(os.path.join(ROOT, "include", "sctl_cuda", f)).read()
        assert "oracle" not in src.lower(), f


def test_kernel_info_matches_oracle_and_python_mirror(sbm, port):
    for kernel in cases.KERNELS:
        info, d = sbm.kernel_info(kernel), port.dims(kernel)
        assert (info["k0"], info["k1"], info["nd"], info["flops"], info["name"]) == (d["k0"], d["k1"], d["nd"], d["flops"], d["name"])
        assert info["scale"] == d["scale64"]
        K = sbm.KERNELS[kernel]
        assert (K.SrcDim(), K.TrgDim(), K.NormalDim(), K.CoordDim(), K.FLOPS(), K.Name()) == (d["k0"], d["k1"], d["nd"], 3, d["flops"], d["name"])
        assert abs(K.uKerScaleFactor(np.float64) - d["scale64"]) < 1e-16
        assert abs(float(K.uKerScaleFactor(np.float32)) - d["scale32"]) < 1e-7
    assert sbm.kernel_info(3)["f_alg"] == 29  # Stokes3D_FxU algorithmic flop per pair
    with pytest.raises(sbm.SctlCudaError):
        sbm.kernel_info(16)


def test_planner_fills_waves(sbm):
    slots, tpb = 148 * 2, 512
    p = sbm.plan_eval(1 << 20, 1 << 20, tpb, slots, 2048, 64)
    assert p["n_tblocks"] == 2048 and p["n_split"] == 1 and p["wave_eff"] > 0.97
    p = sbm.plan_eval(1 << 20, 1 << 17, tpb, slots, 2048, 64)  # the 8-GPU shard of the headline config
    assert p["n_split"] > 1 and p["wave_eff"] > 0.97
    assert p["n_split"] * p["chunk"] >= (1 << 20) and (p["n_split"] - 1) * p["chunk"] < (1 << 20)
    for Ns, Nt in ((1, 1), (5, 100000), (100000, 5), (20000, 20000), (3, 1 << 22)):
        p = sbm.plan_eval(Ns, Nt, tpb, slots, 2048, 64)
        assert p["n_tblocks"] == -(-Nt // tpb) and 1 <= p["n_split"] <= 64
        assert p["n_split"] * p["chunk"] >= Ns and (p["n_split"] - 1) * p["chunk"] < Ns  # chunks tile the sources, none empty


def test_constant_bank_geometry(sbm):
    """ConstGeom (csrc/eval_const_kernel.cuh): slices never overlap, a segment fits its slice, chunks tile the sources,
    the bit row covers every target block, too few chunks fall back to one slice, too many target blocks drop the bits."""
    for kernel in ("Laplace3D_FxU", "Stokes3D_FxU", "Stokes3D_DxU", "BiotSavart3D_FxdU"):
        for rb in (8, 4):
            tpb = 512 if rb == 8 else 512  # CONST_NT * CONST_T: 256 x 2 (FP64), 128 x 4 (FP32)
            for Ns, Nt, banks in ((1 << 20, 1 << 20, 2), (1 << 20, 1 << 17, 4), (5, 40000, 2), (9000, (1 << 18) + 77, 2), (1 << 22, 1 << 22, 2), (3000, 1 << 25, 2)):
                g = sbm.const_geometry(kernel, rb, Ns, Nt, banks)
                assert g["banks"] in (1, banks) and g["banks"] * g["bank_stride"] <= 63 * 1024 and g["bank_stride"] % 16 == 0
                assert g["chunk"] >= 1 and g["chunk"] * g["rec_bytes"] + g["row_words"] * 4 <= g["bank_stride"]
                assert g["seg_bytes"] == g["chunk"] * g["rec_bytes"] + g["row_words"] * 4 and g["seg_bytes"] % 16 == 0
                assert g["n_chunks"] * g["chunk"] >= Ns > (g["n_chunks"] - 1) * g["chunk"]
                assert g["banks"] == 1 or g["n_chunks"] >= 2 * g["banks"]
                n_tb = -(-Nt // tpb)
                assert g["row_words"] == 0 or g["row_words"] * 32 >= n_tb
                assert (g["row_words"] > 0) == (-(-n_tb // 32) * 4 <= g["bank_stride"] // 8)
    g = sbm.const_geometry("Stokes3D_FxU", 8, 1 << 20, 1 << 20, 2)
    assert (g["banks"], g["rec_bytes"], g["row_words"], g["chunk"]) == (2, 48, 64, 666)
    assert sbm.const_geometry("Stokes3D_FxU", 8, 5, 40000, 2)["banks"] == 1
    assert sbm.const_geometry("Stokes3D_FxU", 8, 3000, 1 << 25, 2)["row_words"] == 0


def test_compute_fails_loudly_without_gpu(sbm):
    if sbm.device_count() > 0:
        pytest.skip("a GPU is present")
    K = sbm.Laplace3D_FxU
    with pytest.raises(sbm.SctlCudaError, match="no CUDA device"):
        K.Eval(None, [1.0, 0, 0], [0.0, 0, 0], None, [1.0])
    with pytest.raises(sbm.SctlCudaError, match="no CUDA device"):
        K.KernelMatrix([1.0, 0, 0], [0.0, 0, 0], None)
    with pytest.raises(sbm.SctlCudaError):
        sbm.init(1)
    with pytest.raises(sbm.SctlCudaError):
        sbm.fma_peak(8, 1.0)
    # the session layers have no CPU fallback either
    with pytest.raises(sbm.SctlCudaError, match="no CUDA device"):
        sbm.ParticleDirect(np.float64)
    one = np.array([1], dtype=np.int64)
    with pytest.raises(sbm.SctlCudaError, match="no CUDA device"):
        sbm.NearBlocks(sbm.Laplace3D_FxU, one * 0, one * 0, one, one * 0, one, one * 0)


def test_precondition_violations_raise(sbm):
    K = sbm.Stokes3D_DxU
    with pytest.raises(sbm.SctlCudaError, match="v_src"):
        K.Eval(None, np.zeros(3), np.zeros(6), np.zeros(6), np.zeros(5))  # v_src.Dim() != Ns*KDIM0
    with pytest.raises(sbm.SctlCudaError, match="n_src"):
        K.Eval(None, np.zeros(3), np.zeros(6), np.zeros(3), np.zeros(6))  # n_src.Dim() != Ns*N_DIM
    with pytest.raises(sbm.SctlCudaError, match="r_src"):
        sbm.Laplace3D_FxU.Eval(None, np.zeros(3), np.zeros(4), None, np.zeros(1))
    L = sbm.load_library()
    z = np.zeros(8)
    p = z.ctypes.data_as(ctypes.c_void_p)
    assert L.sctl_cuda_eval(99, 8, -1, 1, 1, p, p, p, p, p, None, 0) != 0 and b"kernel id" in L.sctl_cuda_last_error()
    assert L.sctl_cuda_eval(0, 2, -1, 1, 1, p, p, p, p, p, None, 0) != 0 and b"real_bytes" in L.sctl_cuda_last_error()
    assert L.sctl_cuda_eval(0, 8, -1, -1, 1, p, p, p, p, p, None, 0) != 0 and b"negative" in L.sctl_cuda_last_error()


def test_cpp_host_layer_compiles_standalone(tmp_path):
    """include/sctl_cuda/*.hpp is self-contained C++17 (no SCTL, no CUDA headers needed by callers)."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "sctl_cuda/generic_kernel.hpp"\n#include "sctl_cuda/mini_containers.hpp"\n#include "sctl_cuda/bio_apply.hpp"\n'
                   "int main(){ sctl_cuda::Stokes3D_FxU K; static_assert(decltype(K)::SrcDim()==3 && decltype(K)::TrgDim()==3 && decltype(K)::NormalDim()==0, \"dims\");\n"
                   " return (K.Name()==\"Stokes3D-FxU\" && K.FLOPS()==23) ? 0 : 1; }\n")
    exe = tmp_path / "t"
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", os.path.join(ROOT, "sctl_b200"), "-lsctl_cuda", "-Wl,-rpath," + os.path.join(ROOT, "sctl_b200")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0


def test_sass_evidence_of_the_b200_design(sbm):
    """cuobjdump of the shipped library: the shared-memory kernels stream sources with TMA bulk copies (UBLKCP) behind
    mbarriers (SYNCS), seed the reciprocal square root with MUFU.RSQ64H, and the constant-bank kernels feed FP64
    instructions with uniform-register operands loaded by LDCU (no vector LDC of the source records), and the FP32 pair loops
    use the packed FP32 instructions with broadcast scalar operands; the KernelMatrix kernels store 16-byte vectors and keep
    CTA barriers out of their row loop."""
    import shutil
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    txt = subprocess.run(["cuobjdump", "-sass", sbm.LIB_PATH], capture_output=True, text=True).stdout
    funcs = {m.group(1): m.group(2) for m in re.finditer(r"Function : (\S+)\n(.*?)(?=Function :|\Z)", txt, flags=re.S)}
    smem = [b for n, b in funcs.items() if "eval_pairs_kernel" in n]
    const = [b for n, b in funcs.items() if "eval_pairs_const_kernel" in n]
    helmholtz = {id(b) for n, b in funcs.items() if re.search(r"eval_pairs_const_kernelI[df]Li1[2-5]E", n)}  # kernel ids 12-15
    assert len(smem) >= 50 and len(const) >= 25
    for body in smem:
        assert "UBLKCP" in body and "SYNCS" in body
    for n, body in funcs.items():
        if "eval_pairs" in n and "IdLi" in n:
            assert "MUFU.RSQ64H" in body, n
    for body in const:
        assert re.search(r"LDCU(\.\d+)? UR\d+, c\[0x3\]", body)
        # the only vector loads from the bank are the two one-off reads of the masked/unmasked decision (the probe's flags
        # and this CTA's coincidence-bit word), never the source records inside the loops
        # (the Helmholtz kernels also load sincos polynomial coefficients from the bank: exempt)
        assert id(body) in helmholtz or len(re.findall(r" LDC(\.\d+)? R\d+, c\[0x3\]", body)) <= 2
        assert len(re.findall(r"[DF](FMA|MUL|ADD)2? [^;]*UR\d+", body)) >= 40
    # FP32 pair loops are packed (two targets per 64-bit register pair): FFMA2 / FMUL2 / FADD2 with the source value as a
    # broadcast scalar operand -- a uniform register on the constant-bank path, an LDS result on the ring path
    for n, body in funcs.items():
        if re.search(r"eval_pairs_const_kernelIfLi\d+ELi4E", n):
            assert len(re.findall(r"F(FMA|MUL|ADD)2 [^;]*UR\d+\.F32", body)) >= 10, n
        if re.search(r"eval_pairs_kernelIfLi\d+ELi4E", n):  # T = 4 (main / wide); the few-target variant has T = 1
            assert len(re.findall(r"F(FMA|MUL|ADD)2 [^;]* -?R\d+\.F32", body)) >= 10, n
    # KernelMatrix (csrc/kmat_kernel.cuh): every instantiation writes M with 16-byte vector stores, the staged ones move
    # their stage with 16-byte shared-memory loads, and the only CTA barriers are the two around a source tile (one BAR.SYNC
    # site per __syncthreads() call: two in the kernel, none added by the row loop)
    kmat = {n: b for n, b in funcs.items() if "kernel_matrix_kernel" in n}
    # 14 K1 > 1 kernels x (3 FP64 + 2 FP32 digits classes) staged; the two K1 == 1 kernels in the direct and the lane-shifted form
    assert len(kmat) == 14 * 5 + 2 * 5 * 2
    for n, body in kmat.items():
        mode = int(re.search(r"kernel_matrix_kernelI[df]Li\d+ELi\dELi(\d)EEE", n).group(1))  # KMAT_STAGED / _DIRECT / _SHIFTED
        assert "STG.E.128" in body, n
        assert len(re.findall(r"BAR\.SYNC", body)) <= 2, n
        if mode == 0:
            assert "LDS.128" in body and "WARPSYNC" in body, n
        if mode == 2:
            assert "SHFL.DOWN" in body, n
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", sbm.LIB_PATH], capture_output=True, text=True).stdout


def test_planners_hold_their_invariants_on_random_sizes(sbm):
    """Property test (hypothesis) of the two pieces of host logic that decide how sources are cut: whatever the sizes, chunks
    tile the sources exactly (no source dropped or counted twice, no empty chunk) and the constant-bank layout stays inside
    its slice of the 64 KB bank."""
    from hypothesis import given, settings, strategies as st

    sizes = st.one_of(st.integers(1, 5000), st.integers(1, 1 << 24))

    @settings(max_examples=300, deadline=None)
    @given(Ns=sizes, Nt=sizes, tpb=st.sampled_from([32, 256, 512, 1024, 2048]), slots=st.sampled_from([148, 296, 592]),
           min_tiles=st.integers(1, 8), max_split=st.sampled_from([1, 8, 64, 1024]))
    def ring(Ns, Nt, tpb, slots, min_tiles, max_split):
        p = sbm.plan_eval(Ns, Nt, tpb, slots, 256 * min_tiles, max_split)
        assert p["n_tblocks"] == -(-Nt // tpb)
        assert 1 <= p["n_split"] <= max_split and p["chunk"] >= 1
        assert p["n_split"] * p["chunk"] >= Ns > (p["n_split"] - 1) * p["chunk"]
        assert 0.0 < p["wave_eff"] <= 1.0 + 1e-12

    @settings(max_examples=300, deadline=None)
    @given(Ns=sizes, Nt=sizes, rb=st.sampled_from([8, 4]), banks=st.integers(1, 4),
           kernel=st.sampled_from(["Laplace3D_FxU", "Laplace3D_DxU", "Stokes3D_FxU", "Stokes3D_DxU", "Stokes3D_FxT", "Helmholtz3D_DxU"]))
    def bank(Ns, Nt, rb, banks, kernel):
        g = sbm.const_geometry(kernel, rb, Ns, Nt, banks)
        assert 1 <= g["banks"] <= banks and g["banks"] * g["bank_stride"] <= 63 * 1024 and g["bank_stride"] % 16 == 0
        assert g["chunk"] >= 1 and g["seg_bytes"] == g["chunk"] * g["rec_bytes"] + g["row_words"] * 4 <= g["bank_stride"]
        assert g["n_chunks"] * g["chunk"] >= Ns > (g["n_chunks"] - 1) * g["chunk"]
        assert g["row_words"] == 0 or g["row_words"] * 32 >= -(-Nt // 512)

    ring()
    bank()


def test_near_slice_planner_tiles_the_work_list(sbm):
    """Property test (hypothesis) of the host logic that shards the near-field GEMV of a multi-GPU apply (csrc/near_session.inc
    plan_near_slices): whatever the element sizes, the slices tile the work list exactly and in order, every item's rows of
    K and U_near lie inside its slice's ranges, the U_near ranges of different slices do not overlap, and the bytes of K are
    balanced to within one item; a block list whose U_near ranges do not follow the work-list order is refused."""
    from hypothesis import given, settings, strategies as st

    def block_table(nds, near, blk, k0, k1):
        # BoundaryIntegralOp's layout (boundary_integral.txx:1009-1032): per element, blocks of at most blk near targets
        nds, near = np.asarray(nds), np.asarray(near)
        near_dsp = np.concatenate([[0], np.cumsum(near)[:-1]])
        K_dsp = np.concatenate([[0], np.cumsum(nds * near)[:-1]])
        ns, u_dsp, nt, k_dsp = [], [], [], []
        for e in range(len(nds)):
            for t0 in range(0, near[e], blk):
                ns.append(nds[e]); u_dsp.append(near_dsp[e] + t0); nt.append(min(blk, near[e] - t0)); k_dsp.append(K_dsp[e] * k0 * k1 + t0 * k1 * nds[e] * k0)
        return [np.array(c, dtype=np.int64) for c in (ns, u_dsp, nt, k_dsp)]

    @settings(max_examples=200, deadline=None)
    @given(elems=st.lists(st.tuples(st.integers(0, 70), st.integers(0, 150)), min_size=1, max_size=40), blk=st.integers(1, 64),
           dims=st.sampled_from([(1, 1), (3, 3), (1, 3), (2, 6)]), P=st.integers(1, 8), rb=st.sampled_from([8, 4]))
    def check(elems, blk, dims, P, rb):
        k0, k1 = dims
        ns, u_dsp, nt, k_dsp = block_table([e[0] for e in elems], [e[1] for e in elems], blk, k0, k1)
        slices, n_items, ok = sbm.plan_near_slices(rb, k0, k1, ns, u_dsp, nt, k_dsp, P)
        assert ok
        rows_per_item = 32 if rb == 8 else 64
        items = [(b, r) for b in range(len(nt)) for r in range(0, int(nt[b]) * k1, rows_per_item)]
        assert n_items == len(items)
        assert slices[0]["i0"] == 0 and slices[-1]["i1"] == n_items and all(slices[p]["i1"] == slices[p + 1]["i0"] for p in range(P - 1))
        total = sum((min(int(nt[b]) * k1, r + rows_per_item) - r) * int(ns[b]) * k0 for b, r in items)
        biggest = max([(min(int(nt[b]) * k1, r + rows_per_item) - r) * int(ns[b]) * k0 for b, r in items] + [0])
        prev_uhi = 0
        for p, sl in enumerate(slices):
            mine = items[sl["i0"]:sl["i1"]]
            size = 0
            for b, r in mine:
                r1 = min(int(nt[b]) * k1, r + rows_per_item)
                cols = int(ns[b]) * k0
                assert sl["klo"] <= int(k_dsp[b]) + r * cols and int(k_dsp[b]) + r1 * cols <= sl["khi"]
                assert sl["ulo"] <= int(u_dsp[b]) * k1 + r and int(u_dsp[b]) * k1 + r1 <= sl["uhi"]
                size += (r1 - r) * cols
            if mine:
                assert sl["ulo"] >= prev_uhi
                prev_uhi = sl["uhi"]
            assert size <= total / P + biggest + 1

    check()
    # two blocks whose U_near ranges run against the work-list order: refused, nothing sharded
    ns, u_dsp, nt, k_dsp = [np.array(c, dtype=np.int64) for c in ([5, 5], [40, 0], [40, 40], [0, 200])]
    slices, n_items, ok = sbm.plan_near_slices(8, 1, 1, ns, u_dsp, nt, k_dsp, 2)
    assert not ok and all(sl["i1"] == sl["i0"] == 0 for sl in slices)
