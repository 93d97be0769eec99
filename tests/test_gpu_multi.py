"""Multi-GPU tests (need >= 2 visible GPUs; skipped otherwise): the C library's own single-process sharding
(sctl_cuda_init(n): target shards + NCCL source replication, csrc/sctl_cuda.cu) and the torchrun layer
(sctl_b200.sharded over NCCL) must reproduce the single-GPU result to rounding -- no reduction between GPUs is involved, so the
result agrees to rounding for any number of GPUs (a shard may split the sources differently from the whole target set)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from conftest import rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_gpus(sb, n):
    if sb.device_count() < n:
        pytest.skip("needs %d GPUs, %d visible" % (n, sb.device_count()))


@pytest.mark.parametrize("kernel", ["Stokes3D_FxU", "Laplace3D_DxU"])
def test_library_sharding_matches_single_gpu(sb, cpu_oracle, kernel):
    _need_gpus(sb, 2)
    from sctl_b200 import synth
    k0, k1, nd = cases.KERNELS[kernel]
    K = sb.KERNELS[kernel]
    Ns, Nt = 60001, 40003  # > 1e9 pairs: the host entry shards; ragged so the block partition is uneven
    r_src, n_src, v_src, r_trg = synth.cube(Ns, Nt, k0, nd)
    pre = synth.splitmix_u01(17, Nt * k1) - 0.5
    try:
        assert sb.init(1) == 1
        one = K.Eval(pre.copy(), r_trg, r_src, n_src, v_src)
        n = sb.init(0)  # all visible GPUs
        assert n == sb.device_count() >= 2
        many = K.Eval(pre.copy(), r_trg, r_src, n_src, v_src)
        t = sb.last_timing()
        assert t["kernel_ms"] > 0
    finally:
        sb.init(1)
    idx = np.arange(64) * (Nt // 64)
    want = cpu_oracle.eval(kernel, r_src, n_src, v_src, r_trg.reshape(-1, 3)[idx].reshape(-1), v_trg=pre.reshape(-1, k1)[idx].reshape(-1))
    assert rel_err(many.reshape(-1, k1)[idx].reshape(-1), want) <= 1e-12
    assert rel_err(many, one) <= 1e-13  # shards may pick a different source split than the full problem


def test_torchrun_sharded_eval_two_ranks(sb):
    _need_gpus(sb, 2)
    script = os.path.join(ROOT, "tests", "_torchrun_sharded.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29631", script],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_particle_direct_two_gpus(sb, cpu_oracle):
    """sctl_cuda_direct_* with the targets sharded over 2 GPUs reproduces the single-GPU session to 1e-13 (no
    reduction between GPUs) and the CPU oracle; the C++ mirror's own test program runs with 2 GPUs as well."""
    _need_gpus(sb, 2)
    from sctl_b200 import synth
    Ns, Nt = 7001, 9007
    r_src, n_src, v_src, r_trg = synth.cube(Ns, Nt, 3, 3)
    res = []
    for n in (1, 2):
        pd = sb.ParticleDirect(np.float64, n_gpus=n)
        pd.AddSrc("S"); pd.AddTrg("T")
        pd.SetKernelS2T("S", "T", sb.Stokes3D_DxU)
        pd.SetSrcCoord("S", r_src, n_src); pd.SetTrgCoord("T", r_trg); pd.SetSrcDensity("S", v_src)
        pd.SetAccuracy(-1)
        res.append(pd.EvalDirect(None, "T", 3))
        pd.close()
    assert rel_err(res[0], res[1]) <= 1e-13  # a shard may pick a different source split than the full target set
    assert rel_err(res[1], cpu_oracle.eval("Stokes3D_DxU", r_src, n_src, v_src, r_trg)) <= 1e-12
    exe = os.path.join(ROOT, "tests", "_bin", "test_particle_direct")
    if os.path.exists(exe):
        r = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "PASS (0 check(s) failed)" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_particle_direct_follows_the_current_device(sb, cpu_oracle):
    """One rank per GPU (torchrun, MPI): a session created while device 1 is current must live on device 1 -- it used to
    be hard-wired to device 0, which put every rank's far field on GPU 0."""
    _need_gpus(sb, 2)
    import torch
    from sctl_b200 import synth
    Ns, Nt = 3001, 2003
    r_src, _, v_src, r_trg = synth.cube(Ns, Nt, 3, 0)
    prev = torch.cuda.current_device()
    try:
        torch.cuda.set_device(1)
        free0 = torch.cuda.mem_get_info(0)[0]
        pd = sb.ParticleDirect(np.float64)
        assert pd.devices() == [1]
        pd.AddSrc("S"); pd.AddTrg("T"); pd.SetAccuracy(-1)
        pd.SetKernelS2T("S", "T", sb.Stokes3D_FxU)
        big = np.zeros(3 * (1 << 24))  # 384 MB of target coordinates: must land on device 1
        pd.AddTrg("big"); pd.SetTrgCoord("big", big)
        assert torch.cuda.mem_get_info(0)[0] >= free0 - (64 << 20)
        pd.SetSrcCoord("S", r_src); pd.SetTrgCoord("T", r_trg); pd.SetSrcDensity("S", v_src)
        got = pd.EvalDirect(None, "T", 3)
        assert torch.cuda.current_device() == 1  # the caller's device is left alone
        pd.close()
        two = sb.ParticleDirect(np.float64, devices=[1, 0])
        assert two.devices() == [1, 0]
        two.close()
    finally:
        torch.cuda.set_device(prev)
    assert rel_err(got, cpu_oracle.eval("Stokes3D_FxU", r_src, None, v_src, r_trg)) <= 1e-12


def test_bio_apply_pipeline_two_gpus(sb, cpu_oracle):
    """The whole BoundaryIntegralOp apply with the far field sharded over 2 GPUs: the shards are collected on the first GPU
    (peer copies), the near-field GEMV is sharded too and the accumulation runs on the first GPU; agrees with the 1-GPU
    pipeline to rounding and with the staged apply bit for bit."""
    _need_gpus(sb, 2)
    from sctl_b200 import synth
    from test_gpu_near import block_list, near_target_tables
    kernel, k0, k1 = "Stokes3D_FxU", 3, 3
    tab, n_nodes, n_far, n_near, K_len = block_list(60, k0, k1, seed=11)
    Ntrg = 40003  # ragged shards
    cnt, dsp, scatter = near_target_tables(Ntrg, n_near, seed=3)
    X_far, Xn_far, _, Xtrg = synth.cube(n_nodes, Ntrg, k0, 3)
    wts = 0.01 + 0.001 * synth.splitmix_u01(32, n_nodes)
    K_near = synth.splitmix_u01(22, K_len) - 0.5
    F = synth.splitmix_u01(41, n_nodes * k0) - 0.5
    res = []
    for n in (1, 2):
        near = sb.NearBlocks(sb.KERNELS[kernel], *tab)
        near.set_matrix(K_near)
        near.set_targets(cnt, dsp, scatter)
        bio = sb.BioApply(sb.KERNELS[kernel], X_far, Xn_far, wts, Xtrg, None, digits=-1, n_gpus=n, near=near)
        res.append(bio.compute_potential(F))
        assert np.array_equal(res[-1], bio.compute_potential_staged(F))
        # the near-field GEMV is sharded with the far field: every GPU runs a slice of its work list (same kernel and items,
        # so the staged apply above, which runs all of it on the first GPU, gives the same bits)
        shards = bio.near_shards()
        assert len(shards) == n and (all(x > 0 for x in shards) if n > 1 else shards == [0]), shards
        # a new matrix in the near session reaches the other GPUs' copies of their slices
        near.set_matrix(2.0 * K_near)
        assert np.array_equal(bio.compute_potential(F), bio.compute_potential_staged(F))
        near.set_matrix(K_near)
        assert np.array_equal(bio.compute_potential(F), res[-1])
        bio.close()
        near.close()
    assert rel_err(res[0], res[1]) <= 1e-13
