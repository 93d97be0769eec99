"""CPU test of the N>1 path: world_size-2 (and 3) gloo groups run the partition + source-replication + gather logic of
sctl_b200.sharded with the oracle injected as the compute step; the assembled result must equal a single-rank run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kernel, Ns, Nt, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from sctl_b200 import synth
    from sctl_b200.sharded import ShardedEval, block_range
    P = oracle.PortOracle()

    def compute(kid, r_src, n_src, v_src, r_trg, v_trg, digits):  # oracle as the checker-side compute step
        res = P.eval(kid, r_src.numpy(), None if n_src is None else n_src.numpy(), v_src.numpy(), r_trg.numpy(), v_trg=v_trg.numpy(), omp=False)
        v_trg.copy_(torch.from_numpy(res))

    sh = ShardedEval(kernel, Ns, Nt, dtype=torch.float64, device=torch.device("cpu"), compute=compute)
    k0, k1, nd = sh.k0, sh.k1, sh.nd
    s0, s1 = block_range(Ns, rank, world)
    t0, t1 = block_range(Nt, rank, world)
    # every rank generates ONLY its own slices (counter-based generator)
    r_src, n_src, v_src, r_trg = synth.cube(s1 - s0, t1 - t0, k0, nd, first_src=s0, first_trg=t0)
    tt = lambda a: None if a is None else torch.from_numpy(a)
    pre = torch.full(((t1 - t0) * k1,), 0.25, dtype=torch.float64)  # accumulate semantics survive sharding
    v_loc = sh.eval(tt(r_src), tt(n_src), tt(v_src), tt(r_trg), v_trg_local=pre)
    full = sh.gather(v_loc)
    assert torch.equal(sh.r_src, torch.from_numpy(synth.cube(Ns, 0, k0, nd)[0]))  # replication is exact
    np.save(os.path.join(out_dir, "full_%d.npy" % rank), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kernel,Ns,Nt", [(2, "Stokes3D_FxU", 301, 257), (2, "Laplace3D_DxU", 64, 5), (3, "Stokes3D_DxU", 100, 100), (2, "Laplace3D_FxU", 1, 1)])
def test_sharded_matches_single_rank(tmp_path, world, kernel, Ns, Nt):
    sys.path.insert(0, ROOT)
    import oracle
    from sctl_b200 import synth
    port = _free_port()
    mp.spawn(_worker, args=(world, port, kernel, Ns, Nt, str(tmp_path)), nprocs=world, join=True)
    P = oracle.PortOracle()
    k0, k1, nd = oracle.KERNEL_DIMS[oracle.KERNEL_IDS[kernel]]
    r_src, n_src, v_src, r_trg = synth.cube(Ns, Nt, k0, nd)
    want = P.eval(kernel, r_src, n_src, v_src, r_trg, v_trg=np.full(Nt * k1, 0.25), omp=False)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "full_%d.npy" % r))
        assert np.array_equal(got, want), (r, np.abs(got - want).max())  # same per-target order => bit-identical


def test_block_range_covers_everything():
    sys.path.insert(0, ROOT)
    from sctl_b200.sharded import block_range
    for n in (0, 1, 7, 8, 1000, 2 ** 20 + 3):
        for world in (1, 2, 3, 8):
            edges = [block_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
