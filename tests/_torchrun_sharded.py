"""Worker for tests/test_gpu_multi.py (launched by torchrun, one rank per GPU, NCCL)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (tests may use the checker)
import sctl_b200 as sb  # noqa: E402
from sctl_b200 import synth  # noqa: E402
from sctl_b200.sharded import ShardedEval, block_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    for kernel, Ns, Nt in (("Stokes3D_DxU", 30011, 50021), ("Laplace3D_FxU", 5, 70001)):
        K = sb.KERNELS[kernel]
        k0, k1, nd = K.SrcDim(), K.TrgDim(), K.NormalDim()
        s0, s1 = block_range(Ns, rank, world)
        t0, t1 = block_range(Nt, rank, world)
        r_src, n_src, v_src, r_trg = synth.cube(s1 - s0, t1 - t0, k0, nd, first_src=s0, first_trg=t0)
        td = lambda a: None if a is None else torch.from_numpy(a).to(dev)
        sh = ShardedEval(kernel, Ns, Nt, device=dev)
        v = sh.eval(td(r_src), td(n_src), td(v_src), td(r_trg))
        full = sh.gather(v).cpu().numpy()
        if rank == 0:
            R, N, V, T = synth.cube(Ns, Nt, k0, nd)
            idx = np.arange(50) * (Nt // 50)
            want = oracle.PortOracle().eval(kernel, R, N, V, T.reshape(-1, 3)[idx].reshape(-1))
            got = full.reshape(-1, k1)[idx].reshape(-1)
            err = float(np.abs(got - want).max() / np.abs(want).max())
            assert err <= 1e-12, (kernel, err)
            single = torch.zeros(Nt * k1, dtype=torch.float64, device=dev)
            sb.eval_device(sb.KERNELS[kernel], td(R), td(N), td(V), td(T), single)
            torch.cuda.synchronize()
            assert float(np.abs(single.cpu().numpy() - full).max() / np.abs(full).max()) <= 1e-13
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
