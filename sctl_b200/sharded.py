"""Multi-GPU layer for one process per GPU (torchrun): target shards + source replication over NCCL.

Reference analogue: ``ParticleFMM::EvalDirect`` (include/sctl/fmm-wrapper.txx:551-623) -- contiguous target blocks
per rank (:565-569) and an np-step source ring (:598-619) that serialises communication and compute.  On an
NVSwitch box every GPU reaches every peer at full bandwidth, so the ring is replaced by one round of broadcasts:
rank q broadcasts its source slice into the same offsets of every rank's full source arrays (an all-gather with
ragged counts), then every rank evaluates its own target shard against all sources.  No reduction is needed, so the
per-target sum involves no cross-GPU partial sums; results agree to rounding for any number of ranks (a shard may
split the sources differently from the whole target set).

``torch.distributed`` is plumbing only; the arithmetic goes through ``sctl_cuda_eval_device``.  The compute step is
injectable so the partition / exchange logic can be tested with the gloo backend on CPU (tests/test_sharded_gloo.py).
"""
import torch
import torch.distributed as dist

from . import KERNEL_IDS, GenericKernel, eval_device


def block_range(n, rank, world):
    """Contiguous block partition [n*rank/world, n*(rank+1)/world) -- the rule of fmm-wrapper.txx:568."""
    return (n * rank) // world, (n * (rank + 1)) // world


def _meta(kernel):
    k = kernel if isinstance(kernel, GenericKernel) else GenericKernel(kernel)
    return k.kernel_id, k.SrcDim(), k.TrgDim(), k.NormalDim()


def _cuda_compute(kernel_id, r_src, n_src, v_src, r_trg, v_trg, digits, ctx=None):
    eval_device(kernel_id, r_src, n_src, v_src, r_trg, v_trg, digits=digits, ctx=ctx)


class ShardedEval:
    """Holds the replicated-source buffers of one rank and runs the sharded Eval.

    Ns, Nt are GLOBAL counts.  Each rank passes its own source slice (block_range(Ns, rank, world)) and its own
    target shard (block_range(Nt, rank, world)); ``eval`` returns the rank's slice of v_trg (accumulated into
    ``v_trg_local`` when given).  ``gather`` assembles the full result on every rank.
    """

    def __init__(self, kernel, Ns, Nt, dtype=torch.float64, device=None, group=None, compute=None, ctx=None):
        self.kernel_id, self.k0, self.k1, self.nd = _meta(kernel)
        self.Ns, self.Nt, self.dtype, self.group = int(Ns), int(Nt), dtype, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
        self.compute = compute or (lambda *a: _cuda_compute(*a, ctx=ctx))
        self.s0, self.s1 = block_range(self.Ns, self.rank, self.world)
        self.t0, self.t1 = block_range(self.Nt, self.rank, self.world)
        new = lambda n: torch.empty(n, dtype=dtype, device=self.device)
        self.r_src = new(self.Ns * 3)
        self.v_src = new(self.Ns * self.k0)
        self.n_src = new(self.Ns * self.nd) if self.nd else None

    def replicate(self, r_src_local, n_src_local, v_src_local):
        """Place this rank's slice and broadcast every rank's slice into the full arrays (async, one wait)."""
        s0, s1 = self.s0, self.s1
        self.r_src[s0 * 3:s1 * 3].copy_(r_src_local, non_blocking=True)
        self.v_src[s0 * self.k0:s1 * self.k0].copy_(v_src_local, non_blocking=True)
        if self.nd:
            self.n_src[s0 * self.nd:s1 * self.nd].copy_(n_src_local, non_blocking=True)
        if self.world == 1:
            return
        work = []
        for q in range(self.world):
            q0, q1 = block_range(self.Ns, q, self.world)
            if q1 == q0:
                continue
            src = dist.get_global_rank(self.group, q) if self.group is not None else q
            work.append(dist.broadcast(self.r_src[q0 * 3:q1 * 3], src=src, group=self.group, async_op=True))
            work.append(dist.broadcast(self.v_src[q0 * self.k0:q1 * self.k0], src=src, group=self.group, async_op=True))
            if self.nd:
                work.append(dist.broadcast(self.n_src[q0 * self.nd:q1 * self.nd], src=src, group=self.group, async_op=True))
        for w in work:
            w.wait()

    def eval(self, r_src_local, n_src_local, v_src_local, r_trg_local, v_trg_local=None, digits=-1):
        nt = self.t1 - self.t0
        assert r_trg_local.numel() == nt * 3, "r_trg_local must be this rank's target shard"
        self.replicate(r_src_local, n_src_local, v_src_local)
        if v_trg_local is None or v_trg_local.numel() != nt * self.k1:
            v_trg_local = torch.zeros(nt * self.k1, dtype=self.dtype, device=self.device)
        if nt and self.Ns:
            self.compute(self.kernel_id, self.r_src, self.n_src, self.v_src, r_trg_local, v_trg_local, digits)
        return v_trg_local

    def gather(self, v_trg_local):
        """All ranks get the full v_trg[Nt*K1] (ragged all-gather by broadcasts)."""
        full = torch.empty(self.Nt * self.k1, dtype=self.dtype, device=self.device)
        full[self.t0 * self.k1:self.t1 * self.k1].copy_(v_trg_local)
        if self.world == 1:
            return full
        work = []
        for q in range(self.world):
            q0, q1 = block_range(self.Nt, q, self.world)
            if q1 == q0:
                continue
            src = dist.get_global_rank(self.group, q) if self.group is not None else q
            work.append(dist.broadcast(full[q0 * self.k1:q1 * self.k1], src=src, group=self.group, async_op=True))
        for w in work:
            w.wait()
        return full
