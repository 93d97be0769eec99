#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: FP64 pair-interactions/s of Stokes3D_FxU Eval at Ns = Nt = 2^20.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one full Eval (all 2^20 targets against all 2^20 sources).
  value     pairs/s with inputs resident in HBM: per step the timed region is, on every rank, source replication
            (NCCL broadcasts, N > 1) + source packing + the pair kernel (+ the partial-sum reduce) + the result
            gather (N > 1); CUDA events on the launching stream, max over ranks, summed over K steps.
  e2e       the same metric through the host-facing API with HOST (pinned) buffers: H2D of that rank's inputs,
            the work above, D2H of the result, every step.
  roofline  the pair kernel against the FP64 CUDA-core peak measured in this run by a DFMA micro-benchmark
            (MEASURED_PEAKS.json has no FP64 entry); achieved = pairs x F_alg(29 flop) / kernel time.
            roofline.hbm_view: the same kernel in the contract's HBM form (it is not the binding roof);
            roofline.kernel_matrix_view (N = 1): the path's HBM-bound kernel, KernelMatrix of the same micro-kernel at
            8192 x 8192, bytes written / time against MEASURED_PEAKS.json's copy bandwidth, measured after the timed region.
  cpu_baseline  the reference's own OpenMP/AVX Eval (oracle/_ref, the unmodified reference compiled by
            oracle/Makefile; falls back to the C port) on all sources x a strided target subsample, on the box's
            host cores, rank 0, N = 1 only.  The same call provides the parity figure printed in the line.
--impl reference times that CPU arm alone on the same workload (bounded sample per step) and prints the same line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64_pair_interactions_per_s"  # BASELINE.json metric; "fp32_..." when run with --real f32
UNIT = "pairs/s"
F_ALG = {"Laplace3D_FxU": 11, "Laplace3D_DxU": 19, "Laplace3D_FxdU": 18, "Stokes3D_FxU": 29, "Stokes3D_DxU": 30}
SCTL_FLOPS = {"Laplace3D_FxU": 6, "Laplace3D_DxU": 14, "Laplace3D_FxdU": 11, "Stokes3D_FxU": 23, "Stokes3D_DxU": 26}  # uKernel::FLOPS()
FP_INSTR = {"Laplace3D_FxU": 12, "Laplace3D_DxU": 17, "Laplace3D_FxdU": 17, "Stokes3D_FxU": 22, "Stokes3D_DxU": 25}  # FP-pipe instructions per pair (f64)
# FP32 at digits -1: the MUFU.RSQ seed alone (XU pipe), so the 5 refinement instructions go; packed two pairs per instruction,
# which halves the issue slots but not the pipe cycles
FP_INSTR32 = {k: v - 5 for k, v in FP_INSTR.items()}
DIMS = {"Laplace3D_FxU": (1, 1, 0), "Laplace3D_DxU": (1, 1, 3), "Laplace3D_FxdU": (1, 3, 0), "Stokes3D_FxU": (3, 3, 0), "Stokes3D_DxU": (3, 3, 3)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # samples under load = upper half (the sampler also sees the gaps between steps)
        load = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


def hbm_peak_gbs():
    """Measured HBM copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the profiling recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback of /opt/skills/guides/B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host thread this process may
    run on, so the OpenMP runtime the checker links (libgomp) is told so directly."""
    import ctypes
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(n)  # read by libgomp if it is not loaded yet
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def cpu_arm(kernel, N, nsub, steps, warmup, want_result=False, dtype=np.float64, points="cube", log2n=0):
    """The reference's CPU Eval on all N sources x nsub strided targets; returns (pairs/s, info, subset result)."""
    import oracle
    from sctl_b200 import synth
    if os.environ.get("OMP_NUM_THREADS") == "1" and "TORCHELASTIC_RUN_ID" in os.environ:
        use_all_host_threads()
    orc = oracle.best_cpu_oracle()
    k0, k1, nd = DIMS[kernel]
    r_src, n_src, v_src, r_trg = synth.cube(N, N, k0, nd, dtype=dtype)
    if points == "torus":
        X, Nrm, _, _ = synth.torus(1 << ((log2n + 1) // 2), 1 << (log2n // 2), dtype=dtype)
        r_src, r_trg, n_src = X, X, (Nrm if nd else None)
    idx = np.arange(nsub) * (N // nsub)
    sub = np.ascontiguousarray(r_trg.reshape(-1, 3)[idx]).reshape(-1)
    for _ in range(max(warmup, 2)):  # the reference's first calls pay OpenMP start-up and its scratch-pool first touch
        orc.eval(kernel, r_src, n_src, v_src, sub, dtype=dtype)
    times, res = [], None
    for _ in range(steps):
        t0 = time.perf_counter()
        res = orc.eval(kernel, r_src, n_src, v_src, sub, dtype=dtype)
        times.append(time.perf_counter() - t0)
    pairs = float(N) * nsub
    info = {"kind": orc.kind, "cores": orc.threads(), "sample": "all %d sources x %d strided targets (%.3g pairs) per step" % (N, nsub, pairs),
            "isa": getattr(orc, "isa", "portable-C")}
    return pairs * steps / sum(times), info, (idx, res) if want_result else None, sum(times) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", default="Stokes3D_FxU", choices=list(DIMS))
    ap.add_argument("--log2n", type=int, default=20, help="Ns = Nt = 2^log2n (BASELINE: 20)")
    ap.add_argument("--digits", type=int, default=-1)
    ap.add_argument("--points", default="cube", choices=["cube", "torus"],
                    help="cube: sources and targets uniform in the unit cube, distinct sets (BASELINE); torus: surface points, targets == sources (every target has one r2 == 0 pair)")
    ap.add_argument("--real", default="f64", choices=["f64", "f32"], help="arithmetic type (BASELINE metric: f64; configs[3] also names f32)")
    ap.add_argument("--cpu-targets", type=int, default=1 << 14, help="strided target subsample for the CPU baseline / parity")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-kmat-view", action="store_true", help="skip roofline.kernel_matrix_view (the KernelMatrix write-bandwidth measurement after the timed region)")
    args = ap.parse_args()

    # stdout carries exactly one JSON line (rank 0): everything else that libraries print to fd 1 (NCCL's version banner
    # when the box sets NCCL_DEBUG=VERSION, torch warnings) is sent to stderr from here on
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N = 1 << args.log2n
    kernel = args.kernel
    k0, k1, nd = DIMS[kernel]
    global METRIC
    METRIC = "fp64_pair_interactions_per_s" if args.real == "f64" else "fp32_pair_interactions_per_s"
    npdt = np.float64 if args.real == "f64" else np.float32
    rb = 8 if args.real == "f64" else 4
    workload = "%s GenericKernel::Eval %s, Ns=Nt=2^%d uniform in unit cube%s" % (kernel, "FP64" if rb == 8 else "FP32", args.log2n,
                                                                                 " (BASELINE configs[2])" if (kernel, args.log2n, rb) == ("Stokes3D_FxU", 20, 8) else "")
    if args.points == "torus":
        workload = workload.replace("uniform in unit cube", "torus surface nodes (R=1, a=0.25, %d x %d grid), targets == sources" % (1 << ((args.log2n + 1) // 2), 1 << (args.log2n // 2)))
    config = {"workload": workload, "kernel": kernel, "Ns": N, "Nt": N, "digits": args.digits, "inputs": "SplitMix64 streams (SURVEY 8d), generated in f64%s" % ("" if rb == 8 else ", cast to f32"),
              "sharding": "targets by contiguous block over %d GPU(s), sources replicated by NCCL broadcast" % args.gpus,
              "l2": "a 256 MiB buffer is overwritten between timed steps (L2 flush)"}

    if args.impl == "reference":
        if rank != 0:
            return
        nsub = min(N, 1 << 13)
        val, info, _, sec = cpu_arm(kernel, N, nsub, args.steps, args.warmup, dtype=npdt, points=args.points, log2n=args.log2n)
        info["value"] = val
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.real, "data": "synthetic",
                "config": config, "cpu_baseline": info, "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "note": "reference CPU arm: SCTL's own OpenMP/Vec<> Eval on the host cores, bounded target sample per step"}
        out.write(json.dumps(line) + "\n")
        out.flush()
        return

    import torch
    import torch.distributed as dist

    import sctl_b200 as sb
    from sctl_b200 import synth
    from sctl_b200.sharded import ShardedEval, block_range

    sb.load_library()
    if not torch.cuda.is_available() or sb.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)

    # ---- inputs: every rank generates only its own slices (counter-based generator) ----------------------
    s0, s1 = block_range(N, rank, world)
    t0, t1 = block_range(N, rank, world)
    r_src, n_src, v_src, r_trg = synth.cube(s1 - s0, t1 - t0, k0, nd, dtype=npdt, first_src=s0, first_trg=t0)
    if args.points == "torus":  # surface-point set; coordinates and normals from the torus, densities as above
        X, Nrm, _, _ = synth.torus(1 << ((args.log2n + 1) // 2), 1 << (args.log2n // 2), dtype=npdt)
        r_src, r_trg = X[s0 * 3:s1 * 3].copy(), X[t0 * 3:t1 * 3].copy()
        n_src = Nrm[s0 * 3:s1 * 3].copy() if nd else None
    tdt = torch.float64 if rb == 8 else torch.float32
    pin = lambda a: None if a is None else torch.from_numpy(a).pin_memory()
    h_r_src, h_n_src, h_v_src, h_r_trg = pin(r_src), pin(n_src), pin(v_src), pin(r_trg)
    h_v_trg = torch.zeros((t1 - t0) * k1, dtype=tdt).pin_memory()
    to_dev = lambda h: None if h is None else h.to(dev, non_blocking=True)
    d_r_src, d_n_src, d_v_src, d_r_trg = to_dev(h_r_src), to_dev(h_n_src), to_dev(h_v_src), to_dev(h_r_trg)
    d_v_trg = torch.zeros((t1 - t0) * k1, dtype=tdt, device=dev)
    sh = ShardedEval(kernel, N, N, dtype=tdt, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    h_out = torch.empty((t1 - t0) * k1, dtype=tdt).pin_memory()
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        d_v_trg.zero_()
        v = sh.eval(d_r_src, d_n_src, d_v_src, d_r_trg, v_trg_local=d_v_trg, digits=args.digits)
        return sh.gather(v) if world > 1 else v

    K = sb.KERNELS[kernel]
    np_view = lambda h: None if h is None else h.numpy()

    def step_e2e_cabi():
        # N = 1: the reference-facing entry itself -- sctl_cuda_eval on HOST buffers (H2D of every input, pack, probe, pair
        # kernels, D2H of the result, all inside the call); the caller's v_trg is zero here, Eval accumulates into it
        h_v_trg.zero_()
        return K.Eval(np_view(h_v_trg), np_view(h_r_trg), np_view(h_r_src), np_view(h_n_src), np_view(h_v_src), digits=args.digits, dtype=npdt)

    def step_e2e():
        if world == 1:
            return step_e2e_cabi()
        a, b, c, d = to_dev(h_r_src), to_dev(h_n_src), to_dev(h_v_src), to_dev(h_r_trg)
        v = to_dev(h_v_trg)  # Eval accumulates: the caller's v_trg goes up with the inputs
        v = sh.eval(a, b, c, d, v_trg_local=v, digits=args.digits)
        h_out.copy_(v, non_blocking=True)
        stream.synchronize()
        return h_out

    def timed(fn, steps):
        per_step = []
        for _ in range(steps):
            flush.fill_(1)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            per_step.append(float(ms.item()))
        return per_step

    # ---- peak (roofline denominator) and warm-up -----------------------------------------------------------
    peak_burst = sb.fma_peak(rb, 50.0)
    for _ in range(args.warmup):
        step_resident()
    barrier()
    plan = sb.last_plan()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res_ms = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    full = step_resident()  # one more, untimed, to keep the result for the parity figure
    torch.cuda.synchronize()

    # kernel-only time: pack + pair kernel (+ reduce) on this rank, no collectives
    ker_ms = []
    for _ in range(max(2, min(args.steps, 3))):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d_v_trg.zero_()
        e0.record(stream)
        sb.eval_device(kernel, sh.r_src, sh.n_src, sh.v_src, d_r_trg, d_v_trg, digits=args.digits)
        e1.record(stream)
        torch.cuda.synchronize()
        ker_ms.append(e0.elapsed_time(e1))
    peak_sustained = sb.fma_peak(rb, 1500.0)  # right after the timed region: same thermal / power state

    e2e = None
    if not args.no_e2e:
        step_e2e()
        e2e_ms = timed(step_e2e, args.steps)
        bi = sum(h.numel() * rb for h in (h_r_src, h_n_src, h_v_src, h_r_trg, h_v_trg) if h is not None)
        e2e = {"value": float(N) * N * args.steps / (sum(e2e_ms) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": bi * world, "d2h_bytes_per_step": h_out.numel() * rb * world,
               "ms_per_step": sum(e2e_ms) / args.steps, "api": ("sctl_cuda_eval (C ABI, host pointers: H2D + pack + probe + pair kernels + D2H inside the call)" if world == 1 else
                       "sctl_b200.sharded.ShardedEval.eval on pinned host buffers (H2D of each rank's slices + NCCL replicate + Eval + D2H per step)")}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    pairs = float(N) * N
    total_s = sum(res_ms) * 1e-3
    value = pairs * args.steps / total_s
    my_pairs = float(N) * (t1 - t0)
    k_ms = float(np.mean(ker_ms))
    achieved = my_pairs * F_ALG[kernel] / (k_ms * 1e-3) * 1e-12
    # DRAM bytes of one pair-kernel launch from the ncu --set full capture of this same workload
    # (profiles/r1_ncu_eval_const_final_stokes_fxu_f64.md): 50.38 MB read + 0.02 MB written per constant-bank launch at 2^20
    # targets (r_trg + the v_trg read-modify-write; the write-back stays in L2); null for any other workload
    traffic = 50.40e6 if (kernel, args.log2n, rb, world) == ("Stokes3D_FxU", 20, 8, 1) else None
    const_path = plan["n_launches"] > 10  # one launch per source chunk (the ring paths need at most 5)
    # launches of one Eval on the constant-bank path: pack + 2 probe kernels + one pair kernel per chunk + the bank combine
    pair_launches = max(1, plan["n_launches"] - 4) if const_path else 1
    # FP-pipe instructions per pair of the loop that ran: FP64 by digits class (cubic / first-order / seed only), FP32 seed
    # only at digits -1 .. 6 and one Newton step (3 instructions) above
    if rb == 8:
        fp_instr = FP_INSTR[kernel] - (0 if (args.digits < 0 or args.digits >= 12) else 1 if args.digits >= 6 else 5)
    else:
        fp_instr = FP_INSTR32[kernel] + (3 if args.digits >= 7 else 0)
    roofline = {"bound": "fp64" if rb == 8 else "fp32", "achieved": achieved, "peak": peak_sustained, "unit": "TFLOP/s", "frac": achieved / peak_sustained, "traffic": traffic,
                "peak_source": "%s micro-benchmark in this run (sctl_cuda_fma_peak, 1.5 s sustained, FMA = 2 flop); burst 50 ms figure %.2f; MEASURED_PEAKS.json has no FP64/FP32 CUDA-core entry" % ("DFMA" if rb == 8 else "FFMA", peak_burst),
                "note": "compute-bound on the %s CUDA-core pipe (pointwise rsqrt/FMA, ~0 bytes per pair): neither 'hbm' nor 'tensor' applies" % ("FP64" if rb == 8 else "FP32"),
                "kernel": "eval_pairs%s_kernel<%s, %s>" % ("_const" if const_path else "", "double" if rb == 8 else "float", kernel),
                "kernel_ms_per_step": k_ms, "pair_kernel_launches_per_step": pair_launches,
                "launch_overlap": "chunk i runs from constant-bank slice i %% %d on its own stream, so consecutive pair-kernel launches overlap; kernel_ms_per_step is the wall time of the whole chain on the device" % (2 if rb == 8 else 4) if const_path else None,
                "pairs_per_step": my_pairs, "flop_per_pair": F_ALG[kernel],
                "frac_sctl_flops_convention": my_pairs * SCTL_FLOPS[kernel] / (k_ms * 1e-3) * 1e-12 / peak_sustained,
                "fp_instr_per_pair": fp_instr, "pipe_ceiling_frac": F_ALG[kernel] / (2.0 * fp_instr)}
    if traffic is not None and const_path:
        # the contract's HBM view of the same kernel, to show that it is not the binding roof: DRAM bytes of one pair-kernel
        # launch (ncu) x launches per step / kernel time, against the measured copy bandwidth
        hbm_peak, hbm_src = hbm_peak_gbs()
        gbs = traffic * pair_launches / (k_ms * 1e-3) * 1e-9
        roofline["hbm_view"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": traffic, "peak_source": hbm_src,
                                "note": "targets and v_trg are re-read once per source chunk (%d launches per step); at %.1f %% of the HBM roof the kernel is not memory-bound" % (pair_launches, 100 * gbs / hbm_peak)}

    if world == 1 and not args.no_kmat_view:
        # The path's one HBM-bound kernel, in the contract's own roofline form: KernelMatrix of the same micro-kernel and
        # precision (generic-kernel.txx:213-329; csrc/kmat_kernel.cuh), 8192 x 8192 device-resident, 8 launches back to back
        # per sample, best of 3, CUDA events on the launching stream.  Algorithmic bytes = sizeof(Real)*K0*K1 written per pair.
        nk = 8192
        km_rs, km_ns, _, km_rt = synth.cube(nk, nk, k0, nd, dtype=npdt)
        d_rs, d_rt = torch.from_numpy(km_rs).cuda(), torch.from_numpy(km_rt).cuda()
        d_ns = torch.from_numpy(km_ns).cuda() if nd else None
        d_M = torch.empty((nk * k0, nk * k1), dtype=d_rs.dtype, device="cuda")
        sb.kernel_matrix_device(kernel, d_rs, d_ns, d_rt, d_M)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(8):
                sb.kernel_matrix_device(kernel, d_rs, d_ns, d_rt, d_M)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 8)
        alg = float(nk) * k0 * nk * k1 * rb
        hbm_peak, hbm_src = hbm_peak_gbs()
        gbs = alg / (best * 1e-3) * 1e-9
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/r2k_ncu_kernel_matrix.md)
        km_traffic = {("Stokes3D_FxU", 8): 4.775e9, ("Stokes3D_FxU", 4): 2.358e9, ("Laplace3D_FxU", 8): 0.479e9}.get((kernel, rb))
        roofline["kernel_matrix_view"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": km_traffic,
                                          "algorithmic_bytes_per_launch": alg, "ms_per_launch": best, "peak_source": hbm_src,
                                          "kernel": "kernel_matrix_kernel<%s, %s>" % ("double" if rb == 8 else "float", kernel),
                                          "note": "KernelMatrix %d x %d of the same micro-kernel: %d bytes written per pair, no reads to speak of; the peak is a read+write copy figure, a write-only stream can exceed it" % (nk, nk, k0 * k1 * rb)}
        del d_M

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(res_ms) / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.real, "data": "synthetic", "config": config, "clocks": clocks,
            "gpu_launches": int(plan["n_launches"]) * args.steps, "launch_plan": plan, "roofline": roofline}
    if e2e:
        line["e2e"] = e2e

    if not args.no_cpu_baseline and world == 1:
        nsub = min(N, args.cpu_targets)
        val, info, (idx, cpu_res), _ = cpu_arm(kernel, N, nsub, 2, 2, want_result=True, dtype=npdt, points=args.points, log2n=args.log2n)
        info.update({"value": val, "unit": UNIT})
        line["cpu_baseline"] = info
        got = full.cpu().numpy().reshape(-1, k1)[idx].reshape(-1)
        line["parity"] = {"max_abs_err_over_vinf": float(np.abs(got.astype(np.float64) - cpu_res.astype(np.float64)).max() / np.abs(full.cpu().numpy()).max()),
                          "gate": 1e-12 if rb == 8 else "1e-5 for Ns <= 2^14; beyond that two FP32 summation orders differ by more (SURVEY 7.3)",
                          "checked": "%d strided targets x all sources vs the CPU arm above" % nsub}
    if world > 1 and not args.no_cpu_baseline:
        # multi-GPU runs carry no cpu_baseline (N = 1 only) but still check the gathered result against the CPU reference
        nsub = min(N, 1 << 12)
        _, info, (idx, cpu_res), _ = cpu_arm(kernel, N, nsub, 1, 1, want_result=True, dtype=npdt, points=args.points, log2n=args.log2n)
        got = full.cpu().numpy().reshape(-1, k1)[idx].reshape(-1)
        line["parity"] = {"max_abs_err_over_vinf": float(np.abs(got.astype(np.float64) - cpu_res.astype(np.float64)).max() / np.abs(full.cpu().numpy()).max()),
                          "gate": 1e-12 if rb == 8 else "1e-5 for Ns <= 2^14; beyond that two FP32 summation orders differ by more (SURVEY 7.3)",
                          "checked": "%d strided targets x all sources of the gathered %d-GPU result vs the reference CPU Eval (%s, %d threads)" % (nsub, world, info["kind"], info["cores"])}
    out.write(json.dumps(line) + "\n")
    out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
