/* sctl_cuda.h -- C ABI of the B200-native (sm_100a) dense kernel-summation module.
 *
 * Drop-in boundary for SCTL's GenericKernel<uKernel>::Eval / KernelMatrix (reference paths are relative to
 * /root/reference/include/sctl/):
 *
 *   sctl_cuda_eval            replaces the body of GenericKernel<uK>::Eval<Real,omp,digits>
 *                             (generic-kernel.txx:102-211) and is what the runtime-digits trampoline
 *                             Eval<Real,omp>(..., Integer digits, ConstIterator<char> self)
 *                             (generic-kernel.hpp:99, generic-kernel.txx:72-100) ends up calling.
 *   sctl_cuda_kernel_matrix   replaces GenericKernel<uK>::KernelMatrix<Real,omp,digits>
 *                             (generic-kernel.txx:213-329).
 *   sctl_cuda_kernel_info     mirrors CoordDim/NormalDim/SrcDim/TrgDim (generic-kernel.txx:48-62),
 *                             uKernel::FLOPS() and uKerScaleFactor<Real>() (kernel_functions.hpp:37-42, ...).
 *
 * All arrays are the reference's AoS Vector<Real> layouts (generic-kernel.txx:141,145-147,154):
 *   r_src[Ns*3], n_src[Ns*ND] (may be NULL when ND == 0), v_src[Ns*K0], r_trg[Nt*3], v_trg[Nt*K1],
 *   M[(Ns*K0) x (Nt*K1)] row-major with leading dimension ldM;  Real = double (real_bytes 8) or float (4).
 * (K0, K1, ND) per kernel id: FxU (1,1,0)  DxU (1,1,3)  FxdU (1,3,0)  Stokes FxU (3,3,0)  Stokes DxU (3,3,3);
 * ids 5-11 are listed with the enum below.
 *
 * Semantics kept from the reference:
 *   - Eval ACCUMULATES: v_trg += scale * sum_s K(x_t - x_s, n_s) f_s  (generic-kernel.txx:204-208).  The
 *     "wrong size => ReInit + SetZero" rule (:129-132) is applied by the C++ wrapper before the call.
 *   - KernelMatrix OVERWRITES M (:268,306).
 *   - pairs with r^2 == 0 contribute exactly 0 (approx_rsqrt(r2, r2 > 0), kernel_functions.hpp:45).
 *   - digits in [-1, 15] selects the accuracy of the reciprocal square root exactly like
 *     approx_rsqrt<digits> (vec.txx:375-382): the result is accurate to at least `digits` decimal digits;
 *     -1 = working precision of Real.  Values >= 16 are treated as -1 (generic-kernel.txx:97).
 *   - inputs are never modified; no pointer is retained after return.
 * Errors: the reference aborts through SCTL_ASSERT (common.hpp:98-105).  Here every entry returns 0 on success
 * and nonzero on failure, with a thread-local message in sctl_cuda_last_error(); the C++ wrapper turns that
 * into SCTL_ERROR.  There is NO CPU fallback: without a usable CUDA device every compute entry fails.
 * Threading: every entry may be called concurrently from many host threads (the reference calls
 * KernelMatrix<Real,false> inside an OpenMP parallel loop, boundary_integral.txx:1169-1211); each host thread
 * gets its own stream.
 */
#ifndef SCTL_CUDA_H_
#define SCTL_CUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum sctl_cuda_kernel_id {
  SCTL_CUDA_LAPLACE3D_FXU = 0,  /* sctl::Laplace3D_FxU   kernel_functions.hpp:32-59   */
  SCTL_CUDA_LAPLACE3D_DXU = 1,  /* sctl::Laplace3D_DxU   kernel_functions.hpp:61-79   */
  SCTL_CUDA_LAPLACE3D_FXDU = 2, /* sctl::Laplace3D_FxdU  kernel_functions.hpp:81-123  */
  SCTL_CUDA_STOKES3D_FXU = 3,   /* sctl::Stokes3D_FxU    kernel_functions.hpp:125-178 */
  SCTL_CUDA_STOKES3D_DXU = 4,   /* sctl::Stokes3D_DxU    kernel_functions.hpp:180-218 */
  /* beyond the north-star five: the other non-Helmholtz kernels */
  SCTL_CUDA_STOKES3D_FXT = 5,       /* sctl::Stokes3D_FxT      kernel_functions.hpp:220-254  (3,9,0) */
  SCTL_CUDA_STOKES3D_FSXU = 6,      /* sctl::Stokes3D_FSxU     kernel_functions.hpp:256-324  (4,3,0) */
  SCTL_CUDA_STOKES3D_FXUP = 7,      /* sctl::Stokes3D_FxUP     kernel_functions.hpp:326-364  (3,4,0) */
  SCTL_CUDA_LAPLACE3D_FXD2U = 8,    /* sctl::Laplace3D_Fxd2U   kernel_functions.hpp:366-389  (1,9,0) */
  SCTL_CUDA_LAPLACE3D_DXDU = 9,     /* sctl::Laplace3D_DxdU    kernel_functions.hpp:391-412  (1,3,3) */
  SCTL_CUDA_BIOTSAVART3D_FXU = 10,  /* sctl::BiotSavart3D_FxU  kernel_functions.hpp:414-448  (3,3,0) */
  SCTL_CUDA_BIOTSAVART3D_FXDU = 11, /* sctl::BiotSavart3D_FxdU kernel_functions.hpp:450-483  (3,9,0) */
  /* Helmholtz: complex values as adjacent (re, im) pairs; ctx = HOST pointer to the wavenumber, one Real */
  SCTL_CUDA_HELMHOLTZ3D_FXU = 12,      /* sctl::Helmholtz3D_FxU      kernel_functions.hpp:490-533  (2,2,0) */
  SCTL_CUDA_HELMHOLTZ3D_DXU = 13,      /* sctl::Helmholtz3D_DxU      kernel_functions.hpp:535-578  (2,2,3) */
  SCTL_CUDA_HELMHOLTZ3D_FXDU = 14,     /* sctl::Helmholtz3D_FxdU     kernel_functions.hpp:580-627  (2,6,0) */
  SCTL_CUDA_HELMHOLTZDIFF3D_FXDU = 15, /* sctl::HelmholtzDiff3D_FxdU kernel_functions.hpp:629-659  (2,6,0) */
  SCTL_CUDA_NUM_KERNELS = 16
};

/* ---- lifecycle ------------------------------------------------------------------------------ */

/* Number of CUDA devices visible to this process (0 when there is no driver / device). */
int sctl_cuda_device_count(void);

/* Select how many GPUs the host-pointer entries shard over (0 = all visible).  Optional: the first compute
 * call initialises one GPU (the calling thread's current device).  n_gpus > 1 creates one NCCL communicator
 * over devices 0..n_gpus-1 of this process (sources are replicated with ncclBroadcast over NVLink, targets
 * are sharded by contiguous ranges -- the block rule of ParticleFMM::EvalDirect, fmm-wrapper.txx:565-569). */
int sctl_cuda_init(int n_gpus);
int sctl_cuda_finalize(void);
/* GPUs the host-pointer entries currently shard over (>= 1 once initialised, 0 before). */
int sctl_cuda_num_gpus(void);

/* Thread-local description of the last failure on the calling thread ("" if none). */
const char* sctl_cuda_last_error(void);

/* ---- kernel metadata ------------------------------------------------------------------------ */

/* Any out pointer may be NULL.  name: "Laplace3D-FxU" ... (uKernel::Name()).  scale: uKerScaleFactor<double>().
 * f_alg: algorithmic flop per pair used for the roofline. */
int sctl_cuda_kernel_info(int kernel_id, int* k0, int* k1, int* nd, int* flops, int* f_alg, double* scale, const char** name);

/* ---- the hot path: host pointers (the drop-in entries) -------------------------------------- */

int sctl_cuda_eval(int kernel_id, int real_bytes, int digits, int64_t Ns, int64_t Nt,
                   const void* r_src, const void* n_src, const void* v_src, const void* r_trg, void* v_trg,
                   const void* ctx, int64_t ctx_bytes /* uKernel context (GetCtxPtr, generic-kernel.hpp:79): a HOST pointer; the
                                                         Helmholtz kernels read one Real (the wavenumber) from it, the others
                                                         ignore it; ctx_bytes = real_bytes or 0 */);

int sctl_cuda_kernel_matrix(int kernel_id, int real_bytes, int digits, int64_t Ns, int64_t Nt,
                            const void* r_src, const void* n_src, const void* r_trg, void* M, int64_t ldM,
                            const void* ctx, int64_t ctx_bytes);

/* ---- the hot path: device pointers on the current device, enqueued on `stream` -------------- */
/* (cudaStream_t passed as void*; NULL = the legacy default stream.)  Asynchronous: returns after enqueueing.
 * Pointers need no particular alignment (beyond that of Real).  Scratch comes from the stream-ordered pool (cudaMallocAsync).
 * KernelMatrix writes M with 16-byte vector stores whatever d_M and ldM are (rows that do not start on a 16-byte boundary
 * get a few scalar elements where warps meet); it is fastest when d_M is 16-byte aligned and ldM is a multiple of
 * 16 / real_bytes: an odd row pitch makes every warp store straddle one more 32-byte sector (70 - 90 % of the aligned rate). */

int sctl_cuda_eval_device(int kernel_id, int real_bytes, int digits, int64_t Ns, int64_t Nt,
                          const void* d_r_src, const void* d_n_src, const void* d_v_src, const void* d_r_trg, void* d_v_trg,
                          const void* ctx /* HOST pointer */, int64_t ctx_bytes, void* stream);

int sctl_cuda_kernel_matrix_device(int kernel_id, int real_bytes, int digits, int64_t Ns, int64_t Nt,
                                   const void* d_r_src, const void* d_n_src, const void* d_r_trg, void* d_M, int64_t ldM,
                                   const void* ctx /* HOST pointer */, int64_t ctx_bytes, void* stream);

/* ---- measurement and test hooks ------------------------------------------------------------- */

/* Device-side twin of approx_rsqrt<digits>(x, x > 0) on n device values (contract test, vec-test.hpp:431-465). */
int sctl_cuda_rsqrt_device(int real_bytes, int digits, int64_t n, const void* d_x, void* d_y, void* stream);

/* Every host thread that calls a host-pointer entry owns a stream, four events and (after its first small call) a 1 MB
 * pinned + 1 MB device arena per device.  A thread that ends parks them in a process-wide pool; the next new thread
 * takes them over; sctl_cuda_finalize() releases what is parked.  Returns how many sets are parked for `device`. */
int sctl_cuda_idle_thread_resources(int device);

/* Timings (ms, CUDA events on the call's stream) of the calling thread's last host-pointer call. */
int sctl_cuda_last_timing(double* h2d_ms, double* kernel_ms, double* d2h_ms);

/* Launch geometry the last device-level Eval of the calling thread used (any pointer may be NULL). */
int sctl_cuda_last_plan(int* grid_x, int* grid_y, int* threads, int* smem_bytes, int* ctas_per_sm, int* n_launches, double* wave_eff);

/* ---- Device-resident direct evaluator -------------------------------------------------
 * The ParticleFMM::EvalDirect analogue (fmm-wrapper.txx:551-623; driven once per GMRES iteration by
 * BoundaryIntegralOp::ComputeFarField, boundary_integral.txx:1286-1305): source and target coordinates are uploaded
 * once and stay in HBM, an apply ships only the density in and the potential out.  Sources and targets are
 * addressed by small non-negative slot numbers chosen by the caller (the C++ mirror maps ParticleFMM's names to
 * them).  With n_gpus > 1 the targets are sharded by contiguous block over the session's GPUs (the rule of
 * fmm-wrapper.txx:568) and every GPU holds all sources; no reduction between GPUs (results agree to rounding
 * for any n_gpus: a shard may split the sources differently from the whole target set).
 *   set_src_coord    ParticleFMM::SetSrcCoord(name, X, Xn)   (fmm-wrapper.hpp:166)  normal_dim 0 or 3
 *   set_src_density  ParticleFMM::SetSrcDensity(name, F)     (:174)  n_reals = Ns * K0 of the kernels that use it
 *   set_trg_coord    ParticleFMM::SetTrgCoord(name, X)       (:182)
 *   set_kernel_s2t   ParticleFMM::SetKernelS2T(src, trg, k)  (:143)  kernel_id < 0 removes the interaction
 *   eval             ParticleFMM::EvalDirect(U, trg)         (:198)  U[Nt*K1] is OVERWRITTEN with the sum over every
 *                    source slot that has a kernel into `trg` (fmm-wrapper.txx:562-563,598-619); digits as in Eval.
 * All pointers are host pointers; nothing is retained after return.  A session serialises its own calls. */
typedef struct sctl_cuda_direct_s* sctl_cuda_direct_t;
/* n_gpus 0 or 1: the session lives on the device that is CURRENT on the creating thread (with one rank per GPU --
 * torchrun, MPI -- every rank's far field stays on its own GPU, like sctl_cuda_eval and the near-field session);
 * n_gpus > 1: devices 0..n_gpus-1.  create_on takes the device ordinals explicitly (distinct, visible). */
int sctl_cuda_direct_create(int real_bytes, int n_gpus, sctl_cuda_direct_t* out);
int sctl_cuda_direct_create_on(int real_bytes, int n_gpus, const int* devices, sctl_cuda_direct_t* out);
/* The session's GPUs: count and (up to `cap`) CUDA device ordinals; either pointer may be NULL. */
int sctl_cuda_direct_devices(sctl_cuda_direct_t h, int* n_gpus, int* devices, int cap);
int sctl_cuda_direct_destroy(sctl_cuda_direct_t h);
int sctl_cuda_direct_set_src_coord(sctl_cuda_direct_t h, int src, int64_t Ns, const void* X, const void* Xn, int normal_dim);
int sctl_cuda_direct_set_src_density(sctl_cuda_direct_t h, int src, int64_t n_reals, const void* F);
int sctl_cuda_direct_set_trg_coord(sctl_cuda_direct_t h, int trg, int64_t Nt, const void* X);
int sctl_cuda_direct_set_kernel_s2t(sctl_cuda_direct_t h, int src, int trg, int kernel_id, const void* ctx, int64_t ctx_bytes);
int sctl_cuda_direct_eval(sctl_cuda_direct_t h, int trg, int digits, void* U, int64_t n_reals);

/* ---- Batched near-field blocks of BoundaryIntegralOp -------------------------------------
 * One session = one block list: the near_blk_* work list of BoundaryIntegralOp (boundary_integral.txx:1009-1032).
 * Block b couples ns[b] source nodes with nt[b] near targets through a row-major, target-major operator block
 * K_b of (nt[b]*K1) x (ns[b]*K0) Reals at K + k_dsp[b] (:1183); its density is F + f_dsp[b]*K0, its result
 * U + u_dsp[b]*K1, its far-field nodes start at far_dsp[b] in Xs/Xn/wts, its targets at u_dsp[b] in Xt.  In the
 * reference's variables: f_dsp = elem_nds_dsp[elem], far_dsp = elem_nds_dsp_far[elem], ns = elem_nds_cnt[elem],
 * u_dsp = near_elem_dsp[elem] + t0, nt = near_blk_cnt[blk], k_dsp = K_near_dsp[elem]*K0*K1 + t0*K1*ns*K0.
 *   set_matrix       upload K_near as the CPU singular/near quadrature left it; it then stays in HBM
 *   subtract_direct  "KNearSubtract" (:1165-1240) for element lists without a density operator:
 *                    K_b[(t*K1+k1)][(s*K0+k0)] -= wts[s] * KernelMatrix[(s*K0+k0)][(t*K1+k1)], every block in one launch.
 *                    With Xn_trg (trg_normal_dot_prod, :1186-1200) the kernel's 3*K1 target components are contracted
 *                    with the target normals first, sum_l KernelMatrix[..][(t*K1+k1)*3+l] * Xn_trg[t*3+l]; the session
 *                    is then created with K1 = (kernel target dim) / 3
 *   apply            "EvalNear" (:1324-1340): U_b = K_b * F_b for every block in one launch (overwrites U; blocks
 *                    without sources give zeros, :1327-1331); F and U are host arrays, only they cross PCIe
 *   set_targets      the tables of ComputeNearInterac's tail (:1366-1381): near_trg_cnt / near_trg_dsp per target and
 *                    near_scatter_index (one process: ScatterForward is out[p] = in[scatter[p]], comm.txx:1499-1510)
 *   apply_accumulate "EvalNear" + that tail in one call: U[i*K1+k] += sum_{j < near_trg_cnt[i]}
 *                    U_near[near_scatter_index[near_trg_dsp[i]+j]*K1+k], j ascending onto what U holds (the reference's
 *                    order); U has n_trg*K1 values, so per apply F and U go up and U comes down -- U_near (one entry per
 *                    near PAIR, typically 10x the targets) never leaves the device
 *   get_matrix       download K (tests)
 * All tables are copied at create; all pointers are host pointers; nothing is retained after return. */
typedef struct sctl_cuda_near_s* sctl_cuda_near_t;
int sctl_cuda_near_create(int real_bytes, int k0, int k1, int64_t n_blk, const int64_t* f_dsp, const int64_t* far_dsp, const int64_t* ns,
                          const int64_t* u_dsp, const int64_t* nt, const int64_t* k_dsp, sctl_cuda_near_t* out);
int sctl_cuda_near_destroy(sctl_cuda_near_t h);
int sctl_cuda_near_sizes(sctl_cuda_near_t h, int64_t* K_len, int64_t* nF, int64_t* nU, int64_t* n_far);
int sctl_cuda_near_set_matrix(sctl_cuda_near_t h, const void* K, int64_t n_reals);
int sctl_cuda_near_get_matrix(sctl_cuda_near_t h, void* K, int64_t n_reals);
int sctl_cuda_near_subtract_direct(sctl_cuda_near_t h, int kernel_id, int digits, int64_t Ns, const void* Xs, const void* Xn, const void* wts,
                                   int64_t Nt, const void* Xt, const void* Xn_trg /* NULL, or trg_normal_dot_prod */, const void* ctx, int64_t ctx_bytes);
int sctl_cuda_near_apply(sctl_cuda_near_t h, const void* F, int64_t nF, void* U, int64_t nU);
int sctl_cuda_near_set_targets(sctl_cuda_near_t h, int64_t n_trg, const int64_t* near_trg_cnt, const int64_t* near_trg_dsp, int64_t n_near,
                               const int64_t* near_scatter_index);
int sctl_cuda_near_apply_accumulate(sctl_cuda_near_t h, const void* F, int64_t nF, void* U /* n_trg*K1, accumulated into */, int64_t nU);

/* ---- One BoundaryIntegralOp apply as one device-side pipeline (SURVEY section 8 f1 + f3) ------------------------
 * BoundaryIntegralOp::ComputePotential = ComputeFarField + ComputeNearInterac (boundary_integral.txx:716-722,
 * 1248-1309, 1311-1383), once per GMRES iteration with fixed geometry.  The session borrows a direct session (one
 * interaction src -> trg, coordinates set) and, optionally, a near-field session on the direct session's first GPU
 * (matrix and target tables set); both must outlive it and must not be used while an apply is running.
 *   create   wts_far[Nsrc]: quadrature weights of the far-field nodes (:1270-1283); trg_normal_dot_prod with Xn_trg[Ntrg*3]
 *            contracts the kernel's 3*KDIM1_ target components with the target normals (:1293-1303)
 *   apply    U[Ntrg*KDIM1_] = far field of (F_far * wts_far) + near correction of F.  F: the node density (Nnodes*K0);
 *            F_far_nodes: the density at the far-field nodes where the element list has a far-field density operator
 *            (GetFarFieldDensity, :1267), or NULL / 0 when the far-field nodes are the nodes (every element list of the
 *            reference tree, :527-529) and F itself is used.  Per apply the density goes up once per GPU and U comes down
 *            once; F * wts, the normal contraction and the per-target accumulation are device kernels; the near-field GEMV
 *            (HBM-bound) runs on its own stream beside the far-field launches (FP64-bound); the shards of a multi-GPU far
 *            field are collected on the first GPU with peer copies.  Same arithmetic and summation order as
 *            sctl_cuda_direct_eval followed by sctl_cuda_near_apply_accumulate (bit-identical without trg_normal_dot_prod).
 *   last_timing  device times of the last apply (ms): until the far field has landed on the first GPU; the tail behind it
 *            (contraction, accumulation); the whole apply including the download. */
typedef struct sctl_cuda_bio_s* sctl_cuda_bio_t;
int sctl_cuda_bio_create(sctl_cuda_direct_t far, int src, int trg, sctl_cuda_near_t near /* may be NULL */, int trg_normal_dot_prod,
                         const void* wts_far, int64_t n_wts, const void* Xn_trg, int64_t n_Xn_trg, sctl_cuda_bio_t* out);
int sctl_cuda_bio_destroy(sctl_cuda_bio_t h);
int sctl_cuda_bio_apply(sctl_cuda_bio_t h, int digits, const void* F, int64_t nF, const void* F_far_nodes, int64_t nF_far, void* U, int64_t nU);
int sctl_cuda_bio_last_timing(sctl_cuda_bio_t h, double* far_ms, double* tail_ms, double* total_ms);
/* Work items of the near-field GEMV that each GPU of the far-field session runs (items[p] for the p-th GPU, up to cap
 * entries); all zero when the GEMV is not sharded (one GPU, no near session, or a block list whose U_near ranges do not
 * follow the work-list order), in which case the near session's own GPU runs all of it. */
int sctl_cuda_bio_near_shards(sctl_cuda_bio_t h, int64_t* items, int cap);

/* Tuning knobs (process-wide).  Eval picks its kernel by target count:
 *   "const_min_targets" (default 2^15): at or above, sources are streamed through the constant bank in chunks (FP64,
 *                       and FP32 unless "const_fp32" is 0, which keeps FP32 on the shared-memory ring);
 *   "const_min_launch_flop" (default 5e8): ... and, in FP64, only if one launch of that path carries at least this much
 *                       work (records per launch x targets x F_alg): cheap kernels switch at 2^16 targets;
 *   "wide_min_targets"  (default 2^13): at or above (FP64), 4 targets per thread, one CTA per SM;
 *   "few_max_targets"   (default 2048): at or below, the warps of a CTA share 32 targets and split the sources.
 *   "ring_min_tiles" / "ring_max_split" (defaults 1 / 64): shared-memory ring paths: shortest source chunk (in tiles) and
 *                       most chunks the grid planner may use when it splits the sources to fill the SMs;
 *   "const_banks" / "const_banks_fp32" (defaults 2 / 4): slices of the constant buffer in flight on the constant-bank
 *                       path (FP64 / FP32); chunk i streams through slice i % banks on its own stream so consecutive
 *                       launches overlap (1..4);
 *   "const_ctas_per_sm" / "const_ctas_per_sm_overlap" (defaults 6 / 1): CTAs per SM a launch should have before its
 *                       chunk is split across a thread-block cluster (one launch at a time / overlapping launches);
 *   "small_call_arena"  (default 1): host-pointer calls whose arrays fit 1 MB use a per-thread pinned/device arena
 *                       (one upload, one download, no pool allocations);
 *   "probe_per_block"   (default 1): on the constant-bank path the coincidence check also records which (source chunk,
 *                       target block) pairs hold coincident points, so only those CTAs keep the mask;
 *   "probe_min_pairs"   (default 2e9): at or above Ns*Nt, Eval first checks (O(Ns+Nt), on the device) whether any source
 *                       and target (nearly) coincide and lets the pair kernels drop the r^2 == 0 mask when none do;
 *                       <= 0 disables the check.  Results are bit-identical either way.
 * Results agree to rounding whichever path runs; the knobs exist for tuning and for tests. */
int sctl_cuda_set_option(const char* name, int64_t value);
/* Back to the built-in defaults. */
int sctl_cuda_reset_options(void);

/* Layout of the constant-bank path for one call, exposed for host-side tests (no GPU needed): slices in flight, records
 * per launch, 32-bit words of coincidence bits per chunk (0: none), bytes between slices, bytes per record, bytes per
 * chunk segment in HBM, number of launches. */
int sctl_cuda_const_geometry(int kernel_id, int real_bytes, int64_t Ns, int64_t Nt, int banks_wanted, int* banks, int* chunk, int* row_words,
                             int* bank_stride, int* rec_bytes, int64_t* seg_bytes, int64_t* n_chunks);

/* Test hook: the slices the multi-GPU apply pipeline cuts the near-field GEMV into for a block list (columns as in
 * sctl_cuda_near_create) and n_gpus GPUs.  slices: 6 values per GPU -- items [i0, i1) of the GEMV work list, Reals of K read
 * [klo, khi), Reals of U_near written [ulo, uhi); n_items: length of the work list.  Returns 0, or -1 when the block list
 * cannot be sharded (its U_near ranges do not follow the work-list order). */
int sctl_cuda_plan_near_slices(int real_bytes, int k0, int k1, int64_t n_blk, const int64_t* ns, const int64_t* u_dsp, const int64_t* nt, const int64_t* k_dsp, int n_gpus,
                               int64_t* slices, int64_t* n_items);

/* The grid planner, exposed for host-side tests (no GPU needed): for Nt targets in blocks of `tpb`, Ns sources
 * and `slots` resident CTAs (SMs x CTAs/SM) returns the number of source chunks, their length and the fraction
 * of the last wave that does useful work. */
int sctl_cuda_plan_eval(int64_t Ns, int64_t Nt, int tpb, int slots, int min_chunk, int max_split,
                        int64_t* n_tblocks, int* n_split, int64_t* chunk, double* wave_eff);

/* FMA-throughput micro-benchmark of the current device: dependent-free DFMA (real_bytes 8) or FFMA (4)
 * chains for about `ms` milliseconds; returns achieved TFLOP/s (FMA = 2 flop) -- the roofline denominator. */
int sctl_cuda_fma_peak(int real_bytes, double ms, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* SCTL_CUDA_H_ */
