// libsctl_cuda.so -- C ABI (include/sctl_cuda.h) over the sm_100a kernels in eval_kernel.cuh / kmat_kernel.cuh.
// Host side of the drop-in for GenericKernel<uKernel>::Eval / KernelMatrix
// (reference: include/sctl/generic-kernel.txx:102-329).  No CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is dlopen'ed in sctl_cuda_init(n > 1)

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/sctl_cuda.h"
#include "coincidence.cuh"
#include "eval_const_kernel.cuh"
#include "eval_kernel.cuh"
#include "kmat_kernel.cuh"
#include "near_kernel.cuh"

using namespace sctl_cuda;

namespace {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
thread_local std::string tl_err;

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  tl_err = buf;
  return 1;
}

#define CU_TRY(expr)                                                                          \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// per-thread state
// ---------------------------------------------------------------------------------------------
constexpr int MAX_DEV = 16;

// Per (host thread, device) resources.  A thread that ends hands them back to a process-wide pool instead of leaking
// them (applications that create and destroy many threads would otherwise leak 1 MB of pinned memory per thread); the
// next new thread takes them over, and sctl_cuda_finalize() releases whatever sits in the pool.
struct DevRes {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {};
  // small-call arena: one pinned host buffer and one device buffer, see eval_host_small()
  char* arena_h = nullptr;
  char* arena_d = nullptr;
  bool empty() const { return !stream && !arena_h; }
};
struct ResPool {
  std::mutex mu;
  std::vector<DevRes> idle[MAX_DEV];
};
// never destroyed: thread_local destructors (which return resources) may run after static destructors would have
ResPool& res_pool() {
  static ResPool* p = new ResPool();
  return *p;
}

struct ThreadState {
  DevRes res[MAX_DEV];
  double h2d_ms = 0, kernel_ms = 0, d2h_ms = 0;
  int grid_x = 0, grid_y = 0, threads = 0, smem = 0, ctas_per_sm = 0, n_launches = 0;
  double wave_eff = 0;
  ~ThreadState() {
    // no CUDA calls here (the primary context may already be gone at process exit): just park the handles
    ResPool& pool = res_pool();
    std::lock_guard<std::mutex> lk(pool.mu);
    for (int d = 0; d < MAX_DEV; d++)
      if (!res[d].empty()) pool.idle[d].push_back(res[d]);
  }
};
thread_local ThreadState tl;

struct DeviceInfo {
  int sms = 0;
  bool pool_configured = false;
  // The constant bank that the chunk-streaming kernel reads is one buffer per device: calls that use it are
  // enqueued under this mutex and chained on the GPU through `const_done`, whatever stream they run on.
  std::mutex const_mu;
  cudaEvent_t const_done = nullptr;
  cudaEvent_t const_fork = nullptr;
  cudaStream_t bank_stream[CONST_MAX_BANKS] = {};  // [0] unused: bank 0 runs on the caller's stream
  cudaEvent_t bank_done[CONST_MAX_BANKS] = {};
};

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct Options {
  // Eval path selection by target count (sctl_cuda_set_option):
  //   Nt >= const_min_targets : constant-bank chunk streaming (eval_const_kernel.cuh)
  //   Nt <= few_max_targets   : shared-memory ring, all warps of a CTA share 32 targets and split the sources
  //   Nt >= wide_min_targets  : shared-memory ring, 4 targets per thread, one CTA per SM
  //   otherwise               : shared-memory ring, 2 targets per thread, two CTAs per SM
  int64_t const_min_targets = 1 << 15;
  // ... and (FP64) only when one launch carries at least this much work: records per launch x targets x F_alg
  int64_t const_min_launch_flop = 500000000LL;
  int64_t few_max_targets = 2048;
  // shared-memory ring: shortest source chunk (in tiles) the grid planner may give a CTA when it splits the sources
  int64_t ring_min_tiles = 1;
  int64_t ring_max_split = 64;
  int64_t wide_min_targets = 1 << 13;
  // Constant-bank path: slices of the buffer in flight (1 = one launch at a time), and the CTAs per SM a launch should
  // have before its chunk is split across a cluster (stand-alone launches / overlapping launches).
  int64_t const_banks = 2;
  int64_t const_ctas_per_sm = 6;
  int64_t const_ctas_per_sm_overlap = 1;
  // FP32 on the constant-bank path: 0 never, otherwise from const_min_targets like FP64.  With the packed FP32 loops,
  // unrolled over 4 sources, it is the faster path from 2^15 targets for every kernel but Laplace3D_FxU, which is within
  // 1 % there and 10 % ahead from 2^17 (profiles/r2c_ab_rate32_packed_vs_scalar.md, r2e_ab_rate32_const_shapes.md); round 1's
  // scalar loops only won from 2^17.  const_banks_fp32 slices in flight (FP32 launches are half as long, so four slices hide
  // their boundaries better).
  int64_t const_fp32 = -1;
  int64_t const_banks_fp32 = 4;
  // Above this many pairs Eval first runs the O(Ns+Nt) coincidence probe (coincidence.cuh) and lets the pair kernels
  // drop the r2 == 0 mask when no source/target pair can need it.  <= 0 disables the probe (always mask).
  int64_t probe_min_pairs = 2000000000LL;
  // Constant-bank path: when coincident pairs exist, let the probe also mark WHICH (chunk, target block) pairs hold
  // them, so every other CTA still takes the unmasked loop (targets == sources: one coincident pair per target).
  int64_t probe_per_block = 1;
  // host-pointer calls whose arrays fit 1 MB go through a per-thread pinned/device arena (one upload, one download)
  int64_t small_call_arena = 1;
};

struct Global {
  Options opt;
  std::mutex mu;
  int n_gpus = 0;  // 0 = not initialised
  DeviceInfo dev[MAX_DEV];
  NcclApi nccl;
  ncclComm_t comms[MAX_DEV] = {};
  bool have_comms = false;
  // sharded (multi-GPU) host calls run one at a time on persistent per-device streams
  std::mutex multi_mu;
  cudaStream_t mstream[MAX_DEV] = {};
  cudaEvent_t mev[MAX_DEV][4] = {};
};
Global g;

int ensure_device_info(int dev) {
  std::lock_guard<std::mutex> lk(g.mu);
  {
    if (!g.dev[dev].pool_configured) {
      cudaDeviceProp prop;
      CU_TRY(cudaGetDeviceProperties(&prop, dev));
      if (prop.major < 10) return fail("device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, prop.major, prop.minor);
      g.dev[dev].sms = prop.multiProcessorCount;
      cudaMemPool_t pool;
      CU_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
      uint64_t thresh = UINT64_MAX;  // keep scratch cached between calls
      CU_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
      g.dev[dev].pool_configured = true;
    }
  }
  return 0;
}

int ensure_device(int dev) {
  if (ensure_device_info(dev)) return 1;
  DevRes& r = tl.res[dev];
  if (!r.stream) {
    {  // a finished thread's resources, if any
      ResPool& pool = res_pool();
      std::lock_guard<std::mutex> lk(pool.mu);
      if (!pool.idle[dev].empty()) {
        r = pool.idle[dev].back();
        pool.idle[dev].pop_back();
      }
    }
    if (r.stream) return 0;
    int cur;
    CU_TRY(cudaGetDevice(&cur));
    CU_TRY(cudaSetDevice(dev));
    CU_TRY(cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; i++) CU_TRY(cudaEventCreate(&r.ev[i]));
    CU_TRY(cudaSetDevice(cur));
  }
  return 0;
}

int visible_devices() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// ---------------------------------------------------------------------------------------------
// digits -> refinement of the reciprocal-square-root seed  (approx_rsqrt<digits>, vec.txx:375-382)
// ---------------------------------------------------------------------------------------------
// approx_rsqrt<digits> promises `digits` decimal digits and the reference picks the cheapest Newton count that delivers
// them for the ISA it runs on (intrin-wrapper.hpp:3384 AVX-512, :2395 AVX, :1566 SSE): the same rule
// here, with the B200 seed and steps (errors measured on the device, tools/seed_error.py, profiles/r2_seed_error.md):
//   FP64: digits <= 5   seed only (MUFU.RSQ64H, max error 9.3e-7);
//         6 .. 11       one first-order step, 4 FP64 instructions (max error 8.6e-13 = 12.07 digits, zero mean) --
//                       this is the class of ParticleFMM's default 10 and BoundaryIntegralOp's 11 (fmm-wrapper.txx:205,
//                       boundary_integral.txx:562); the reference's own AVX/SSE path delivers 1.9e-13 there (11-bit
//                       seed, 2 steps), its AVX-512 path full precision;
//         >= 12, -1     the cubic step, 5 instructions (< 1 ulp).
//   FP32: digits <= 6 seed only (MUFU.RSQ, 2^-22); 7 and up one Newton step.
template <class Real> int rsq_mode_for_digits(int digits) {
  if (sizeof(Real) == 8) {
    if (digits < 0 || digits >= 16) digits = 15;  // (Integer)(52*0.30103)
    return digits <= 5 ? RSQ_SEED : digits <= 11 ? RSQ_NEWTON : RSQ_CUBIC;
  } else {
    if (digits < 0 || digits >= 16) digits = 6;  // (Integer)(23*0.30103)
    return digits <= 6 ? RSQ_SEED : RSQ_NEWTON;
  }
}

// ---------------------------------------------------------------------------------------------
// launch configurations
// ---------------------------------------------------------------------------------------------
// MAIN: many targets.  FEW: few targets -- all warps of a CTA share 32*T targets and split the sources.
// CONST: very many targets -- sources streamed through the constant bank into uniform registers (eval_const_kernel.cuh).
// the constant-bank shapes (FP64, FP32) can be overridden at build time for tuning runs (tools/ab_build.sh)
#ifndef SCTL_TUNE_CONST_NT
#define SCTL_TUNE_CONST_NT 256
#endif
#ifndef SCTL_TUNE_CONST_U
#define SCTL_TUNE_CONST_U 4
#endif
#ifndef SCTL_TUNE_CONST32_T
#define SCTL_TUNE_CONST32_T 4
#endif
#ifndef SCTL_TUNE_CONST32_NT
#define SCTL_TUNE_CONST32_NT 128
#endif
#ifndef SCTL_TUNE_CONST32_U
#define SCTL_TUNE_CONST32_U 4
#endif
#ifndef SCTL_TUNE_CONST32_MINB
#define SCTL_TUNE_CONST32_MINB 4
#endif
template <class Real> struct Tune;
template <> struct Tune<double> {
  static constexpr int T = 2, NW = 8, TILE = 256, STAGES = 3, MINB = 2;
  static constexpr int WIDE_T = 4, WIDE_NW = 8, WIDE_TILE = 256, WIDE_STAGES = 3, WIDE_MINB = 1;
  static constexpr int FEW_T = 1, FEW_NW = 8, FEW_TILE = 256, FEW_STAGES = 3, FEW_MINB = 2;
  static constexpr int CONST_T = 2, CONST_NT = SCTL_TUNE_CONST_NT, CONST_U = SCTL_TUNE_CONST_U, CONST_MINB = 2;
};
template <> struct Tune<float> {
  static constexpr int T = 4, NW = 8, TILE = 512, STAGES = 3, MINB = 2;
  static constexpr int WIDE_T = 4, WIDE_NW = 8, WIDE_TILE = 512, WIDE_STAGES = 3, WIDE_MINB = 2;
  static constexpr int FEW_T = 1, FEW_NW = 8, FEW_TILE = 512, FEW_STAGES = 3, FEW_MINB = 2;
  static constexpr int CONST_T = SCTL_TUNE_CONST32_T, CONST_NT = SCTL_TUNE_CONST32_NT, CONST_U = SCTL_TUNE_CONST32_U, CONST_MINB = SCTL_TUNE_CONST32_MINB;
};

struct DevBuf {
  void* p = nullptr;
  cudaStream_t st = nullptr;
  int alloc(size_t bytes, cudaStream_t s) {
    st = s;
    if (bytes == 0) return 0;
    CU_TRY(cudaMallocAsync(&p, bytes, s));
    return 0;
  }
  ~DevBuf() {
    if (p) cudaFreeAsync(p, st);
  }
};

struct OccCache {
  std::mutex mu;
  int ctas[MAX_DEV] = {};
};

template <class Real, int KID, int T, int NW, int R, int RSQ, int TILE, int STAGES, int MINB>
int launch_eval_cfg(int dev, int64_t Ns, int64_t Nt, const Real* pack, const Real* r_trg, Real* v_trg, const int* must_mask, Real ctx0, cudaStream_t stream, int min_chunk, int max_split) {
  using Cfg = EvalConfig<Real, KID, T, NW, R, RSQ, TILE, STAGES>;
  auto kern = eval_pairs_kernel<Real, KID, T, NW, R, RSQ, TILE, STAGES, MINB>;
  static OccCache occ;
  int ctas_per_sm;
  {
    std::lock_guard<std::mutex> lk(occ.mu);
    if (!occ.ctas[dev]) {
      CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
      int nb = 0;
      CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::THREADS, Cfg::SMEM_BYTES));
      if (nb < 1) return fail("eval kernel does not fit on an SM (threads %d, smem %zu)", Cfg::THREADS, Cfg::SMEM_BYTES);
      occ.ctas[dev] = nb;
    }
    ctas_per_sm = occ.ctas[dev];
  }
  const int slots = g.dev[dev].sms * ctas_per_sm;
  EvalPlan plan = plan_eval(Ns, Nt, Cfg::TPB, slots, min_chunk, max_split);
  if (plan.n_tblocks > 2147483647LL) return fail("too many target blocks");

  EvalArgs<Real> a;
  a.pack = pack;
  a.r_trg = r_trg;
  a.v_trg = v_trg;
  a.partial = nullptr;
  a.Ns = Ns;
  a.Nt = Nt;
  a.chunk = plan.chunk;
  a.n_split = plan.n_split;
  a.scale = scale_factor<Real>(KID);
  a.must_mask = must_mask;
  a.ctx0 = ctx0;
  constexpr int K1 = KerTraits<KID>::K1;
  int n_launches = 1;
  DevBuf partial;  // stream-ordered free on every exit path
  if (plan.n_split > 1) {
    if (partial.alloc(sizeof(Real) * (size_t)plan.n_split * Nt * K1, stream)) return 1;
    a.partial = (Real*)partial.p;
  }
  dim3 grid((unsigned)plan.n_tblocks, (unsigned)plan.n_split);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(a);
  CU_TRY(cudaGetLastError());
  if (plan.n_split > 1) {
    const int64_t n = Nt * K1;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 4096);
    reduce_partials_kernel<Real><<<blocks, 256, 0, stream>>>(n, plan.n_split, a.partial, a.scale, v_trg);
    CU_TRY(cudaGetLastError());
    n_launches = 2;
  }
  tl.grid_x = (int)grid.x;
  tl.grid_y = (int)grid.y;
  tl.threads = Cfg::THREADS;
  tl.smem = (int)Cfg::SMEM_BYTES;
  tl.ctas_per_sm = ctas_per_sm;
  tl.n_launches = n_launches;
  tl.wave_eff = plan.wave_eff;
  return 0;
}

// Chunk-streaming path: the packed sources go through the constant bank one chunk segment at a time (device to device,
// stream ordered), one eval_pairs_const_kernel launch over all targets per chunk.  Consecutive chunks use different
// slices of the bank on different streams, so a launch's ramp and tail overlap its neighbours' (eval_const_kernel.cuh).

template <class Real> int const_banks_wanted() {
  return (int)std::max<int64_t>(1, std::min<int64_t>(sizeof(Real) == 8 ? g.opt.const_banks : g.opt.const_banks_fp32, CONST_MAX_BANKS));
}
template <class Real> int64_t const_target_blocks(int64_t Nt) {
  using Tn = Tune<Real>;
  return (Nt + Tn::CONST_NT * Tn::CONST_T - 1) / (Tn::CONST_NT * Tn::CONST_T);
}

// `stage`: the packed sources in chunk segments (ConstGeom); `flags`: the probe's two ints, or nullptr when it did not run
template <class Real, int KID, int RSQ>
int launch_eval_const(int dev, const ConstGeom& cg, bool use_bits, int64_t Ns, int64_t Nt, const Real* stage, const Real* r_trg, Real* v_trg, const int* flags, Real ctx0, cudaStream_t stream) {
  using Tn = Tune<Real>;
  constexpr int T = Tn::CONST_T, NT = Tn::CONST_NT;
  constexpr int K1 = KerTraits<KID>::K1;
  auto kern = eval_pairs_const_kernel<Real, KID, T, NT, RSQ, Tn::CONST_U, Tn::CONST_MINB>;
  const int64_t n_tb = const_target_blocks<Real>(Nt);
  if (n_tb > 2147483647LL) return fail("too many target blocks");
  DeviceInfo& di = g.dev[dev];
  const int banks = cg.banks;
  // Cluster split of a chunk along grid.y when target blocks alone cannot fill the SMs.  With one bank every launch
  // stands alone and wants several CTAs per SM; with overlapping launches one CTA slot's worth per SM is enough.
  const int64_t want_ctas = (banks > 1 ? g.opt.const_ctas_per_sm_overlap : g.opt.const_ctas_per_sm) * (int64_t)di.sms;
  int cl = 1;
  while (cl < 8 && n_tb * cl < want_ctas) cl *= 2;
  ConstEvalArgs<Real> a = {};
  a.r_trg = r_trg;
  a.Nt = Nt;
  a.allow_unmasked = flags != nullptr;
  a.ctx0 = ctx0;
  a.scale = scale_factor<Real>(KID);
  a.zero = 0;
  a.rsq_k = (Real)rsq_poly_k<RSQ>();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)n_tb, (unsigned)cl);
  cfg.blockDim = dim3(NT);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = (unsigned)cl;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cl > 1 ? 1 : 0;

  std::lock_guard<std::mutex> lk(di.const_mu);
  if (!di.const_done) {
    CU_TRY(cudaEventCreateWithFlags(&di.const_done, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&di.const_fork, cudaEventDisableTiming));
    for (int b = 1; b < CONST_MAX_BANKS; b++) {
      CU_TRY(cudaStreamCreateWithFlags(&di.bank_stream[b], cudaStreamNonBlocking));
      CU_TRY(cudaEventCreateWithFlags(&di.bank_done[b], cudaEventDisableTiming));
    }
  } else {
    CU_TRY(cudaStreamWaitEvent(stream, di.const_done, 0));  // the previous user of the constant bank, on any stream
  }
  // Whatever happens below, the shared per-device state must be left consistent for the next caller (on any stream):
  // the bank streams joined back into `stream`, the scratch buffer freed behind everything that uses it, and const_done
  // recorded behind the last reader of c_src_pack / c_eval_must_mask.  The guard does that on every exit path.
  struct Join {
    DeviceInfo& di;
    cudaStream_t stream;
    int banks;
    bool forked = false, joined = false;
    void* alt = nullptr;
    void join() {
      if (forked && !joined)
        for (int b = 1; b < banks; b++)
          if (cudaEventRecord(di.bank_done[b], di.bank_stream[b]) == cudaSuccess) cudaStreamWaitEvent(stream, di.bank_done[b], 0);
      joined = true;
    }
    ~Join() {
      join();
      if (alt) cudaFreeAsync(alt, stream);
      cudaEventRecord(di.const_done, stream);
    }
  } guard{di, stream, banks};
  if (flags) CU_TRY(cudaMemcpyToSymbolAsync(c_eval_must_mask, flags, 2 * sizeof(int), 0, cudaMemcpyDeviceToDevice, stream));
  // banks 1.. accumulate into their own zeroed buffers (launches of different banks overlap and must not race on v_trg)
  Real* alt = nullptr;
  const size_t nv = (size_t)Nt * K1;
  if (banks > 1) {
    CU_TRY(cudaMallocAsync((void**)&alt, sizeof(Real) * nv * (banks - 1), stream));
    guard.alt = alt;
    CU_TRY(cudaMemsetAsync(alt, 0, sizeof(Real) * nv * (banks - 1), stream));
    CU_TRY(cudaEventRecord(di.const_fork, stream));
    guard.forked = true;
    for (int b = 1; b < banks; b++) CU_TRY(cudaStreamWaitEvent(di.bank_stream[b], di.const_fork, 0));
  }
  int launches = 0;
  const int64_t seg_reals = cg.seg_bytes / (int64_t)sizeof(Real);
  for (int64_t i = 0; i < cg.n_chunks; i++) {
    const int b = (int)(i % banks);
    cudaStream_t st = b ? di.bank_stream[b] : stream;
    const int n = (int)std::min<int64_t>(cg.chunk, Ns - i * cg.chunk);
    // one copy per launch: the chunk's records and, behind them, its coincidence bits
    const size_t bytes = use_bits ? (size_t)cg.seg_bytes : (size_t)n * cg.rec_bytes;
    CU_TRY(cudaMemcpyToSymbolAsync(c_src_pack, stage + i * seg_reals, bytes, (size_t)b * cg.bank_stride, cudaMemcpyDeviceToDevice, st));
    a.v_trg = b ? alt + (size_t)(b - 1) * nv : v_trg;
    a.n_rec = n;
    a.sub_len = (n + cl - 1) / cl;
    a.full_ranks = n / a.sub_len;
    a.tail_len = n % a.sub_len;
    a.bank_off = b * (cg.bank_stride / (int)sizeof(Real));
    a.bits_off = use_bits ? (b * cg.bank_stride + cg.chunk * cg.rec_bytes) / 4 : -1;
    cfg.stream = st;
    CU_TRY(cudaLaunchKernelEx(&cfg, kern, a));
    launches++;
  }
  if (banks > 1) {
    guard.join();
    const int blocks = (int)std::min<int64_t>(((int64_t)nv + 255) / 256, 4096);
    reduce_partials_kernel<Real><<<blocks, 256, 0, stream>>>((int64_t)nv, banks - 1, alt, (Real)1, v_trg);  // bank order: deterministic
    CU_TRY(cudaGetLastError());
    launches++;
  }
  tl.grid_x = (int)n_tb;
  tl.grid_y = cl;
  tl.threads = NT;
  tl.smem = (int)(sizeof(Real) * T * K1 * NT);
  tl.ctas_per_sm = 0;
  tl.n_launches = launches;
  tl.wave_eff = 0;
  return 0;
}

template <class Real, int KID, int RSQ>
int eval_device_t(int dev, int64_t Ns, int64_t Nt, const Real* r_src, const Real* n_src, const Real* v_src, const Real* r_trg, Real* v_trg, Real ctx0, cudaStream_t stream) {
  using Tn = Tune<Real>;
  constexpr int REC = RecTraits<Real, KID>::REC;
  const bool fp32_const = g.opt.const_fp32 != 0;
  bool use_const = Nt >= g.opt.const_min_targets && (sizeof(Real) == 8 || fp32_const);
  // coincidence probe (big problems only): lets the pair kernels skip the r2 == 0 mask where it cannot matter
  const bool probe = g.opt.probe_min_pairs > 0 && (double)Ns * (double)Nt >= (double)g.opt.probe_min_pairs;
  ConstGeom cg = {};
  // the bit rows are part of the layout whether or not this call fills them: the chunk length, and with it the summation
  // order, must not depend on the probe options (masked and probed runs are bit-identical)
  if (use_const) cg = const_geometry<Real, KID>(Ns, const_target_blocks<Real>(Nt), const_banks_wanted<Real>(), true);
  // FP64: a launch of the constant-bank path must carry enough work to pay for itself.  Measured on B200 at 2^15 targets
  // (profiles/r2_small_shard_paths.md): the ring wins by 7 % / 5 % for Laplace3D_FxU / Stokes3D_DxU (3.6e8 / 4.0e8 flop per
  // launch), the constant bank by 1 % / 4 % for Laplace3D_FxdU / Stokes3D_FxU (6.0e8 / 6.4e8); from 2^16 it wins for all.
  if (use_const && sizeof(Real) == 8 && (double)cg.chunk * (double)Nt * KerTraits<KID>::F_ALG < (double)g.opt.const_min_launch_flop) use_const = false;
  const bool use_bits = use_const && cg.row_words && probe && g.opt.probe_per_block != 0;
  // packed sources: contiguous records, or chunk segments (records + coincidence bits) for the constant-bank path
  const size_t pack_bytes = use_const ? (size_t)cg.n_chunks * cg.seg_bytes : sizeof(Real) * (size_t)Ns * REC;
  DevBuf pack_buf, table_buf;  // freed (stream-ordered) on every exit path, after everything enqueued below
  if (pack_buf.alloc(pack_bytes, stream)) return 1;
  Real* pack = (Real*)pack_buf.p;
  if (use_bits) CU_TRY(cudaMemsetAsync(pack, 0, pack_bytes, stream));  // the bit rows start clear
  {
    const int blocks = (int)std::min<int64_t>((Ns + 255) / 256, 8192);
    const int64_t chunk_len = use_const ? cg.chunk : std::max<int64_t>(Ns, 1);
    pack_sources_kernel<Real, KID><<<blocks, 256, 0, stream>>>(Ns, r_src, n_src, v_src, pack, chunk_len, use_const ? cg.seg_bytes / (int64_t)sizeof(Real) : 0);
    CU_TRY(cudaGetLastError());
  }
  int* flags = nullptr;
  unsigned long long* table = nullptr;
  int n_probe_launches = 0;
  if (probe) {
    uint64_t slots = 1024;
    while (slots < 2 * (uint64_t)Ns) slots <<= 1;
    const bool bits = use_bits;
    // [slots keys | 2 flags (one 64-bit word) | pad] [cmin | cmax]
    const size_t key_bytes = (slots + 2) * sizeof(unsigned long long);
    if (table_buf.alloc(key_bytes + (bits ? 2 * slots * sizeof(unsigned) : 0), stream)) return 1;
    table = (unsigned long long*)table_buf.p;
    CU_TRY(cudaMemsetAsync(table, 0, key_bytes + (bits ? 2 * slots * sizeof(unsigned) : 0), stream));
    flags = reinterpret_cast<int*>(table + slots);
    CoincidenceBits cb = {};
    if (bits) {
      cb.cmin = reinterpret_cast<unsigned*>(table + slots + 2);
      cb.cmax = cb.cmin + slots;
      CU_TRY(cudaMemsetAsync(cb.cmin, 0xFF, slots * sizeof(unsigned), stream));
      cb.bits = reinterpret_cast<unsigned*>(pack) + (int64_t)cg.chunk * cg.rec_bytes / 4;
      cb.row_stride = cg.seg_bytes / 4;
      cb.chunk_len = cg.chunk;
      cb.trg_per_blk = Tn::CONST_NT * Tn::CONST_T;
    }
    coincidence_insert_kernel<Real><<<(int)std::min<int64_t>((Ns + 255) / 256, 4096), 256, 0, stream>>>(Ns, r_src, table, slots - 1, cb);
    coincidence_probe_kernel<Real><<<(int)std::min<int64_t>((Nt + 255) / 256, 4096), 256, 0, stream>>>(Nt, r_trg, table, slots - 1, flags, cb);
    CU_TRY(cudaGetLastError());
    n_probe_launches = 2;
  }
  int rc;
  if (use_const)
    rc = launch_eval_const<Real, KID, RSQ>(dev, cg, use_bits, Ns, Nt, pack, r_trg, v_trg, flags, ctx0, stream);
  else if (Nt >= g.opt.wide_min_targets && sizeof(Real) == 8 && KerTraits<KID>::K1 <= 4)  // 9 accumulators x 4 targets would spill
    rc = launch_eval_cfg<Real, KID, (KerTraits<KID>::K1 <= 4 ? Tn::WIDE_T : Tn::T), Tn::WIDE_NW, 1, RSQ, Tn::WIDE_TILE, Tn::WIDE_STAGES, Tn::WIDE_MINB>(dev, Ns, Nt, pack, r_trg, v_trg, flags, ctx0, stream, (int)g.opt.ring_min_tiles * Tn::WIDE_TILE, (int)g.opt.ring_max_split);
  else if (Nt <= g.opt.few_max_targets)
    rc = launch_eval_cfg<Real, KID, Tn::FEW_T, Tn::FEW_NW, Tn::FEW_NW, RSQ, Tn::FEW_TILE, Tn::FEW_STAGES, Tn::FEW_MINB>(dev, Ns, Nt, pack, r_trg, v_trg, flags, ctx0, stream, Tn::FEW_TILE, 1024);
  else
    rc = launch_eval_cfg<Real, KID, Tn::T, Tn::NW, 1, RSQ, Tn::TILE, Tn::STAGES, Tn::MINB>(dev, Ns, Nt, pack, r_trg, v_trg, flags, ctx0, stream, (int)g.opt.ring_min_tiles * Tn::TILE, (int)g.opt.ring_max_split);
  tl.n_launches += n_probe_launches;
  if (rc) return rc;
  tl.n_launches += 1;  // the pack kernel
  return 0;
}

template <class Real, int KID>
int eval_device_k(int dev, int digits, int64_t Ns, int64_t Nt, const Real* r_src, const Real* n_src, const Real* v_src, const Real* r_trg, Real* v_trg, Real ctx0, cudaStream_t stream) {
  int mode = rsq_mode_for_digits<Real>(digits);
  // Helmholtz kernels feed r = r2 * rinv into sin/cos and cancel leading terms (HelmholtzDiff): the 2^-22 seed alone
  // leaves FP32 results further from the reference than its own AVX-512 recipe (seed + 1 Newton step at digits 5-8)
  if (KerNeedsCtx<KID>::value && sizeof(Real) == 4 && mode == RSQ_SEED && (digits < 0 || digits >= 5)) mode = RSQ_NEWTON;
  switch (mode) {
    case RSQ_SEED: return eval_device_t<Real, KID, RSQ_SEED>(dev, Ns, Nt, r_src, n_src, v_src, r_trg, v_trg, ctx0, stream);
    case RSQ_NEWTON: return eval_device_t<Real, KID, RSQ_NEWTON>(dev, Ns, Nt, r_src, n_src, v_src, r_trg, v_trg, ctx0, stream);
    default:
      if (sizeof(Real) == 4) return eval_device_t<Real, KID, RSQ_NEWTON>(dev, Ns, Nt, r_src, n_src, v_src, r_trg, v_trg, ctx0, stream);
      return eval_device_t<Real, KID, (sizeof(Real) == 8 ? RSQ_CUBIC : RSQ_NEWTON)>(dev, Ns, Nt, r_src, n_src, v_src, r_trg, v_trg, ctx0, stream);
  }
}

template <class Real>
int eval_device_r(int dev, int kid, int digits, int64_t Ns, int64_t Nt, const void* r_src, const void* n_src, const void* v_src, const void* r_trg, void* v_trg, double ctx0, cudaStream_t stream) {
#define SCTL_CUDA_CASE(K) \
  case K: return eval_device_k<Real, K>(dev, digits, Ns, Nt, (const Real*)r_src, (const Real*)n_src, (const Real*)v_src, (const Real*)r_trg, (Real*)v_trg, (Real)ctx0, stream);
  switch (kid) {
    SCTL_CUDA_CASE(LAPLACE3D_FXU)
    SCTL_CUDA_CASE(LAPLACE3D_DXU)
    SCTL_CUDA_CASE(LAPLACE3D_FXDU)
    SCTL_CUDA_CASE(STOKES3D_FXU)
    SCTL_CUDA_CASE(STOKES3D_DXU)
    SCTL_CUDA_CASE(STOKES3D_FXT)
    SCTL_CUDA_CASE(STOKES3D_FSXU)
    SCTL_CUDA_CASE(STOKES3D_FXUP)
    SCTL_CUDA_CASE(LAPLACE3D_FXD2U)
    SCTL_CUDA_CASE(LAPLACE3D_DXDU)
    SCTL_CUDA_CASE(BIOTSAVART3D_FXU)
    SCTL_CUDA_CASE(BIOTSAVART3D_FXDU)
    SCTL_CUDA_CASE(HELMHOLTZ3D_FXU)
    SCTL_CUDA_CASE(HELMHOLTZ3D_DXU)
    SCTL_CUDA_CASE(HELMHOLTZ3D_FXDU)
    SCTL_CUDA_CASE(HELMHOLTZDIFF3D_FXDU)
    default: return fail("unknown kernel id %d", kid);
  }
#undef SCTL_CUDA_CASE
}

int check_common(int kid, int real_bytes, int64_t Ns, int64_t Nt) {
  if (kid < 0 || kid >= NUM_KERNELS) return fail("unknown kernel id %d", kid);
  if (real_bytes != 8 && real_bytes != 4) return fail("real_bytes must be 8 (double) or 4 (float), got %d", real_bytes);
  if (Ns < 0 || Nt < 0) return fail("negative size (Ns=%lld, Nt=%lld)", (long long)Ns, (long long)Nt);
  return 0;
}

// The kernel context (uKernel ctx_ptr, generic-kernel.hpp:79): for the Helmholtz kernels a HOST pointer to one Real, the
// wavenumber (kernel_functions.hpp:484-488); copied by value into the launch.  Other kernels ignore it.
int decode_ctx(int kid, int real_bytes, const void* ctx, int64_t ctx_bytes, double* ctx0) {
  *ctx0 = 0;
  if (kid < HELMHOLTZ3D_FXU) return 0;
  if (!ctx) return fail("kernel %s needs its context: ctx must point to the wavenumber (one %s)", kernel_info(kid).name, real_bytes == 8 ? "double" : "float");
  if (ctx_bytes != 0 && ctx_bytes != real_bytes) return fail("ctx_bytes (%lld) must equal real_bytes (%d) for kernel %s", (long long)ctx_bytes, real_bytes, kernel_info(kid).name);
  *ctx0 = real_bytes == 8 ? *(const double*)ctx : (double)*(const float*)ctx;
  return 0;
}

int eval_device_any(int dev, int kid, int real_bytes, int digits, int64_t Ns, int64_t Nt, const void* r_src, const void* n_src, const void* v_src, const void* r_trg, void* v_trg, double ctx0, cudaStream_t stream) {
  const KernelInfo ki = kernel_info(kid);
  tl.n_launches = 0;
  if (Ns == 0 || Nt == 0) return 0;  // nothing to add
  if (!r_src || !v_src || !r_trg || !v_trg) return fail("null array pointer");
  if (ki.nd && !n_src) return fail("kernel %s needs source normals (n_src is NULL)", ki.name);
  if (real_bytes == 8) return eval_device_r<double>(dev, kid, digits, Ns, Nt, r_src, n_src, v_src, r_trg, v_trg, ctx0, stream);
  return eval_device_r<float>(dev, kid, digits, Ns, Nt, r_src, n_src, v_src, r_trg, v_trg, ctx0, stream);
}

// ---------------------------------------------------------------------------------------------
// KernelMatrix
// ---------------------------------------------------------------------------------------------
// CTAs per SM a large matrix is cut into, for K1 > 1 and K1 == 1 kernels (tuning builds: tools/ab_build.sh;
// profiles/r2l_kmat_grid_ab.md: the K1 == 1 kernels want fewer, longer CTAs -- more rows per target load -- FP32 odd
// pitches +15 %, nothing lost elsewhere; K1 > 1 is flat within 3 % between 32 and 128)
#ifndef SCTL_TUNE_KMAT_CTAS_PER_SM
#define SCTL_TUNE_KMAT_CTAS_PER_SM 64
#endif
#ifndef SCTL_TUNE_KMAT_CTAS_PER_SM_K1
#define SCTL_TUNE_KMAT_CTAS_PER_SM_K1 32
#endif
template <class Real, int KID, int MODE>
int kmat_launch(int rsq, dim3 grid, int64_t Ns, int64_t Nt, int64_t rows_per_cta, const Real* r_src, const Real* n_src, const Real* r_trg, Real* M, int64_t ldM, Real ctx0, cudaStream_t stream) {
  const Real scale = scale_factor<Real>(KID);
  switch (rsq) {
    case RSQ_SEED: kernel_matrix_kernel<Real, KID, RSQ_SEED, MODE><<<grid, KMAT_BT, 0, stream>>>(Ns, Nt, rows_per_cta, r_src, n_src, r_trg, M, ldM, scale, ctx0); break;
    case RSQ_NEWTON: kernel_matrix_kernel<Real, KID, RSQ_NEWTON, MODE><<<grid, KMAT_BT, 0, stream>>>(Ns, Nt, rows_per_cta, r_src, n_src, r_trg, M, ldM, scale, ctx0); break;
    default: kernel_matrix_kernel<Real, KID, (sizeof(Real) == 8 ? RSQ_CUBIC : RSQ_NEWTON), MODE><<<grid, KMAT_BT, 0, stream>>>(Ns, Nt, rows_per_cta, r_src, n_src, r_trg, M, ldM, scale, ctx0); break;
  }
  CU_TRY(cudaGetLastError());
  return 0;
}

template <class Real, int KID>
int kmat_device_k(int dev, int digits, int64_t Ns, int64_t Nt, const Real* r_src, const Real* n_src, const Real* r_trg, Real* M, int64_t ldM, Real ctx0, cudaStream_t stream) {
  constexpr int T = KmatShape<Real, KID, false>::T;  // the same for both paths
  // A CTA keeps its KMAT_BT * T targets in registers and walks a contiguous range of source rows.  Enough CTAs for several
  // waves of the machine (8 - 12 CTAs of 128 threads are resident per SM) when the matrix is that large, so that the last
  // wave is short against the whole; never fewer CTAs than one per (target block, row) allows -- small blocks (hundreds x
  // hundreds, boundary_integral.txx:1191-1211) are latency-bound and want every row on its own CTA.
  const int64_t n_tb = (Nt + KMAT_BT * T - 1) / (KMAT_BT * T);
  if (n_tb > 2147483647LL) return fail("too many target blocks");
  constexpr int per_sm = KerTraits<KID>::K1 == 1 ? SCTL_TUNE_KMAT_CTAS_PER_SM_K1 : SCTL_TUNE_KMAT_CTAS_PER_SM;
  const int64_t want_y = std::max<int64_t>(1, ((int64_t)g.dev[dev].sms * per_sm + n_tb - 1) / n_tb);
  const int64_t gy0 = std::min<int64_t>({Ns, want_y, 65535});
  const int64_t rows_per_cta = (Ns + gy0 - 1) / gy0;
  dim3 grid((unsigned)n_tb, (unsigned)((Ns + rows_per_cta - 1) / rows_per_cta));
  const int rsq = rsq_mode_for_digits<Real>(digits);
  if constexpr (KerTraits<KID>::K1 == 1) {
    // vectors go straight from registers to M; rows that do not all start on a 16-byte boundary take the lane-shifted form
    constexpr int V = KmatShape<Real, KID, true>::V;
    if ((uintptr_t)M % 16 == 0 && ldM % V == 0) return kmat_launch<Real, KID, KMAT_DIRECT>(rsq, grid, Ns, Nt, rows_per_cta, r_src, n_src, r_trg, M, ldM, ctx0, stream);
    return kmat_launch<Real, KID, KMAT_SHIFTED>(rsq, grid, Ns, Nt, rows_per_cta, r_src, n_src, r_trg, M, ldM, ctx0, stream);
  } else {
    return kmat_launch<Real, KID, KMAT_STAGED>(rsq, grid, Ns, Nt, rows_per_cta, r_src, n_src, r_trg, M, ldM, ctx0, stream);  // warp-private stage
  }
}

template <class Real>
int kmat_device_r(int dev, int kid, int digits, int64_t Ns, int64_t Nt, const void* r_src, const void* n_src, const void* r_trg, void* M, int64_t ldM, double ctx0, cudaStream_t stream) {
#define SCTL_CUDA_CASE(K) \
  case K: return kmat_device_k<Real, K>(dev, digits, Ns, Nt, (const Real*)r_src, (const Real*)n_src, (const Real*)r_trg, (Real*)M, ldM, (Real)ctx0, stream);
  switch (kid) {
    SCTL_CUDA_CASE(LAPLACE3D_FXU)
    SCTL_CUDA_CASE(LAPLACE3D_DXU)
    SCTL_CUDA_CASE(LAPLACE3D_FXDU)
    SCTL_CUDA_CASE(STOKES3D_FXU)
    SCTL_CUDA_CASE(STOKES3D_DXU)
    SCTL_CUDA_CASE(STOKES3D_FXT)
    SCTL_CUDA_CASE(STOKES3D_FSXU)
    SCTL_CUDA_CASE(STOKES3D_FXUP)
    SCTL_CUDA_CASE(LAPLACE3D_FXD2U)
    SCTL_CUDA_CASE(LAPLACE3D_DXDU)
    SCTL_CUDA_CASE(BIOTSAVART3D_FXU)
    SCTL_CUDA_CASE(BIOTSAVART3D_FXDU)
    SCTL_CUDA_CASE(HELMHOLTZ3D_FXU)
    SCTL_CUDA_CASE(HELMHOLTZ3D_DXU)
    SCTL_CUDA_CASE(HELMHOLTZ3D_FXDU)
    SCTL_CUDA_CASE(HELMHOLTZDIFF3D_FXDU)
    default: return fail("unknown kernel id %d", kid);
  }
#undef SCTL_CUDA_CASE
}

int kmat_device_any(int dev, int kid, int real_bytes, int digits, int64_t Ns, int64_t Nt, const void* r_src, const void* n_src, const void* r_trg, void* M, int64_t ldM, double ctx0, cudaStream_t stream) {
  const KernelInfo ki = kernel_info(kid);
  if (Ns == 0 || Nt == 0) return 0;
  if (!r_src || !r_trg || !M) return fail("null array pointer");
  if (ki.nd && !n_src) return fail("kernel %s needs source normals (n_src is NULL)", ki.name);
  if (ldM < Nt * ki.k1) return fail("ldM (%lld) smaller than Nt*K1 (%lld)", (long long)ldM, (long long)(Nt * ki.k1));
  if (real_bytes == 8) return kmat_device_r<double>(dev, kid, digits, Ns, Nt, r_src, n_src, r_trg, M, ldM, ctx0, stream);
  return kmat_device_r<float>(dev, kid, digits, Ns, Nt, r_src, n_src, r_trg, M, ldM, ctx0, stream);
}

// ---------------------------------------------------------------------------------------------
// small device utilities
// ---------------------------------------------------------------------------------------------
template <class Real, int MODE> __global__ void rsqrt_test_kernel(int64_t n, const Real* x, Real* y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = rsqrt_masked<MODE>(x[i]);
}

template <class Real> __global__ void __launch_bounds__(256) fma_peak_kernel(int iters, Real* out, Real seed) {
  Real a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const Real m = (Real)0.999999, c = (Real)1e-6;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const Real s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == (Real)-1) out[0] = s;  // never true; keeps the chains alive
}

template <class Real> int fma_peak_t(double ms, double* tflops) {
  int dev;
  CU_TRY(cudaGetDevice(&dev));
  if (ensure_device(dev)) return 1;
  cudaStream_t st = tl.res[dev].stream;
  Real* out;
  CU_TRY(cudaMallocAsync((void**)&out, sizeof(Real), st));
  const int blocks = g.dev[dev].sms * 8, threads = 256;
  int iters = 2000;
  float t = 0;
  double best = 0;
  for (int rep = 0; rep < 6; rep++) {
    CU_TRY(cudaEventRecord(tl.res[dev].ev[0], st));
    fma_peak_kernel<Real><<<blocks, threads, 0, st>>>(iters, out, (Real)0.5);
    CU_TRY(cudaEventRecord(tl.res[dev].ev[1], st));
    CU_TRY(cudaEventSynchronize(tl.res[dev].ev[1]));
    CU_TRY(cudaEventElapsedTime(&t, tl.res[dev].ev[0], tl.res[dev].ev[1]));
    const double flop = 2.0 * 64.0 * iters * (double)blocks * threads;
    if (rep > 0) best = std::max(best, flop / (t * 1e-3) * 1e-12);
    if (t < ms && rep < 4) iters = (int)std::min<double>(2.0e6, iters * std::max(1.5, ms / std::max(t, 0.01f)));
  }
  CU_TRY(cudaFreeAsync(out, st));
  CU_TRY(cudaStreamSynchronize(st));
  *tflops = best;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// host-pointer paths
// ---------------------------------------------------------------------------------------------
int current_device_ready(int* dev) {
  if (visible_devices() < 1) return fail("no CUDA device available (libsctl_cuda has no CPU fallback)");
  CU_TRY(cudaGetDevice(dev));
  if (*dev >= MAX_DEV) return fail("device index %d too large", *dev);
  {
    std::lock_guard<std::mutex> lk(g.mu);
    if (g.n_gpus == 0) g.n_gpus = 1;
  }
  return ensure_device(*dev);
}

// Small calls (all arrays together below SMALL_ARENA bytes) are latency-bound: five pageable uploads, five pool
// allocations and a pageable download cost more than the kernel.  They go through a per-thread arena instead: the inputs
// are gathered into ONE pinned buffer, uploaded with one copy into a persistent device buffer, and the result comes back
// through the same pinned buffer.  (BoundaryIntegralOp's near-field setup and FMM leaf interactions issue thousands of
// such calls, many of them concurrently from OpenMP threads: boundary_integral.txx:1169-1211.)
constexpr size_t SMALL_ARENA = 1 << 20;
inline size_t arena_align(size_t n) { return (n + 255) & ~(size_t)255; }

int ensure_arena(int dev) {
  if (tl.res[dev].arena_h) return 0;
  CU_TRY(cudaHostAlloc((void**)&tl.res[dev].arena_h, SMALL_ARENA, cudaHostAllocDefault));
  CU_TRY(cudaMalloc((void**)&tl.res[dev].arena_d, SMALL_ARENA));
  return 0;
}

int eval_host_small(int dev, int kid, int real_bytes, int digits, int64_t Ns, int64_t Nt, const char* r_src, const char* n_src, const char* v_src, const char* r_trg,
                    char* v_trg, double ctx0) {
  const KernelInfo ki = kernel_info(kid);
  const size_t rb = (size_t)real_bytes;
  const size_t b_rs = rb * Ns * 3, b_vs = rb * Ns * ki.k0, b_ns = ki.nd ? rb * Ns * ki.nd : 0, b_rt = rb * Nt * 3, b_vt = rb * Nt * ki.k1;
  const size_t o_vt = 0, o_rs = o_vt + arena_align(b_vt), o_vs = o_rs + arena_align(b_rs), o_ns = o_vs + arena_align(b_vs), o_rt = o_ns + arena_align(b_ns);
  const size_t total = o_rt + arena_align(b_rt);
  CU_TRY(cudaSetDevice(dev));
  if (ensure_device(dev) || ensure_arena(dev)) return 1;
  cudaStream_t st = tl.res[dev].stream;
  char *h = tl.res[dev].arena_h, *d = tl.res[dev].arena_d;
  memcpy(h + o_vt, v_trg, b_vt);
  memcpy(h + o_rs, r_src, b_rs);
  memcpy(h + o_vs, v_src, b_vs);
  if (b_ns) memcpy(h + o_ns, n_src, b_ns);
  memcpy(h + o_rt, r_trg, b_rt);
  CU_TRY(cudaEventRecord(tl.res[dev].ev[0], st));
  CU_TRY(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaEventRecord(tl.res[dev].ev[1], st));
  if (eval_device_any(dev, kid, real_bytes, digits, Ns, Nt, d + o_rs, b_ns ? d + o_ns : nullptr, d + o_vs, d + o_rt, d + o_vt, ctx0, st)) return 1;
  CU_TRY(cudaEventRecord(tl.res[dev].ev[2], st));
  CU_TRY(cudaMemcpyAsync(h + o_vt, d + o_vt, b_vt, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaEventRecord(tl.res[dev].ev[3], st));
  CU_TRY(cudaStreamSynchronize(st));
  memcpy(v_trg, h + o_vt, b_vt);
  float a = 0, b = 0, c = 0;
  CU_TRY(cudaEventElapsedTime(&a, tl.res[dev].ev[0], tl.res[dev].ev[1]));
  CU_TRY(cudaEventElapsedTime(&b, tl.res[dev].ev[1], tl.res[dev].ev[2]));
  CU_TRY(cudaEventElapsedTime(&c, tl.res[dev].ev[2], tl.res[dev].ev[3]));
  tl.h2d_ms = a;
  tl.kernel_ms = b;
  tl.d2h_ms = c;
  return 0;
}

// One GPU: upload everything, run, download the target range [t0, t1).
int eval_host_one(int dev, int kid, int real_bytes, int digits, int64_t Ns, int64_t t0, int64_t t1, const char* r_src, const char* n_src, const char* v_src,
                  const char* r_trg, char* v_trg, double ctx0, bool timing) {
  const KernelInfo ki = kernel_info(kid);
  const size_t rb = (size_t)real_bytes;
  const int64_t nt = t1 - t0;
  CU_TRY(cudaSetDevice(dev));
  if (ensure_device(dev)) return 1;
  cudaStream_t st = tl.res[dev].stream;
  DevBuf d_rs, d_ns, d_vs, d_rt, d_vt;
  if (d_rs.alloc(rb * Ns * 3, st) || d_vs.alloc(rb * Ns * ki.k0, st) || d_rt.alloc(rb * nt * 3, st) || d_vt.alloc(rb * nt * ki.k1, st)) return 1;
  if (ki.nd && d_ns.alloc(rb * Ns * ki.nd, st)) return 1;
  if (timing) CU_TRY(cudaEventRecord(tl.res[dev].ev[0], st));
  CU_TRY(cudaMemcpyAsync(d_rs.p, r_src, rb * Ns * 3, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(d_vs.p, v_src, rb * Ns * ki.k0, cudaMemcpyHostToDevice, st));
  if (ki.nd) CU_TRY(cudaMemcpyAsync(d_ns.p, n_src, rb * Ns * ki.nd, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(d_rt.p, r_trg + rb * t0 * 3, rb * nt * 3, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(d_vt.p, v_trg + rb * t0 * ki.k1, rb * nt * ki.k1, cudaMemcpyHostToDevice, st));
  if (timing) CU_TRY(cudaEventRecord(tl.res[dev].ev[1], st));
  if (eval_device_any(dev, kid, real_bytes, digits, Ns, nt, d_rs.p, d_ns.p, d_vs.p, d_rt.p, d_vt.p, ctx0, st)) return 1;
  if (timing) CU_TRY(cudaEventRecord(tl.res[dev].ev[2], st));
  CU_TRY(cudaMemcpyAsync(v_trg + rb * t0 * ki.k1, d_vt.p, rb * nt * ki.k1, cudaMemcpyDeviceToHost, st));
  if (timing) CU_TRY(cudaEventRecord(tl.res[dev].ev[3], st));
  CU_TRY(cudaStreamSynchronize(st));
  if (timing) {
    float a = 0, b = 0, c = 0;
    CU_TRY(cudaEventElapsedTime(&a, tl.res[dev].ev[0], tl.res[dev].ev[1]));
    CU_TRY(cudaEventElapsedTime(&b, tl.res[dev].ev[1], tl.res[dev].ev[2]));
    CU_TRY(cudaEventElapsedTime(&c, tl.res[dev].ev[2], tl.res[dev].ev[3]));
    tl.h2d_ms = a;
    tl.kernel_ms = b;
    tl.d2h_ms = c;
  }
  return 0;
}

// P GPUs of this process: GPU p uploads source slice p and target shard p; slices are replicated with one
// grouped set of ncclBroadcast (an all-gather with ragged counts) over NVLink; each GPU evaluates its shard
// against all sources and downloads its slice of v_trg.  No reduction, so per-target summation order is the
// single-GPU order.  (Reference analogue: ParticleFMM::EvalDirect's np-step ring, fmm-wrapper.txx:551-623.)
struct ShardResult {
  int rc = 0;
  std::string err;
  double h2d = 0, ker = 0, d2h = 0;
};

// All P shard threads meet here once, after their fallible local work (device selection, allocations, uploads) and
// before the first collective: if any of them failed, none enters the NCCL group -- otherwise the survivors would
// enqueue broadcasts that can never complete and block in cudaStreamSynchronize for ever.
struct ShardGate {
  explicit ShardGate(int P_) : P(P_) {}
  std::mutex mu;
  std::condition_variable cv;
  int P, arrived = 0;
  bool failed = false;
  bool arrive(bool ok) {
    std::unique_lock<std::mutex> lk(mu);
    if (!ok) failed = true;
    if (++arrived == P) cv.notify_all();
    else cv.wait(lk, [&] { return arrived == P; });
    return !failed;
  }
};

void eval_host_shard(int p, int P, int kid, int real_bytes, int digits, int64_t Ns, int64_t Nt, const char* r_src, const char* n_src, const char* v_src,
                     const char* r_trg, char* v_trg, double ctx0, ShardGate* gate, ShardResult* res) {
  const KernelInfo ki = kernel_info(kid);
  const size_t rb = (size_t)real_bytes;
  const int64_t t0 = Nt * p / P, t1 = Nt * (p + 1) / P, nt = t1 - t0;
  cudaStream_t st = g.mstream[p];
  cudaEvent_t* ev = g.mev[p];
  DevBuf d_rs, d_ns, d_vs, d_rt, d_vt;
  // everything that can fail without the other shards noticing
  auto before_collective = [&]() -> int {
    CU_TRY(cudaSetDevice(p));
    if (d_rs.alloc(rb * Ns * 3, st) || d_vs.alloc(rb * Ns * ki.k0, st) || d_rt.alloc(rb * std::max<int64_t>(nt, 1) * 3, st) ||
        d_vt.alloc(rb * std::max<int64_t>(nt, 1) * ki.k1, st))
      return 1;
    if (ki.nd && d_ns.alloc(rb * Ns * ki.nd, st)) return 1;
    CU_TRY(cudaEventRecord(ev[0], st));
    const int64_t s0 = Ns * p / P, s1 = Ns * (p + 1) / P;
    CU_TRY(cudaMemcpyAsync((char*)d_rs.p + rb * s0 * 3, r_src + rb * s0 * 3, rb * (s1 - s0) * 3, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync((char*)d_vs.p + rb * s0 * ki.k0, v_src + rb * s0 * ki.k0, rb * (s1 - s0) * ki.k0, cudaMemcpyHostToDevice, st));
    if (ki.nd) CU_TRY(cudaMemcpyAsync((char*)d_ns.p + rb * s0 * ki.nd, n_src + rb * s0 * ki.nd, rb * (s1 - s0) * ki.nd, cudaMemcpyHostToDevice, st));
    if (nt) {
      CU_TRY(cudaMemcpyAsync(d_rt.p, r_trg + rb * t0 * 3, rb * nt * 3, cudaMemcpyHostToDevice, st));
      CU_TRY(cudaMemcpyAsync(d_vt.p, v_trg + rb * t0 * ki.k1, rb * nt * ki.k1, cudaMemcpyHostToDevice, st));
    }
    return 0;
  };
  auto collective_and_after = [&]() -> int {
    // replicate the source slices: root q broadcasts its slice into the same offsets on every GPU
    const ncclDataType_t dt = real_bytes == 8 ? ncclFloat64 : ncclFloat32;
    ncclResult_t nr = g.nccl.GroupStart();
    for (int q = 0; q < P && nr == ncclSuccess; q++) {
      const int64_t q0 = Ns * q / P, q1 = Ns * (q + 1) / P;
      if (q1 == q0) continue;
      char* a = (char*)d_rs.p + rb * q0 * 3;
      nr = g.nccl.Broadcast(a, a, (size_t)(q1 - q0) * 3, dt, q, g.comms[p], st);
      if (nr != ncclSuccess) break;
      char* b = (char*)d_vs.p + rb * q0 * ki.k0;
      nr = g.nccl.Broadcast(b, b, (size_t)(q1 - q0) * ki.k0, dt, q, g.comms[p], st);
      if (nr != ncclSuccess) break;
      if (ki.nd) {
        char* c = (char*)d_ns.p + rb * q0 * ki.nd;
        nr = g.nccl.Broadcast(c, c, (size_t)(q1 - q0) * ki.nd, dt, q, g.comms[p], st);
      }
    }
    ncclResult_t ne = g.nccl.GroupEnd();
    if (nr != ncclSuccess || ne != ncclSuccess) return fail("NCCL broadcast failed: %s", g.nccl.GetErrorString(nr != ncclSuccess ? nr : ne));
    CU_TRY(cudaEventRecord(ev[1], st));
    if (nt && eval_device_any(p, kid, real_bytes, digits, Ns, nt, d_rs.p, d_ns.p, d_vs.p, d_rt.p, d_vt.p, ctx0, st)) return 1;
    CU_TRY(cudaEventRecord(ev[2], st));
    if (nt) CU_TRY(cudaMemcpyAsync(v_trg + rb * t0 * ki.k1, d_vt.p, rb * nt * ki.k1, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaEventRecord(ev[3], st));
    CU_TRY(cudaStreamSynchronize(st));
    float a = 0, b = 0, c = 0;
    CU_TRY(cudaEventElapsedTime(&a, ev[0], ev[1]));
    CU_TRY(cudaEventElapsedTime(&b, ev[1], ev[2]));
    CU_TRY(cudaEventElapsedTime(&c, ev[2], ev[3]));
    res->h2d = a;
    res->ker = b;
    res->d2h = c;
    return 0;
  };
  res->rc = before_collective();
  if (res->rc) res->err = tl_err;
  if (!gate->arrive(res->rc == 0)) {
    if (!res->rc) {
      res->rc = 1;
      res->err = "cancelled: another GPU's shard failed before the broadcast";
    }
    cudaStreamSynchronize(st);  // the uploads read the caller's arrays
    return;
  }
  res->rc = collective_and_after();
  if (res->rc) res->err = tl_err;
}

// below this many pairs one GPU finishes before the extra launches and the broadcast pay off
constexpr double MULTI_GPU_MIN_PAIRS = 1.0e9;

}  // namespace

#include "direct_session.inc"
#include "near_session.inc"
#include "bio_session.inc"

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int sctl_cuda_device_count(void) { return visible_devices(); }

const char* sctl_cuda_last_error(void) { return tl_err.c_str(); }

int sctl_cuda_num_gpus(void) {
  std::lock_guard<std::mutex> lk(g.mu);
  return g.n_gpus;
}

namespace {
// caller holds g.multi_mu
int finalize_locked() {
  std::lock_guard<std::mutex> lk(g.mu);
  if (g.have_comms) {
    for (int i = 0; i < g.n_gpus; i++)
      if (g.comms[i]) g.nccl.CommDestroy(g.comms[i]);
    memset(g.comms, 0, sizeof(g.comms));
    g.have_comms = false;
  }
  g.n_gpus = 0;
  // resources parked by threads that have ended
  ResPool& pool = res_pool();
  std::lock_guard<std::mutex> plk(pool.mu);
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess) cur = -1;
  for (int d = 0; d < MAX_DEV; d++) {
    if (pool.idle[d].empty()) continue;
    if (cudaSetDevice(d) == cudaSuccess)
      for (DevRes& r : pool.idle[d]) {
        if (r.stream) {
          cudaStreamSynchronize(r.stream);
          cudaStreamDestroy(r.stream);
        }
        for (cudaEvent_t e : r.ev)
          if (e) cudaEventDestroy(e);
        if (r.arena_h) cudaFreeHost(r.arena_h);
        if (r.arena_d) cudaFree(r.arena_d);
      }
    pool.idle[d].clear();
  }
  if (cur >= 0) cudaSetDevice(cur);
  cudaGetLastError();
  return 0;
}
}  // namespace

int sctl_cuda_init(int n_gpus) {
  // multi_mu first, then g.mu (the order of sctl_cuda_eval's sharded path): a sharded Eval in flight keeps its
  // communicators until it has finished
  std::lock_guard<std::mutex> mlk(g.multi_mu);
  const int vis = visible_devices();
  if (vis < 1) return fail("no CUDA device available (libsctl_cuda has no CPU fallback)");
  if (n_gpus < 0 || n_gpus > vis) return fail("requested %d GPUs, %d visible", n_gpus, vis);
  if (n_gpus == 0) n_gpus = vis;
  if (n_gpus > MAX_DEV) n_gpus = MAX_DEV;
  {
    std::lock_guard<std::mutex> lk(g.mu);
    if (g.have_comms && g.n_gpus == n_gpus) return 0;
  }
  if (finalize_locked()) return 1;
  if (n_gpus > 1) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (!g.nccl.handle) {
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
      if (!h) return fail("cannot load NCCL: %s", dlerror());
      g.nccl.handle = h;
      g.nccl.CommInitAll = (decltype(g.nccl.CommInitAll))dlsym(h, "ncclCommInitAll");
      g.nccl.CommDestroy = (decltype(g.nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
      g.nccl.Broadcast = (decltype(g.nccl.Broadcast))dlsym(h, "ncclBroadcast");
      g.nccl.GroupStart = (decltype(g.nccl.GroupStart))dlsym(h, "ncclGroupStart");
      g.nccl.GroupEnd = (decltype(g.nccl.GroupEnd))dlsym(h, "ncclGroupEnd");
      g.nccl.GetErrorString = (decltype(g.nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
      if (!g.nccl.CommInitAll || !g.nccl.CommDestroy || !g.nccl.Broadcast || !g.nccl.GroupStart || !g.nccl.GroupEnd || !g.nccl.GetErrorString)
        return fail("NCCL library lacks a required symbol");
    }
    int devs[MAX_DEV];
    for (int i = 0; i < n_gpus; i++) devs[i] = i;
    ncclResult_t r = g.nccl.CommInitAll(g.comms, n_gpus, devs);
    if (r != ncclSuccess) return fail("ncclCommInitAll(%d) failed: %s", n_gpus, g.nccl.GetErrorString(r));
    g.have_comms = true;
  }
  int cur = 0;
  CU_TRY(cudaGetDevice(&cur));
  for (int i = 0; i < n_gpus; i++) {
    if (ensure_device_info(i)) return 1;
    if (n_gpus > 1 && !g.mstream[i]) {
      CU_TRY(cudaSetDevice(i));
      CU_TRY(cudaStreamCreateWithFlags(&g.mstream[i], cudaStreamNonBlocking));
      for (int k = 0; k < 4; k++) CU_TRY(cudaEventCreate(&g.mev[i][k]));
    }
  }
  CU_TRY(cudaSetDevice(cur));
  std::lock_guard<std::mutex> lk(g.mu);
  g.n_gpus = n_gpus;
  return 0;
}

int sctl_cuda_finalize(void) {
  std::lock_guard<std::mutex> mlk(g.multi_mu);
  return finalize_locked();
}

int sctl_cuda_kernel_info(int kernel_id, int* k0, int* k1, int* nd, int* flops, int* f_alg, double* scale, const char** name) {
  if (kernel_id < 0 || kernel_id >= NUM_KERNELS) return fail("unknown kernel id %d", kernel_id);
  const KernelInfo ki = kernel_info(kernel_id);
  if (k0) *k0 = ki.k0;
  if (k1) *k1 = ki.k1;
  if (nd) *nd = ki.nd;
  if (flops) *flops = ki.flops;
  if (f_alg) *f_alg = ki.f_alg;
  if (scale) *scale = scale_factor<double>(kernel_id);
  if (name) *name = ki.name;
  return 0;
}

int sctl_cuda_eval_device(int kernel_id, int real_bytes, int digits, int64_t Ns, int64_t Nt, const void* d_r_src, const void* d_n_src, const void* d_v_src,
                          const void* d_r_trg, void* d_v_trg, const void* ctx, int64_t ctx_bytes, void* stream) {
  if (check_common(kernel_id, real_bytes, Ns, Nt)) return 1;
  double ctx0;
  if (decode_ctx(kernel_id, real_bytes, ctx, ctx_bytes, &ctx0)) return 1;
  int dev;
  if (current_device_ready(&dev)) return 1;
  return eval_device_any(dev, kernel_id, real_bytes, digits, Ns, Nt, d_r_src, d_n_src, d_v_src, d_r_trg, d_v_trg, ctx0, (cudaStream_t)stream);
}

int sctl_cuda_kernel_matrix_device(int kernel_id, int real_bytes, int digits, int64_t Ns, int64_t Nt, const void* d_r_src, const void* d_n_src,
                                   const void* d_r_trg, void* d_M, int64_t ldM, const void* ctx, int64_t ctx_bytes, void* stream) {
  if (check_common(kernel_id, real_bytes, Ns, Nt)) return 1;
  double ctx0;
  if (decode_ctx(kernel_id, real_bytes, ctx, ctx_bytes, &ctx0)) return 1;
  int dev;
  if (current_device_ready(&dev)) return 1;
  return kmat_device_any(dev, kernel_id, real_bytes, digits, Ns, Nt, d_r_src, d_n_src, d_r_trg, d_M, ldM, ctx0, (cudaStream_t)stream);
}

int sctl_cuda_eval(int kernel_id, int real_bytes, int digits, int64_t Ns, int64_t Nt, const void* r_src, const void* n_src, const void* v_src,
                   const void* r_trg, void* v_trg, const void* ctx, int64_t ctx_bytes) {
  if (check_common(kernel_id, real_bytes, Ns, Nt)) return 1;
  double ctx0;
  if (decode_ctx(kernel_id, real_bytes, ctx, ctx_bytes, &ctx0)) return 1;
  const KernelInfo ki = kernel_info(kernel_id);
  int dev;
  if (current_device_ready(&dev)) return 1;
  tl.h2d_ms = tl.kernel_ms = tl.d2h_ms = 0;
  if (Ns == 0 || Nt == 0) return 0;
  if (!r_src || !v_src || !r_trg || !v_trg) return fail("null array pointer");
  if (ki.nd && !n_src) return fail("kernel %s needs source normals (n_src is NULL)", ki.name);
  int P;
  bool comms;
  {
    std::lock_guard<std::mutex> lk(g.mu);
    P = g.n_gpus;
    comms = g.have_comms;
  }
  if (P > 1 && comms && (double)Ns * (double)Nt >= MULTI_GPU_MIN_PAIRS && Nt >= P) {
    std::unique_lock<std::mutex> mlk(g.multi_mu);
    {  // init / finalize also take multi_mu: what was read above may be stale by now
      std::lock_guard<std::mutex> lk(g.mu);
      if (g.n_gpus != P || !g.have_comms) P = 0;
    }
    if (P == 0) {
      mlk.unlock();
      return sctl_cuda_eval(kernel_id, real_bytes, digits, Ns, Nt, r_src, n_src, v_src, r_trg, v_trg, ctx, ctx_bytes);
    }
    std::vector<ShardResult> res(P);
    std::vector<std::thread> th;
    ShardGate gate(P);
    for (int p = 0; p < P; p++)
      th.emplace_back(eval_host_shard, p, P, kernel_id, real_bytes, digits, Ns, Nt, (const char*)r_src, (const char*)n_src, (const char*)v_src,
                      (const char*)r_trg, (char*)v_trg, ctx0, &gate, &res[p]);
    for (auto& t : th) t.join();
    for (int p = 0; p < P; p++)  // report the shard that failed first-hand, not one that was merely cancelled
      if (res[p].rc && res[p].err.compare(0, 10, "cancelled:") != 0) return fail("GPU %d: %s", p, res[p].err.c_str());
    for (int p = 0; p < P; p++) {
      if (res[p].rc) return fail("GPU %d: %s", p, res[p].err.c_str());
      tl.h2d_ms = std::max(tl.h2d_ms, res[p].h2d);
      tl.kernel_ms = std::max(tl.kernel_ms, res[p].ker);
      tl.d2h_ms = std::max(tl.d2h_ms, res[p].d2h);
    }
    return 0;
  }
  {
    const size_t rb = (size_t)real_bytes;
    const size_t total = arena_align(rb * Nt * ki.k1) + arena_align(rb * Ns * 3) + arena_align(rb * Ns * ki.k0) + arena_align(rb * Ns * ki.nd) + arena_align(rb * Nt * 3);
    if (total <= SMALL_ARENA && g.opt.small_call_arena)
      return eval_host_small(dev, kernel_id, real_bytes, digits, Ns, Nt, (const char*)r_src, (const char*)n_src, (const char*)v_src, (const char*)r_trg, (char*)v_trg, ctx0);
  }
  return eval_host_one(dev, kernel_id, real_bytes, digits, Ns, 0, Nt, (const char*)r_src, (const char*)n_src, (const char*)v_src, (const char*)r_trg,
                       (char*)v_trg, ctx0, true);
}

int sctl_cuda_kernel_matrix(int kernel_id, int real_bytes, int digits, int64_t Ns, int64_t Nt, const void* r_src, const void* n_src, const void* r_trg,
                            void* M, int64_t ldM, const void* ctx, int64_t ctx_bytes) {
  if (check_common(kernel_id, real_bytes, Ns, Nt)) return 1;
  double ctx0;
  if (decode_ctx(kernel_id, real_bytes, ctx, ctx_bytes, &ctx0)) return 1;
  const KernelInfo ki = kernel_info(kernel_id);
  int dev;
  if (current_device_ready(&dev)) return 1;
  if (Ns == 0 || Nt == 0) return 0;
  if (!r_src || !r_trg || !M) return fail("null array pointer");
  if (ki.nd && !n_src) return fail("kernel %s needs source normals (n_src is NULL)", ki.name);
  const size_t rb = (size_t)real_bytes;
  const int64_t rows = Ns * ki.k0, cols = Nt * ki.k1;
  if (ldM < cols) return fail("ldM (%lld) smaller than Nt*K1 (%lld)", (long long)ldM, (long long)cols);
  cudaStream_t st = tl.res[dev].stream;
  {  // small-call arena (see eval_host_small): one pinned upload, one pinned download, no pool allocations
    const size_t b_rs = rb * Ns * 3, b_ns = ki.nd ? rb * Ns * ki.nd : 0, b_rt = rb * Nt * 3, b_M = rb * rows * cols;
    const size_t o_rs = 0, o_ns = o_rs + arena_align(b_rs), o_rt = o_ns + arena_align(b_ns), o_M = o_rt + arena_align(b_rt);
    if (o_M + arena_align(b_M) <= SMALL_ARENA && g.opt.small_call_arena) {
      if (ensure_arena(dev)) return 1;
      char *h = tl.res[dev].arena_h, *d = tl.res[dev].arena_d;
      memcpy(h + o_rs, r_src, b_rs);
      if (b_ns) memcpy(h + o_ns, n_src, b_ns);
      memcpy(h + o_rt, r_trg, b_rt);
      CU_TRY(cudaMemcpyAsync(d, h, o_M, cudaMemcpyHostToDevice, st));
      if (kmat_device_any(dev, kernel_id, real_bytes, digits, Ns, Nt, d + o_rs, b_ns ? d + o_ns : nullptr, d + o_rt, d + o_M, cols, ctx0, st)) return 1;
      CU_TRY(cudaMemcpyAsync(h + o_M, d + o_M, b_M, cudaMemcpyDeviceToHost, st));
      CU_TRY(cudaStreamSynchronize(st));
      if (ldM == cols) memcpy(M, h + o_M, b_M);
      else
        for (int64_t r = 0; r < rows; r++) memcpy((char*)M + rb * ldM * r, h + o_M + rb * cols * r, rb * cols);
      return 0;
    }
  }
  DevBuf d_rs, d_ns, d_rt, d_M;
  if (d_rs.alloc(rb * Ns * 3, st) || d_rt.alloc(rb * Nt * 3, st) || d_M.alloc(rb * rows * cols, st)) return 1;
  if (ki.nd && d_ns.alloc(rb * Ns * ki.nd, st)) return 1;
  CU_TRY(cudaMemcpyAsync(d_rs.p, r_src, rb * Ns * 3, cudaMemcpyHostToDevice, st));
  if (ki.nd) CU_TRY(cudaMemcpyAsync(d_ns.p, n_src, rb * Ns * ki.nd, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(d_rt.p, r_trg, rb * Nt * 3, cudaMemcpyHostToDevice, st));
  if (kmat_device_any(dev, kernel_id, real_bytes, digits, Ns, Nt, d_rs.p, d_ns.p, d_rt.p, d_M.p, cols, ctx0, st)) return 1;
  CU_TRY(cudaMemcpy2DAsync(M, rb * ldM, d_M.p, rb * cols, rb * cols, rows, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  return 0;
}

int sctl_cuda_rsqrt_device(int real_bytes, int digits, int64_t n, const void* d_x, void* d_y, void* stream) {
  if (real_bytes != 8 && real_bytes != 4) return fail("real_bytes must be 8 or 4");
  int dev;
  if (current_device_ready(&dev)) return 1;
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 4096);
  if (real_bytes == 8) {
    switch (rsq_mode_for_digits<double>(digits)) {
      case RSQ_SEED: rsqrt_test_kernel<double, RSQ_SEED><<<blocks, 256, 0, st>>>(n, (const double*)d_x, (double*)d_y); break;
      case RSQ_NEWTON: rsqrt_test_kernel<double, RSQ_NEWTON><<<blocks, 256, 0, st>>>(n, (const double*)d_x, (double*)d_y); break;
      default: rsqrt_test_kernel<double, RSQ_CUBIC><<<blocks, 256, 0, st>>>(n, (const double*)d_x, (double*)d_y); break;
    }
  } else {
    if (rsq_mode_for_digits<float>(digits) == RSQ_SEED) rsqrt_test_kernel<float, RSQ_SEED><<<blocks, 256, 0, st>>>(n, (const float*)d_x, (float*)d_y);
    else rsqrt_test_kernel<float, RSQ_NEWTON><<<blocks, 256, 0, st>>>(n, (const float*)d_x, (float*)d_y);
  }
  CU_TRY(cudaGetLastError());
  return 0;
}

int sctl_cuda_idle_thread_resources(int device) {
  if (device < 0 || device >= MAX_DEV) return -1;
  ResPool& pool = res_pool();
  std::lock_guard<std::mutex> lk(pool.mu);
  return (int)pool.idle[device].size();
}

int sctl_cuda_last_timing(double* h2d_ms, double* kernel_ms, double* d2h_ms) {
  if (h2d_ms) *h2d_ms = tl.h2d_ms;
  if (kernel_ms) *kernel_ms = tl.kernel_ms;
  if (d2h_ms) *d2h_ms = tl.d2h_ms;
  return 0;
}

int sctl_cuda_last_plan(int* grid_x, int* grid_y, int* threads, int* smem_bytes, int* ctas_per_sm, int* n_launches, double* wave_eff) {
  if (grid_x) *grid_x = tl.grid_x;
  if (grid_y) *grid_y = tl.grid_y;
  if (threads) *threads = tl.threads;
  if (smem_bytes) *smem_bytes = tl.smem;
  if (ctas_per_sm) *ctas_per_sm = tl.ctas_per_sm;
  if (n_launches) *n_launches = tl.n_launches;
  if (wave_eff) *wave_eff = tl.wave_eff;
  return 0;
}

int sctl_cuda_set_option(const char* name, int64_t value) {
  if (!name) return fail("null option name");
  std::lock_guard<std::mutex> lk(g.mu);
  if (!strcmp(name, "const_min_targets")) g.opt.const_min_targets = value;
  else if (!strcmp(name, "const_min_launch_flop")) g.opt.const_min_launch_flop = value;
  else if (!strcmp(name, "few_max_targets")) g.opt.few_max_targets = value;
  else if (!strcmp(name, "ring_min_tiles")) g.opt.ring_min_tiles = std::max<int64_t>(1, value);
  else if (!strcmp(name, "ring_max_split")) g.opt.ring_max_split = std::max<int64_t>(1, std::min<int64_t>(value, 1024));
  else if (!strcmp(name, "wide_min_targets")) g.opt.wide_min_targets = value;
  else if (!strcmp(name, "const_fp32")) g.opt.const_fp32 = value;
  else if (!strcmp(name, "const_banks")) g.opt.const_banks = value;
  else if (!strcmp(name, "const_banks_fp32")) g.opt.const_banks_fp32 = value;
  else if (!strcmp(name, "const_ctas_per_sm")) g.opt.const_ctas_per_sm = value;
  else if (!strcmp(name, "const_ctas_per_sm_overlap")) g.opt.const_ctas_per_sm_overlap = value;
  else if (!strcmp(name, "probe_min_pairs")) g.opt.probe_min_pairs = value;
  else if (!strcmp(name, "probe_per_block")) g.opt.probe_per_block = value;
  else if (!strcmp(name, "small_call_arena")) g.opt.small_call_arena = value;
  else return fail("unknown option '%s'", name);
  return 0;
}

int sctl_cuda_reset_options() {
  std::lock_guard<std::mutex> lk(g.mu);
  g.opt = Options();
  return 0;
}

int sctl_cuda_plan_eval(int64_t Ns, int64_t Nt, int tpb, int slots, int min_chunk, int max_split, int64_t* n_tblocks, int* n_split, int64_t* chunk,
                        double* wave_eff) {
  if (tpb < 1 || slots < 1 || max_split < 1) return fail("bad planner arguments");
  const EvalPlan p = plan_eval(Ns, Nt, tpb, slots, min_chunk, max_split);
  if (n_tblocks) *n_tblocks = p.n_tblocks;
  if (n_split) *n_split = p.n_split;
  if (chunk) *chunk = p.chunk;
  if (wave_eff) *wave_eff = p.wave_eff;
  return 0;
}

int sctl_cuda_const_geometry(int kernel_id, int real_bytes, int64_t Ns, int64_t Nt, int banks_wanted, int* banks, int* chunk, int* row_words, int* bank_stride,
                             int* rec_bytes, int64_t* seg_bytes, int64_t* n_chunks) {
  if (check_common(kernel_id, real_bytes, Ns, Nt)) return 1;
  if (banks_wanted < 1 || banks_wanted > CONST_MAX_BANKS) return fail("banks_wanted must be in [1, %d]", CONST_MAX_BANKS);
  ConstGeom cg = {};
#define SCTL_CUDA_CASE(K)                                                                                         \
  case K:                                                                                                         \
    cg = real_bytes == 8 ? const_geometry<double, K>(Ns, const_target_blocks<double>(Nt), banks_wanted, true)     \
                         : const_geometry<float, K>(Ns, const_target_blocks<float>(Nt), banks_wanted, true);      \
    break;
  switch (kernel_id) {
    SCTL_CUDA_CASE(LAPLACE3D_FXU)
    SCTL_CUDA_CASE(LAPLACE3D_DXU)
    SCTL_CUDA_CASE(LAPLACE3D_FXDU)
    SCTL_CUDA_CASE(STOKES3D_FXU)
    SCTL_CUDA_CASE(STOKES3D_DXU)
    SCTL_CUDA_CASE(STOKES3D_FXT)
    SCTL_CUDA_CASE(STOKES3D_FSXU)
    SCTL_CUDA_CASE(STOKES3D_FXUP)
    SCTL_CUDA_CASE(LAPLACE3D_FXD2U)
    SCTL_CUDA_CASE(LAPLACE3D_DXDU)
    SCTL_CUDA_CASE(BIOTSAVART3D_FXU)
    SCTL_CUDA_CASE(BIOTSAVART3D_FXDU)
    SCTL_CUDA_CASE(HELMHOLTZ3D_FXU)
    SCTL_CUDA_CASE(HELMHOLTZ3D_DXU)
    SCTL_CUDA_CASE(HELMHOLTZ3D_FXDU)
    SCTL_CUDA_CASE(HELMHOLTZDIFF3D_FXDU)
  }
#undef SCTL_CUDA_CASE
  if (banks) *banks = cg.banks;
  if (chunk) *chunk = cg.chunk;
  if (row_words) *row_words = cg.row_words;
  if (bank_stride) *bank_stride = cg.bank_stride;
  if (rec_bytes) *rec_bytes = cg.rec_bytes;
  if (seg_bytes) *seg_bytes = cg.seg_bytes;
  if (n_chunks) *n_chunks = cg.n_chunks;
  return 0;
}

int sctl_cuda_fma_peak(int real_bytes, double ms, double* tflops) {
  if (!tflops) return fail("null output pointer");
  if (visible_devices() < 1) return fail("no CUDA device available");
  if (real_bytes == 8) return fma_peak_t<double>(ms, tflops);
  if (real_bytes == 4) return fma_peak_t<float>(ms, tflops);
  return fail("real_bytes must be 8 or 4");
}

}  // extern "C"
