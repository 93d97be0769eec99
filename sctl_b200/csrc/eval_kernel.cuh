// All-pairs evaluation  v_trg[t] += scale * sum_s K(x_t - x_s, n_s) f_s  on one B200.
//
// Replaces the body of GenericKernel<uKernel>::Eval (reference: include/sctl/generic-kernel.txx:102-211,
// hot loop :176-188).  Design:
//   * a target is a thread-resident register set (coordinates + K1 accumulators), T targets per thread;
//   * sources are first packed into 16-byte-multiple records (pack_sources_kernel), then streamed through
//     shared memory by a dedicated producer warp with cp.async.bulk (TMA 1-D bulk copy, SASS UBLKCP) into
//     a STAGES-deep ring guarded by full/empty mbarriers; consumers read records with broadcast LDS.128;
//   * the (target-block x source-chunk) grid is sized by plan_eval() so the last wave is full; when the
//     source range is split the partial sums go to a scratch array and are combined in chunk order by
//     reduce_partials_kernel (deterministic, no atomics);
//   * R > 1 makes the R warps of a group share one set of targets and split every tile's sources between them
//     (few-target regime, reference :135-156), combined through shared memory in warp order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "pair_math.cuh"

namespace sctl_cuda {

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (global -> shared::cta)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// cp.async.bulk: `bytes` must be a multiple of 16, both addresses 16-byte aligned; completion is
// signalled on `bar` as transaction bytes.
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Source packing:  AoS r_src | n_src | v_src  ->  one record per source
//   FxU, FxdU, Fxd2U : [x y z f]          Stokes FxU/FxT/FxUP, BiotSavart : [x y z fx fy fz]
//   Laplace DxU/DxdU: [x y z f*nx f*ny f*nz]   Stokes DxU : [x y z f n]   Stokes FSxU : [x y z f0..f3]   (+ zero padding to 16 B)
// O(Ns) work in front of an O(Ns*Nt) kernel; it buys 16-byte alignment for arbitrary caller pointers,
// one bulk copy per tile, and one FP64 instruction per pair for Laplace DxU.
// ------------------------------------------------------------------------------------------------
template <class Real, int KID>
__global__ void pack_sources_kernel(int64_t Ns, const Real* __restrict__ r_src, const Real* __restrict__ n_src,
                                    const Real* __restrict__ v_src, Real* __restrict__ pack, int64_t chunk_len, int64_t seg_reals) {
  constexpr int REC = RecTraits<Real, KID>::REC;
  constexpr int K0 = KerTraits<KID>::K0;
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < Ns; s += (int64_t)gridDim.x * blockDim.x) {
    Real rec[REC];
#pragma unroll
    for (int k = 0; k < REC; k++) rec[k] = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) rec[k] = r_src[s * 3 + k];
    if (KID == LAPLACE3D_DXU || KID == LAPLACE3D_DXDU) {  // double-layer Laplace kernels are linear in f*n
      const Real f = v_src[s];
#pragma unroll
      for (int k = 0; k < 3; k++) rec[3 + k] = f * n_src[s * 3 + k];
    } else {
#pragma unroll
      for (int k = 0; k < K0; k++) rec[3 + k] = v_src[s * K0 + k];
      if (KID == STOKES3D_DXU || KID == HELMHOLTZ3D_DXU) {  // normals follow the densities
#pragma unroll
        for (int k = 0; k < 3; k++) rec[3 + K0 + k] = n_src[s * 3 + k];
      }
    }
    // contiguous records (chunk_len >= Ns), or chunk-strided segments for the constant-bank path: records of chunk c at
    // pack + c * seg_reals, followed by that chunk's coincidence bits (coincidence.cuh), so one copy moves both
    Real* dst = pack + (s / chunk_len) * seg_reals + (s % chunk_len) * REC;
#pragma unroll
    for (int k = 0; k < REC; k++) dst[k] = rec[k];
  }
}

// ------------------------------------------------------------------------------------------------
// The pair kernel
// ------------------------------------------------------------------------------------------------
template <class Real> struct EvalArgs {
  const Real* pack;   // [Ns][REC] packed sources, 16-byte aligned
  const Real* r_trg;  // [Nt][3] AoS
  Real* v_trg;        // [Nt][K1] AoS, accumulated into when n_split == 1
  Real* partial;      // [n_split][Nt][K1] unscaled partial sums when n_split > 1
  int64_t Ns, Nt;
  int64_t chunk;      // sources per grid.y slice
  int n_split;
  Real scale;
  const int* must_mask;  // device flag from coincidence_probe_kernel (nullptr: always mask r2 == 0 pairs)
  Real ctx0;             // kernel context by value (Helmholtz wavenumber); 0 otherwise
};

template <class Real, int KID, int T, int NW, int R, int RSQ, int TILE, int STAGES> struct EvalConfig {
  static constexpr int REC = RecTraits<Real, KID>::REC;
  static constexpr int K1 = KerTraits<KID>::K1;
  static constexpr int GROUPS = NW / R;                 // warp groups with distinct targets
  static constexpr int TPB = GROUPS * 32 * T;           // targets per CTA
  static constexpr int THREADS = (NW + 1) * 32;         // + producer warp
  static constexpr int STAGE_REALS = TILE * REC;
  static constexpr size_t STAGE_BYTES = (size_t)STAGE_REALS * sizeof(Real);
  static constexpr size_t SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * sizeof(uint64_t);
  static_assert(NW % R == 0, "R must divide NW");
  static_assert(STAGE_BYTES % 16 == 0, "stage must be a whole number of 16-byte words");
  static_assert(R == 1 || (size_t)NW * T * K1 * 32 * sizeof(Real) <= STAGES * STAGE_BYTES, "reduction scratch must fit in the ring");
};

// One tile of packed records against this thread's T register-resident targets.  V is the type a pair is computed in:
// Real, or f32x2 (two targets per value; the record's scalars are broadcast operands of the packed FP32 instructions,
// pair_math.cuh).
template <class V, class Real, int KID, int T, int R, int RSQ, int U, bool MASK>
__device__ __forceinline__ void consume_tile(const Real* tile, int n, int rep, V (&acc)[T][KerTraits<KID>::K1], const V (&xt)[T][3], V mu) {
  constexpr int REC = RecTraits<Real, KID>::REC;
  constexpr int W = 16 / (int)sizeof(Real);
  using Vec16 = typename std::conditional<sizeof(Real) == 8, double2, float4>::type;
#pragma unroll U
  for (int s = rep; s < n; s += R) {
    Real rec[REC];
    const Vec16* p = reinterpret_cast<const Vec16*>(tile + s * REC);
#pragma unroll
    for (int q = 0; q < REC / W; q++) {
      const Vec16 v = p[q];
#pragma unroll
      for (int c = 0; c < W; c++) rec[q * W + c] = ((const Real*)&v)[c];
    }
#pragma unroll
    for (int j = 0; j < T; j++) pair_update<KID, RSQ, MASK>(acc[j], xt[j], rec, mu);
  }
}

template <class Real, int KID, int T, int NW, int R, int RSQ, int TILE, int STAGES, int MINB, int U = 2>
__global__ void __launch_bounds__((NW + 1) * 32, MINB) eval_pairs_kernel(const EvalArgs<Real> a) {
  using Cfg = EvalConfig<Real, KID, T, NW, R, RSQ, TILE, STAGES>;
  constexpr int REC = Cfg::REC;
  constexpr int K1 = Cfg::K1;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  Real* ring = reinterpret_cast<Real*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int64_t s_begin = (int64_t)blockIdx.y * a.chunk;
  const int64_t s_end = (s_begin + a.chunk < a.Ns) ? s_begin + a.chunk : a.Ns;
  const int64_t n_src = s_end > s_begin ? s_end - s_begin : 0;
  const int n_tiles = (int)((n_src + TILE - 1) / TILE);

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NW);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == NW) {
    // ===== producer warp: one lane streams the chunk's records into the ring =====
    if (lane == 0) {
      const Real* src = a.pack + s_begin * REC;
      for (int i = 0; i < n_tiles; i++) {
        const int st = i % STAGES;
        if (i >= STAGES) mbar_wait(&empty[st], ((i / STAGES) - 1) & 1);
        const int64_t rem = n_src - (int64_t)i * TILE;
        const uint32_t bytes = (uint32_t)((rem < TILE ? rem : TILE) * REC * sizeof(Real));
        mbar_arrive_expect_tx(&full[st], bytes);
        bulk_copy_g2s(ring + st * Cfg::STAGE_REALS, src + (int64_t)i * TILE * REC, bytes, &full[st]);
      }
    }
    return;
  }

  // ===== consumer warps =====
  const int group = warp / R;
  const int rep = warp % R;
  const int64_t t_base = (int64_t)blockIdx.x * Cfg::TPB + (int64_t)group * (32 * T) + lane;

  Real xt[T][3];
  Real acc[T][K1];
#pragma unroll
  for (int j = 0; j < T; j++) {
    const int64_t t = t_base + j * 32;
#pragma unroll
    for (int k = 0; k < 3; k++) xt[j][k] = (t < a.Nt) ? a.r_trg[t * 3 + k] : (Real)0;
#pragma unroll
    for (int k = 0; k < K1; k++) acc[j][k] = 0;
  }

  const bool unmasked = a.must_mask != nullptr && a.must_mask[0] == 0;  // same value in every thread of the grid
  // FP32: targets 2p and 2p+1 of this thread share one packed value (same operations per target as the scalar loop,
  // hence the same bits)
  constexpr bool PACKED = SCTL_CUDA_F32X2 && sizeof(Real) == 4 && T % 2 == 0;
  constexpr int TP = PACKED ? T / 2 : 1;
  f32x2 xt2[TP][3], acc2[TP][K1];
  if constexpr (PACKED) {
#pragma unroll
    for (int p = 0; p < TP; p++) {
#pragma unroll
      for (int k = 0; k < 3; k++) xt2[p][k] = f32x2(xt[2 * p][k], xt[2 * p + 1][k]);
#pragma unroll
      for (int k = 0; k < K1; k++) acc2[p][k] = f32x2(0.0f);
    }
  }
  for (int i = 0; i < n_tiles; i++) {
    const int st = i % STAGES;
    mbar_wait(&full[st], (i / STAGES) & 1);
    const Real* tile = ring + st * Cfg::STAGE_REALS;
    const int64_t rem = n_src - (int64_t)i * TILE;
    const int n = (int)(rem < TILE ? rem : TILE);
    if constexpr (PACKED) {
      if (unmasked) consume_tile<f32x2, Real, KID, TP, R, RSQ, U, false>(tile, n, rep, acc2, xt2, f32x2((float)a.ctx0));
      else consume_tile<f32x2, Real, KID, TP, R, RSQ, U, true>(tile, n, rep, acc2, xt2, f32x2((float)a.ctx0));
    } else {
      if (unmasked) consume_tile<Real, Real, KID, T, R, RSQ, U, false>(tile, n, rep, acc, xt, a.ctx0);
      else consume_tile<Real, Real, KID, T, R, RSQ, U, true>(tile, n, rep, acc, xt, a.ctx0);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }
  if constexpr (PACKED) {
#pragma unroll
    for (int p = 0; p < TP; p++)
#pragma unroll
      for (int k = 0; k < K1; k++) {
        acc[2 * p][k] = acc2[p][k].v.x;
        acc[2 * p + 1][k] = acc2[p][k].v.y;
      }
  }

  if (R > 1) {
    // combine the R replicas of each group in warp order through the (now idle) ring
    named_bar_sync(1, NW * 32);
    Real* red = ring;
#pragma unroll
    for (int j = 0; j < T; j++)
#pragma unroll
      for (int k = 0; k < K1; k++) red[((warp * T + j) * K1 + k) * 32 + lane] = acc[j][k];
    named_bar_sync(1, NW * 32);
    if (rep != 0) return;
#pragma unroll
    for (int j = 0; j < T; j++)
#pragma unroll
      for (int k = 0; k < K1; k++) {
        Real sum = acc[j][k];
        for (int r = 1; r < R; r++) sum += red[(((warp + r) * T + j) * K1 + k) * 32 + lane];
        acc[j][k] = sum;
      }
  }

#pragma unroll
  for (int j = 0; j < T; j++) {
    const int64_t t = t_base + j * 32;
    if (t < a.Nt) {
      if (a.n_split == 1) {
#pragma unroll
        for (int k = 0; k < K1; k++) a.v_trg[t * K1 + k] = fma(acc[j][k], a.scale, a.v_trg[t * K1 + k]);
      } else {
        Real* out = a.partial + ((int64_t)blockIdx.y * a.Nt + t) * K1;
#pragma unroll
        for (int k = 0; k < K1; k++) out[k] = acc[j][k];
      }
    }
  }
}

// v_trg[i] += scale * sum_c partial[c][i], chunks in index order (deterministic).
template <class Real>
__global__ void reduce_partials_kernel(int64_t n, int n_split, const Real* __restrict__ partial, Real scale, Real* __restrict__ v_trg) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    Real sum = 0;
    for (int c = 0; c < n_split; c++) sum += partial[(int64_t)c * n + i];
    v_trg[i] = fma(sum, scale, v_trg[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Grid planning: pick the number of source chunks so that (target blocks x chunks) fills whole waves.
// ------------------------------------------------------------------------------------------------
struct EvalPlan {
  int64_t n_tblocks;
  int n_split;
  int64_t chunk;
  double wave_eff;
};

inline EvalPlan plan_eval(int64_t Ns, int64_t Nt, int tpb, int slots /* SMs x resident CTAs */, int min_chunk, int max_split) {
  EvalPlan best{};
  best.n_tblocks = (Nt + tpb - 1) / tpb;
  best.n_split = 1;
  best.chunk = Ns;
  best.wave_eff = 0;
  if (Ns <= 0 || Nt <= 0) return best;
  int64_t cap = Ns / (min_chunk > 0 ? min_chunk : 1);
  if (cap < 1) cap = 1;
  if (cap > max_split) cap = max_split;
  for (int64_t sp = 1; sp <= cap; sp++) {
    const int64_t chunk = (Ns + sp - 1) / sp;
    const int64_t n_chunks = (Ns + chunk - 1) / chunk;
    const int64_t ctas = best.n_tblocks * n_chunks;
    const int64_t waves = (ctas + slots - 1) / slots;
    // ragged last chunk / partially filled last target block count as lost work too
    const double eff = (double)Ns * (double)Nt / ((double)waves * slots * (double)chunk * tpb);
    if (eff > best.wave_eff + 1e-9) {
      best.n_split = (int)n_chunks;
      best.chunk = chunk;
      best.wave_eff = eff;
    }
    if (best.wave_eff >= 0.97) break;
  }
  return best;
}

}  // namespace sctl_cuda
