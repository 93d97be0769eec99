// Constant-bank pair kernel: the main Eval path for many targets.
//
// Sources are read through the constant cache into UNIFORM registers (SASS LDCU.64 URx, c[0x3][URy+imm]) so the
// FP64 instructions that consume a source value take it as a uniform operand (DADD R, R, -UR / DFMA R, R, UR, R)
// and read at most two vector registers.  Why it matters: on B200 a DFMA/DMUL/DADD with three distinct
// vector-register sources issues every 3 cycles instead of 2 (register-file read bandwidth; tools/probe.cu measures
// 42.4 vs 63.6 lanes/clk/SM).  With sources in shared memory 9 of the 22 FP64 instructions of a Stokes3D_FxU pair
// are of that kind (modelled FP64-pipe ceiling 84.6 %); with sources in uniform registers only the three
// accumulate FMAs are (ceiling 93.6 %).
//
// The constant bank cannot be written from the device, so the host streams the packed sources through it in chunks
// (cudaMemcpyToSymbolAsync, device to device; ConstGeom below) and launches this kernel once per chunk; every
// launch adds scale * (chunk's partial sum) to its accumulation buffer.  Targets stay register-resident for the whole chunk, T per
// thread.  When there are too few target blocks to balance 148 SMs, the chunk is split CL ways along grid.y and the
// CL CTAs that own the same targets form a thread-block cluster: partial sums are combined through distributed
// shared memory in rank order (deterministic, no atomics, no scratch traffic).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pair_math.cuh"

// Tuning switches (tools/ab_build.sh builds variants to measure the difference):
//   SCTL_CUDA_CONST_HOIST  -1: per kernel (ConstHoist below); 0 / 1: the refinement's loop-invariant operands are
//                          literals / kernel arguments for every FP64 kernel
//   SCTL_CUDA_F32X2        (pair_math.cuh) 1: FP32 pair loops on packed values; 0: scalar
#ifndef SCTL_CUDA_CONST_HOIST
#define SCTL_CUDA_CONST_HOIST -1
#endif
//   SCTL_CUDA_CONST_SWP    1: software-pipelined FP64 loop (const_pair_loop_swp); 0 (default): the plain loop
#ifndef SCTL_CUDA_CONST_SWP
#define SCTL_CUDA_CONST_SWP 0
#endif
//   SCTL_CUDA_CONST_PRAGMA_LOOP  1 (default): kernels without hoisted operands use `#pragma unroll U` over the sources;
//                          0: the hand-blocked loop for them too.  Same instructions, but ptxas schedules the two forms
//                          differently: Stokes3D_FxU FP64 807.7 vs 797.1 Gpair/s (profiles/r2b_ab_rate.md, third session)
#ifndef SCTL_CUDA_CONST_PRAGMA_LOOP
#define SCTL_CUDA_CONST_PRAGMA_LOOP 1
#endif
//   SCTL_CUDA_CONST32_BLOCKED_LOOP  the same choice for the packed FP32 loops (1: hand-blocked)
#ifndef SCTL_CUDA_CONST32_BLOCKED_LOOP
#define SCTL_CUDA_CONST32_BLOCKED_LOOP 1
#endif

namespace sctl_cuda {

constexpr int CONST_PACK_BYTES = 64 * 1024 - 1024;  // leave room for the compiler's own bank-3 constants
__constant__ double c_src_pack[CONST_PACK_BYTES / 8];
// [0]: the coincidence probe's verdict for the call in flight (0 = no source/target pair can have r2 < 2^-1022, so the
// unmasked loop is exact); [1]: 1 = the per-CTA bits that refine a verdict of 1 are not usable.  Copied device-to-device
// from the probe's flags, so the choice needs no host round trip, and read from the constant bank, so the branch on it
// is uniform and the loops keep their uniform-register operands.
__constant__ int c_eval_must_mask[4];

template <class Real> struct ConstEvalArgs {
  const Real* r_trg;  // [Nt][3]
  Real* v_trg;        // [Nt][K1], accumulated into
  int64_t Nt;
  int n_rec;          // records in the constant bank for this launch
  int sub_len;        // records per cluster rank (grid.y slice)
  int full_ranks;     // n_rec / sub_len: ranks with a full slice
  int tail_len;       // n_rec % sub_len: records of rank `full_ranks`
  int bank_off;       // first Real of this launch's bank inside c_src_pack (multi-bank streaming, see below)
  int bits_off;       // first 32-bit word of this launch's coincidence bits inside c_src_pack (one bit per CTA), -1: none
  int allow_unmasked; // this call ran the coincidence probe: c_eval_must_mask is meaningful
  int zero;           // 0, as a value ptxas cannot fold: the seed's low word (see the kernel)
  Real scale;
  Real ctx0;          // kernel context by value (Helmholtz wavenumber)
  Real rsq_k;         // rsq_poly_k<RSQ>(): the refinement's register-resident coefficient
};

// Whether the FP64 loop keeps the rsqrt refinement's loop-invariant operands in registers (see the kernel).  It removes
// 10 IMAD.MOV per 8 pairs from the loop, but those issue in the FP64 pipe's shadow and what decides is how ptxas then
// schedules the loop (the hoisted operands also need the hand-blocked loop form, const_pair_loop).  Chosen per kernel from
// measurements on B200 (profiles/r2b_ab_rate.md, 2^19 x 2^19, digits -1 / 11, against the plain loop): Laplace3D_FxU
// +1.9 % / +0.6 %, Laplace3D_FxdU +1.6 % / +1.8 %, Stokes3D_FSxU +0.2 % / +1.4 %, Laplace3D_Fxd2U +1.0 % / +1.6 %,
// BiotSavart3D_FxdU +2.2 % / +0.2 %, Helmholtz3D_DxU +0.8 % / +0.3 %, HelmholtzDiff3D_FxdU 0 / +1.2 %; Stokes3D_FxU
// -1.3 % / -0.1 %, Stokes3D_DxU, Laplace3D_DxU, Stokes3D_FxUP, Laplace3D_DxdU, BiotSavart3D_FxU -0.3 ... -2 %, Stokes3D_FxT,
// Helmholtz3D_FxU and Helmholtz3D_FxdU one way at digits -1 and the other at 11.
template <int KID> constexpr bool const_hoist_measured_faster() {
  return KID == LAPLACE3D_FXU || KID == LAPLACE3D_FXDU || KID == STOKES3D_FSXU || KID == LAPLACE3D_FXD2U || KID == BIOTSAVART3D_FXDU || KID == HELMHOLTZ3D_DXU ||
         KID == HELMHOLTZDIFF3D_FXDU;
}
template <class Real, int KID> struct ConstHoist {
  static constexpr bool value = SCTL_CUDA_CONST_HOIST < 0 ? (sizeof(Real) == 8 && const_hoist_measured_faster<KID>()) : (SCTL_CUDA_CONST_HOIST != 0);
};

template <class Real, int KID> constexpr int const_chunk_records() { return CONST_PACK_BYTES / (RecTraits<Real, KID>::REC * (int)sizeof(Real)); }

// Multi-bank streaming: the buffer is cut into `banks` equal slices and chunk i goes through slice i % banks on its
// own stream, so the launch for chunk i+1 is already queued (its slice filled) while chunk i runs and its CTAs take
// over the SMs as chunk i's last wave drains -- no idle tail and no wave quantisation per launch.  Launches that
// share a slice are ordered by their stream; launches of different slices accumulate into different buffers.
constexpr int CONST_MAX_BANKS = 4;

// Geometry of one call: how the constant buffer is cut and how the packed sources are laid out in HBM.  Every chunk is
// a SEGMENT of seg_bytes: `chunk` records followed (when the call carries coincidence bits) by `row_words` 32-bit words
// with one bit per target block; a launch copies one segment into its slice with one cudaMemcpyToSymbolAsync.
struct ConstGeom {
  int banks;            // slices in flight
  int chunk;            // records per launch
  int row_words;        // 32-bit words of coincidence bits per chunk (0: none), a multiple of 4
  int bank_stride;      // bytes between slices in c_src_pack
  int rec_bytes;
  int64_t seg_bytes;
  int64_t n_chunks;
};
template <class Real, int KID> inline ConstGeom const_geometry(int64_t Ns, int64_t n_tblocks, int banks_wanted, bool want_bits) {
  ConstGeom g;
  g.rec_bytes = RecTraits<Real, KID>::REC * (int)sizeof(Real);
  for (int pass = 0;; pass++) {
    g.banks = pass ? 1 : banks_wanted;
    g.bank_stride = CONST_PACK_BYTES / g.banks / 16 * 16;
    const int64_t words = (n_tblocks + 31) / 32;
    g.row_words = (want_bits && words * 4 <= g.bank_stride / 8) ? (int)((words + 3) / 4 * 4) : 0;  // bits take at most 1/8 of a slice
    g.chunk = (g.bank_stride - g.row_words * 4) / g.rec_bytes;
    g.n_chunks = (Ns + g.chunk - 1) / g.chunk;
    if (g.banks == 1 || g.n_chunks >= 2 * g.banks) break;  // too few chunks to pipeline: one slice
  }
  if (g.n_chunks > 0xFFFFFFF0LL) g.row_words = 0;  // chunk numbers are kept in 32 bits by the probe
  g.seg_bytes = (int64_t)g.chunk * g.rec_bytes + g.row_words * 4;
  return g;
}

// U sources x T targets per iteration.  BLOCKED: unrolled by hand (not #pragma unroll U) so that every unrolled pair has
// its own `lo` operand, the remainder (n % U sources) one source at a time -- the form the hoisted operands need.
// Otherwise `#pragma unroll U` over the plain source loop: the same instructions, which ptxas schedules better (FP64
// Stokes3D_FxU 807.7 vs 797.1 Gpair/s, profiles/r2b_ab_rate.md).  `src` and `n` are warp-uniform: the records are read
// with LDCU into uniform registers.  V is the type a pair is computed in: Real, or f32x2 (two targets per value, packed
// FP32 instructions with the record's scalars as broadcast uniform operands; pair_math.cuh).
template <class V, class Real, int KID, int T, int RSQ, int U, bool MASK, bool BLOCKED>
__device__ __forceinline__ void const_pair_loop(V (&acc)[T][KerTraits<KID>::K1], const V (&xt)[T][3], const Real* src, int n, V ctx0,
                                                const int (&lo)[U][T], V k) {
  constexpr int REC = RecTraits<Real, KID>::REC;
  if constexpr (BLOCKED) {
    int s = 0;
    for (; s + U <= n; s += U) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        const Real* rec = src + (s + u) * REC;  // uniform address: LDCU into uniform registers
#pragma unroll
        for (int j = 0; j < T; j++) pair_update<KID, RSQ, MASK>(acc[j], xt[j], rec, ctx0, lo[u][j], k);
      }
    }
    for (; s < n; s++) {
      const Real* rec = src + s * REC;
#pragma unroll
      for (int j = 0; j < T; j++) pair_update<KID, RSQ, MASK>(acc[j], xt[j], rec, ctx0, lo[0][j], k);
    }
  } else {
#pragma unroll U
    for (int s = 0; s < n; s++) {
      const Real* rec = src + s * REC;
#pragma unroll
      for (int j = 0; j < T; j++) pair_update<KID, RSQ, MASK>(acc[j], xt[j], rec, ctx0, lo[0][j], k);
    }
  }
}

// Software-pipelined form of the FP64 loop (tuning build, SCTL_CUDA_CONST_SWP=1): the front half of group g+1 (loads,
// dx, r2, seed) is issued between the back halves of group g, so no pair waits for its own MUFU / LDCU latency.
template <int KID, int T, int RSQ, int U, bool MASK>
__device__ __forceinline__ void const_pair_loop_swp(double (&acc)[T][KerTraits<KID>::K1], const double (&xt)[T][3], const double* src, int n) {
  constexpr int REC = RecTraits<double, KID>::REC;
  double dx[U][T][3], r2[U][T], y0[U][T];
  int s = 0;
  if (n >= U) {
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < T; j++) pair_front<RSQ, MASK>(dx[u][j], r2[u][j], y0[u][j], xt[j], src + u * REC);
    for (s = U; s + U <= n; s += U) {
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int j = 0; j < T; j++) {
          pair_back<KID, RSQ>(acc[j], dx[u][j], r2[u][j], y0[u][j], src + (s - U + u) * REC);
          pair_front<RSQ, MASK>(dx[u][j], r2[u][j], y0[u][j], xt[j], src + (s + u) * REC);
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < T; j++) pair_back<KID, RSQ>(acc[j], dx[u][j], r2[u][j], y0[u][j], src + (s - U + u) * REC);
  }
  for (; s < n; s++) {
    const double* rec = src + s * REC;
#pragma unroll
    for (int j = 0; j < T; j++) pair_update<KID, RSQ, MASK>(acc[j], xt[j], rec);
  }
}

// FP32 pair loops run on packed values (two targets each)
template <class Real, int KID, int T> __host__ __device__ constexpr bool const_packed() { return SCTL_CUDA_F32X2 && sizeof(Real) == 4 && T % 2 == 0; }

// NT is the LARGEST block the kernel may be launched with; the host picks blockDim.x = 32 * (warps per CTA) <= NT so
// that every SM gets the same number of warps (see launch_eval_const in sctl_cuda.cu).
template <class Real, int KID, int T, int NT, int RSQ, int U, int MINB>
__global__ void __launch_bounds__(NT, MINB) eval_pairs_const_kernel(const ConstEvalArgs<Real> a) {
  constexpr int REC = RecTraits<Real, KID>::REC;
  constexpr int K1 = KerTraits<KID>::K1;
  __shared__ Real red[T * K1 * NT];  // cluster reduction buffer
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t t_base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * (32 * T) + lane;
  Real xt[T][3], acc[T][K1];
#pragma unroll
  for (int j = 0; j < T; j++) {
    const int64_t t = t_base + j * 32;
#pragma unroll
    for (int k = 0; k < 3; k++) xt[j][k] = (t < a.Nt) ? a.r_trg[t * 3 + k] : (Real)0;
#pragma unroll
    for (int k = 0; k < K1; k++) acc[j][k] = 0;
  }
  // Loop bounds are kept in a form ptxas recognises as warp-uniform, otherwise it falls back to vector LDC.
  const Real* src = reinterpret_cast<const Real*>(c_src_pack) + a.bank_off + (int)blockIdx.y * a.sub_len * REC;
  // ranks [0, full_ranks) take sub_len records, rank full_ranks the remainder, later ranks nothing (a short last chunk
  // can have fewer records than the cluster has ranks); all three come from the host so the bound stays uniform
  const int n = ((int)blockIdx.y < a.full_ranks) ? a.sub_len : (((int)blockIdx.y == a.full_ranks) ? a.tail_len : 0);
  // Masked or unmasked loop (bit-identical results wherever no pair has r2 < 2^-1022): unmasked when the probe found no
  // coincident source/target pair at all, or -- targets == sources is the common case with one such pair per target --
  // none between THIS CTA's targets and THIS launch's chunk (coincidence.cuh).  Every operand of the decision comes
  // from the constant bank or blockIdx, so the branch is uniform and both loops keep their LDCU operands.
  bool unmasked = false;
  if (a.allow_unmasked) {
    unmasked = c_eval_must_mask[0] == 0;
    if (!unmasked && a.bits_off >= 0 && c_eval_must_mask[1] == 0) {
      const unsigned w = reinterpret_cast<const unsigned*>(c_src_pack)[a.bits_off + (int)(blockIdx.x >> 5)];
      unmasked = ((w >> (blockIdx.x & 31)) & 1u) == 0;
    }
  }
  if constexpr (const_packed<Real, KID, T>()) {
    // targets 2p and 2p+1 of this thread share one packed value; same operations per target as the scalar loop
    f32x2 xt2[T / 2][3], acc2[T / 2][K1];
    int lo[U][T / 2] = {};
#pragma unroll
    for (int p = 0; p < T / 2; p++) {
#pragma unroll
      for (int k = 0; k < 3; k++) xt2[p][k] = f32x2(xt[2 * p][k], xt[2 * p + 1][k]);
#pragma unroll
      for (int k = 0; k < K1; k++) acc2[p][k] = f32x2(0.0f);
    }
    constexpr bool BLOCKED = SCTL_CUDA_CONST32_BLOCKED_LOOP != 0;
    if (unmasked) const_pair_loop<f32x2, Real, KID, T / 2, RSQ, U, false, BLOCKED>(acc2, xt2, src, n, f32x2((float)a.ctx0), lo, f32x2(0.0f));
    else const_pair_loop<f32x2, Real, KID, T / 2, RSQ, U, true, BLOCKED>(acc2, xt2, src, n, f32x2((float)a.ctx0), lo, f32x2(0.0f));
#pragma unroll
    for (int p = 0; p < T / 2; p++)
#pragma unroll
      for (int k = 0; k < K1; k++) {
        acc[2 * p][k] = acc2[p][k].v.x;
        acc[2 * p + 1][k] = acc2[p][k].v.y;
      }
  } else if constexpr (SCTL_CUDA_CONST_SWP && sizeof(Real) == 8 && !KerNeedsCtx<KID>::value) {
    if (unmasked) const_pair_loop_swp<KID, T, RSQ, U, false>(acc, xt, src, n);
    else const_pair_loop_swp<KID, T, RSQ, U, true>(acc, xt, src, n);
  } else {
    // The refinement's loop-invariant operands (pair_math.cuh, rsqrt_masked): handed in as kernel arguments so that
    // ptxas cannot fold them back into literals.  One zero low word per unrolled pair: each then lives in the low half of
    // the register pair its MUFU.RSQ64H writes the high half of, for the whole loop.
    constexpr bool HOIST = ConstHoist<Real, KID>::value;
    constexpr bool BLOCKED = HOIST || !SCTL_CUDA_CONST_PRAGMA_LOOP;
    int lo[U][T];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < T; j++) lo[u][j] = HOIST ? a.zero : 0;
    const Real k = HOIST ? a.rsq_k : (Real)rsq_poly_k<RSQ>();
    if (unmasked) const_pair_loop<Real, Real, KID, T, RSQ, U, false, BLOCKED>(acc, xt, src, n, a.ctx0, lo, k);
    else const_pair_loop<Real, Real, KID, T, RSQ, U, true, BLOCKED>(acc, xt, src, n, a.ctx0, lo, k);
  }
  if (gridDim.y > 1) {
    // grid.y == cluster size: rank 0 adds the other ranks' partial sums in rank order
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    if (blockIdx.y != 0) {
#pragma unroll
      for (int j = 0; j < T; j++)
#pragma unroll
        for (int k = 0; k < K1; k++) red[(j * K1 + k) * NT + threadIdx.x] = acc[j][k];
    }
    cluster.sync();
    if (blockIdx.y == 0) {
      for (unsigned r = 1; r < gridDim.y; r++) {
        const Real* remote = cluster.map_shared_rank(red, r);
#pragma unroll
        for (int j = 0; j < T; j++)
#pragma unroll
          for (int k = 0; k < K1; k++) acc[j][k] += remote[(j * K1 + k) * NT + threadIdx.x];
      }
    }
    cluster.sync();  // keep the remote buffers alive until rank 0 has read them
    if (blockIdx.y != 0) return;
  }
#pragma unroll
  for (int j = 0; j < T; j++) {
    const int64_t t = t_base + j * 32;
    if (t < a.Nt) {
#pragma unroll
      for (int k = 0; k < K1; k++) a.v_trg[t * K1 + k] = fma(acc[j][k], a.scale, a.v_trg[t * K1 + k]);
    }
  }
}

}  // namespace sctl_cuda
