// Batched near-field blocks of BoundaryIntegralOp on the device.
//
// Reference (paths relative to /root/reference/include/sctl/):
//   setup  boundary_integral.txx:1165-1240 ("KNearSubtract"): for every near block, KernelMatrix of the element's
//          far-field nodes against the block's near targets, rows scaled by the quadrature weights, transposed and
//          subtracted from the precomputed target-major operator block K_near_;
//   apply  boundary_integral.txx:1324-1340 ("EvalNear"): per block U_ = K_near_ * F_ (row-major GEMV).
// The reference walks the block list with OpenMP and calls KernelMatrix / GEMM once per block (hundreds x hundreds
// each); through a host-pointer drop-in that is one H2D + launch + D2H per block.  Here the whole list is ONE launch
// per phase over a (block, row-chunk) work list, and the operator stays in HBM between applies.
//
// Layouts (all the reference's): block b has ns_b source nodes and nt_b targets; K_b is row-major
// (nt_b*K1) x (ns_b*K0) at K + k_dsp[b]  (target-major, :1183); F_b = F + f_dsp[b]*K0; U_b = U + u_dsp[b]*K1.
// Bounds: subtract is HBM read-modify-write of K (2 * sizeof(Real) * K0*K1 bytes per pair, ~25 flop per pair);
// apply is an HBM read of K (sizeof(Real) bytes per 2 flop).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pair_math.cuh"

namespace sctl_cuda {

struct NearBlocks {
  int64_t n_blk;
  const int64_t* f_dsp;    // [n_blk] first source node of the block in F (and in the block's element)
  const int64_t* far_dsp;  // [n_blk] first far-field node of the block in Xs / Xn / wts
  const int64_t* ns;       // [n_blk] source nodes
  const int64_t* u_dsp;    // [n_blk] first target of the block in Xt / U
  const int64_t* nt;       // [n_blk] targets
  const int64_t* k_dsp;    // [n_blk] offset of K_b in K (Reals)
  // work list: item w covers rows/targets [item_t0[w], item_t0[w] + ROWS) of block item_blk[w]
  int64_t n_items;
  const int32_t* item_blk;
  const int32_t* item_t0;
};

constexpr int NEAR_SUB_TRG = 32;    // targets per work item in the subtract kernel
constexpr int NEAR_SUB_THREADS = 128;
constexpr int NEAR_APPLY_WARPS = 8;  // warps per CTA in the apply kernel
// rows per warp, processed together (independent loads in flight, F loaded once): 32 bytes of each column step per lane
template <class Real> __host__ __device__ constexpr int near_apply_rpw() { return 32 / (int)sizeof(Real); }
template <class Real> __host__ __device__ constexpr int near_apply_rows() { return NEAR_APPLY_WARPS * near_apply_rpw<Real>(); }  // matrix rows per work item

// K_b[(t*K1+k1)][(s*K0+k0)] -= wts[s] * scale * K_{k0,k1}(x_t - x_s, n_s)
// One CTA per work item (NEAR_SUB_TRG targets of one block).  The CTA's threads are laid out as (source lane, target
// group): lanes cover the block's sources (for a fixed row they touch consecutive columns: coalesced), groups take
// alternate targets, so short source lists (64 nodes per element is typical) still use every thread.  The kernel is
// an HBM read-modify-write stream of K: each thread issues the loads of TU targets' entries before it computes and
// stores, which keeps several independent DRAM requests in flight per thread (a plain `*p -= x` loop serialises one
// round trip per target).  The pair arithmetic is the KernelMatrix twin PairMatrix<KID> (pair_math.cuh), scaled per
// entry like generic-kernel.txx:268,306, then by the weight (:1203).
// NDOT: the trg_normal_dot_prod variant (:1186-1200): the kernel's K1 = 3*K1r target components are contracted with the
// target normal, entry (k0, k1r) = sum_l u[k0][k1r*3+l] * n_t[l], and the operator block has K1r rows per target.
template <class Real, int KID, int RSQ, bool NDOT>
__global__ void __launch_bounds__(NEAR_SUB_THREADS) near_subtract_kernel(const NearBlocks nb, const Real* __restrict__ Xs, const Real* __restrict__ Xn, const Real* __restrict__ wts,
                                                                         const Real* __restrict__ Xt, const Real* __restrict__ Xn_trg, Real* __restrict__ K, Real scale, Real ctx0) {
  constexpr int K0 = KerTraits<KID>::K0, KU = KerTraits<KID>::K1, ND = KerTraits<KID>::ND;
  constexpr int K1 = NDOT ? (KU >= 3 ? KU / 3 : 1) : KU;  // rows per target in the operator block
  constexpr int KK = K0 * K1;
  constexpr int TU = KK == 1 ? 8 : (KK == 2 ? 4 : (KK <= 4 ? 2 : 1));  // targets whose entries are loaded together
  __shared__ Real xt_s[NEAR_SUB_TRG][3];
  __shared__ Real nt_s[NEAR_SUB_TRG][3];
  for (int64_t w = blockIdx.x; w < nb.n_items; w += gridDim.x) {
    const int b = nb.item_blk[w];
    const int64_t t0 = nb.item_t0[w];
    const int64_t ns = nb.ns[b], nt = nb.nt[b];
    const int n_t = (int)min((int64_t)NEAR_SUB_TRG, nt - t0);
    __syncthreads();
    if (threadIdx.x < n_t * 3) {  // NEAR_SUB_TRG * 3 <= NEAR_SUB_THREADS
      xt_s[threadIdx.x / 3][threadIdx.x % 3] = Xt[(nb.u_dsp[b] + t0) * 3 + threadIdx.x];
      if (NDOT) nt_s[threadIdx.x / 3][threadIdx.x % 3] = Xn_trg[(nb.u_dsp[b] + t0) * 3 + threadIdx.x];
    }
    __syncthreads();
    int lanes = NEAR_SUB_THREADS;  // source lanes: the smallest power of two >= ns, at least 32 (whole warps share a row)
    while (lanes > 32 && (lanes >> 1) >= ns) lanes >>= 1;
    const int groups = NEAR_SUB_THREADS / lanes, grp = threadIdx.x / lanes;
    const int64_t ld = ns * K0;
    Real* Kb = K + nb.k_dsp[b] + t0 * K1 * ld;
    for (int64_t s = threadIdx.x % lanes; s < ns; s += lanes) {
      const int64_t gs = nb.far_dsp[b] + s;
      const Real xs[3] = {Xs[gs * 3 + 0], Xs[gs * 3 + 1], Xs[gs * 3 + 2]};
      Real n[3] = {0, 0, 0};
      if (ND) {
        n[0] = Xn[gs * 3 + 0];
        n[1] = Xn[gs * 3 + 1];
        n[2] = Xn[gs * 3 + 2];
      }
      const Real ws = wts[gs] * scale;
      for (int tb = grp * TU; tb < n_t; tb += groups * TU) {
        Real old[TU][KK];
#pragma unroll
        for (int j = 0; j < TU; j++)
          if (tb + j < n_t) {
#pragma unroll
            for (int k1 = 0; k1 < K1; k1++)
#pragma unroll
              for (int k0 = 0; k0 < K0; k0++) old[j][k0 * K1 + k1] = Kb[(int64_t)((tb + j) * K1 + k1) * ld + s * K0 + k0];
          }
#pragma unroll
        for (int j = 0; j < TU; j++)
          if (tb + j < n_t) {
            const int t = tb + j;
            Real dx[3] = {xt_s[t][0] - xs[0], xt_s[t][1] - xs[1], xt_s[t][2] - xs[2]};
            const Real r2 = fma(dx[2], dx[2], fma(dx[1], dx[1], dx[0] * dx[0]));
            const Real rinv = rsqrt_masked<RSQ>(r2);
            Real uu[K0 * KU], u[KK];
            if constexpr (KerNeedsCtx<KID>::value) PairMatrix<KID, Real>::run_ctx(uu, dx, rinv, n, r2, ctx0);
            else PairMatrix<KID, Real>::run(uu, dx, rinv, n);
#pragma unroll
            for (int k0 = 0; k0 < K0; k0++)
#pragma unroll
              for (int k1 = 0; k1 < K1; k1++) {
                if constexpr (NDOT) u[k0 * K1 + k1] = fma(uu[k0 * KU + k1 * 3 + 2], nt_s[t][2], fma(uu[k0 * KU + k1 * 3 + 1], nt_s[t][1], uu[k0 * KU + k1 * 3 + 0] * nt_s[t][0]));
                else u[k0 * K1 + k1] = uu[k0 * KU + k1];
              }
#pragma unroll
            for (int k1 = 0; k1 < K1; k1++)
#pragma unroll
              for (int k0 = 0; k0 < K0; k0++) Kb[(int64_t)(t * K1 + k1) * ld + s * K0 + k0] = fma(-u[k0 * K1 + k1], ws, old[j][k0 * K1 + k1]);
          }
      }
    }
  }
}

// U_b[r] = sum_c K_b[r][c] * F_b[c].  A warp owns RPW consecutive rows and walks the columns once: per step one F value
// and RPW matrix values per lane, all independent loads; fixed summation order and shuffle tree per row
// (deterministic).  Register use is capped so that five (FP64) or four (FP32, eight rows per warp) CTAs stay resident per SM: the kernel is a latency-bound
// stream of short rows and needs the bytes in flight.  Blocks with no sources write zeros (boundary_integral.txx:1327-1331).
template <class Real>
__global__ void __launch_bounds__(32 * NEAR_APPLY_WARPS, (sizeof(Real) == 8 ? 5 : 4)) near_apply_kernel(const NearBlocks nb, int K0, int K1, const Real* __restrict__ K, const Real* __restrict__ F, Real* __restrict__ U) {
  constexpr int RPW = near_apply_rpw<Real>();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t w = blockIdx.x; w < nb.n_items; w += gridDim.x) {
    const int b = nb.item_blk[w];
    const int64_t r0 = (int64_t)nb.item_t0[w] + warp * RPW;  // first row of this warp in the block's matrix
    const int64_t rows = nb.nt[b] * K1, cols = nb.ns[b] * K0;
    if (r0 >= rows) continue;
    const Real* Kr = K + nb.k_dsp[b] + r0 * cols;
    const Real* Fb = F + nb.f_dsp[b] * K0;
    const int nr = (int)min((int64_t)RPW, rows - r0);
    Real acc[RPW];
#pragma unroll
    for (int j = 0; j < RPW; j++) acc[j] = 0;
    const int icols = (int)cols;  // a block row fits 32 bits; keeps the per-row offsets out of 64-bit registers
    const Real* p = Kr + lane;
    if (nr == RPW) {
      for (int c = lane; c < icols; c += 32, p += 32) {
        const Real f = Fb[c];
#pragma unroll
        for (int j = 0; j < RPW; j++) acc[j] = fma(p[j * icols], f, acc[j]);
      }
    } else {
      for (int c = lane; c < icols; c += 32, p += 32) {
        const Real f = Fb[c];
#pragma unroll
        for (int j = 0; j < RPW; j++)
          if (j < nr) acc[j] = fma(p[j * icols], f, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < RPW; j++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
    if (lane < nr) {
      Real v = acc[0];
#pragma unroll
      for (int j = 1; j < RPW; j++) v = (lane == j) ? acc[j] : v;
      U[nb.u_dsp[b] * K1 + r0 + lane] = v;
    }
  }
}

// Tail of ComputeNearInterac (boundary_integral.txx:1366-1381): comm_.ScatterForward(U_near, near_scatter_index)
// (one process: out[p] = in[scatter[p]], comm.txx:1499-1510) followed by
//   for j < near_trg_cnt[i]:  U[i*K1+k] += U_near_scattered[(near_trg_dsp[i]+j)*K1+k]
// fused into one gather: a thread owns one (target, component) of U and adds its near contributions in the reference's
// order (j ascending, onto the value U already holds), so the result does not depend on the launch geometry.
// HBM bound: per near pair one 8-byte index and K1 Reals read at scattered addresses; U read and written once.
template <class Real>
__global__ void __launch_bounds__(256) near_gather_accumulate_kernel(int64_t n_trg, int K1, const int64_t* __restrict__ cnt, const int64_t* __restrict__ dsp,
                                                                      const int64_t* __restrict__ scatter, const Real* __restrict__ U_near, Real* __restrict__ U) {
  const int64_t n = n_trg * K1;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / K1;
    const int k = (int)(idx - i * K1);
    const int64_t c = cnt[i];
    const int64_t* sp = scatter + dsp[i];
    Real u = U[idx];
    for (int64_t j = 0; j < c; j++) u += U_near[sp[j] * K1 + k];
    U[idx] = u;
  }
}

}  // namespace sctl_cuda
