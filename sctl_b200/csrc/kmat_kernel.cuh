// KernelMatrix on one B200:  M[(s*K0+k0)*ldM + t*K1+k1] = scale * u_{k0,k1}(x_t - x_s, n_s), overwriting.
// Replaces GenericKernel<uKernel>::KernelMatrix (reference: include/sctl/generic-kernel.txx:213-329; layout
// :268,306).  Store-bound: sizeof(Real)*K0*K1 bytes written per pair against 13..50 flop, so the kernel is shaped for
// the HBM write stream:
//   * a CTA owns BT*T consecutive targets (register-resident, T per thread) of a contiguous range of source rows, so
//     every matrix row it touches is one contiguous segment of BT*T*K1 Reals;
//   * every global store is a 16-byte vector.  K1 == 1: a thread's T = 16/sizeof(Real) consecutive targets are one
//     vector, stored straight from registers (a warp writes 512 contiguous bytes per instruction); where rows do not all
//     start on a 16-byte boundary (odd ldM, matrix views: KMAT_SHIFTED) every lane completes the aligned vector inside
//     its columns with the first values of the next lane (one __shfl_down per value).  K1 > 1: the K0 x (T*K1) values
//     of a thread are staged through shared memory, WARP by warp: a warp owns the 32*T*K1 contiguous columns of its
//     lanes' targets, stages them in its own slice of shared memory and writes them out itself as aligned vectors,
//     scalar elements only before the first and after the last 16-byte boundary of its part -- rows need not be 16-byte
//     aligned, the staging index is shifted per row so that the shared-memory side of every vector is aligned too.
//     Only __syncwarp() orders the two phases: the row loop has no CTA barrier, so warps that are computing (the
//     Helmholtz kernels spend ~70 FP64 instructions per pair) never wait for warps that are storing;
//   * source coordinates (and normals) of a tile of rows are brought into shared memory by the whole CTA before the
//     tile's rows are computed, so the row loop has no global load in it (round 1's kernel loaded them per row and
//     spent its time on the long scoreboard: profiles/r1_ncu_kernel_matrix.md, 62 % / 40 % of the HBM copy bandwidth
//     for Laplace3D_FxU FP64 / FP32).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pair_math.cuh"

namespace sctl_cuda {

constexpr int KMAT_BT = 128;      // threads per CTA
constexpr int KMAT_SRC_TILE = 32;  // source rows whose coordinates are staged together

template <class Real> struct KmatVec;
template <> struct KmatVec<double> { using type = double2; };
template <> struct KmatVec<float> { using type = float4; };

// How a thread's values reach M.  KMAT_STAGED: K1 > 1.  K1 == 1: KMAT_DIRECT when every row starts on a 16-byte boundary (M
// and the row pitch aligned; chosen by the host), KMAT_SHIFTED otherwise.
enum : int { KMAT_STAGED = 0, KMAT_DIRECT = 1, KMAT_SHIFTED = 2 };

// Targets per thread T: K1 == 1 -> one 16-byte vector of consecutive targets; K1 > 1 -> 8 bytes' worth of targets (the
// stage holds K0 * BT * T * K1 Reals: 27 KB for the K0 x K1 = 3 x 9 kernels in either precision).
// Rows per round R: the source rows a thread computes before it stores.  Direct forms: 2 (independent pairs in flight);
// staged: 1, the K1 > 1 kernels do enough work per row.
template <class Real, int KID, bool DIRECT> struct KmatShape {
  static constexpr int V = 16 / (int)sizeof(Real);
  static constexpr int T = KerTraits<KID>::K1 == 1 ? V : 8 / (int)sizeof(Real);
  static constexpr int R = DIRECT ? 2 : 1;
};

template <class Real, int KID, int RSQ, int MODE>
__global__ void __launch_bounds__(KMAT_BT) kernel_matrix_kernel(int64_t Ns, int64_t Nt, int64_t rows_per_cta, const Real* __restrict__ r_src, const Real* __restrict__ n_src,
                                                              const Real* __restrict__ r_trg, Real* __restrict__ M, int64_t ldM, Real scale, Real mu) {
  constexpr int K0 = KerTraits<KID>::K0, K1 = KerTraits<KID>::K1, ND = KerTraits<KID>::ND;
  constexpr bool DIRECT = MODE != KMAT_STAGED;
  using Shape = KmatShape<Real, KID, DIRECT>;
  constexpr int BT = KMAT_BT, SB = KMAT_SRC_TILE, V = Shape::V, T = Shape::T, R = Shape::R;
  using Vec = typename KmatVec<Real>::type;
  static_assert(DIRECT == (K1 == 1) && (!DIRECT || T == V), "K1 == 1: one vector of consecutive targets per row, straight from registers; K1 > 1: staged");
  static_assert(SB % R == 0, "a source tile is a whole number of rounds");
  constexpr int WW = 32 * T * K1;  // columns of a warp's part of the segment
  __shared__ Real xs_s[SB][3];
  __shared__ Real ns_s[ND ? SB : 1][3];
  __shared__ __align__(16) Real stage[DIRECT ? 1 : BT / 32][DIRECT ? 1 : R * K0][DIRECT ? V : WW + V];

  const int64_t t_blk = (int64_t)blockIdx.x * (BT * T);
  const int64_t t0 = t_blk + (int64_t)threadIdx.x * T;
  Real xt[T][3];
#pragma unroll
  for (int j = 0; j < T; j++)
#pragma unroll
    for (int k = 0; k < 3; k++) xt[j][k] = (t0 + j < Nt) ? r_trg[(t0 + j) * 3 + k] : (Real)0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t c_warp = t_blk * K1 + (int64_t)warp * WW;  // first column of this warp's part
  const int64_t left = Nt * K1 - c_warp;
  const int n_col = (int)(left < 0 ? 0 : (left < WW ? left : WW));  // valid columns of this warp's part
  // alignment of this warp's part (and, K1 == 1, of every thread's first column).  a0 = elements between the last 16-byte
  // boundary and the part's first column in matrix row 0, lmod = ldM mod V; `simple`: all rows of the part aligned and the
  // part full
  constexpr int NV = WW / V;  // vectors per full part
  const int a0 = (int)(((uintptr_t)(M + c_warp) / sizeof(Real)) & (V - 1));
  const int lmod = (int)(ldM & (V - 1));
  const bool simple = a0 == 0 && lmod == 0 && n_col == WW;
  const int64_t s_begin = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t s_end = (s_begin + rows_per_cta < Ns) ? s_begin + rows_per_cta : Ns;

  for (int64_t tile = s_begin; tile < s_end; tile += SB) {
    const int nsb = (int)((s_end - tile) < SB ? (s_end - tile) : SB);
    __syncthreads();  // the previous tile's readers are done
    for (int i = threadIdx.x; i < nsb * 3; i += BT) {
      xs_s[i / 3][i % 3] = r_src[tile * 3 + i];
      if (ND) ns_s[i / 3][i % 3] = n_src[tile * 3 + i];
    }
    __syncthreads();
    if (n_col == 0) continue;  // (warp-uniform; no barrier below) nothing of this warp's part lies inside the matrix
#pragma unroll(R == 1 ? 2 : 1)
    for (int si0 = 0; si0 < nsb; si0 += R) {
      Real u[R][T][K0 * K1];
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int si = si0 + r < nsb ? si0 + r : nsb - 1;  // a ragged last round recomputes the last row (not stored)
        const Real xs[3] = {xs_s[si][0], xs_s[si][1], xs_s[si][2]};
        Real n[3] = {0, 0, 0};
        if (ND) {
#pragma unroll
          for (int k = 0; k < 3; k++) n[k] = ns_s[si][k];
        }
#pragma unroll
        for (int j = 0; j < T; j++) {
          const Real dx[3] = {xt[j][0] - xs[0], xt[j][1] - xs[1], xt[j][2] - xs[2]};
          const Real r2 = fma(dx[2], dx[2], fma(dx[1], dx[1], dx[0] * dx[0]));
          const Real rinv = rsqrt_masked<RSQ>(r2);
          if constexpr (KerNeedsCtx<KID>::value) PairMatrix<KID, Real>::run_ctx(u[r][j], dx, rinv, n, r2, mu);
          else PairMatrix<KID, Real>::run(u[r][j], dx, rinv, n);
#pragma unroll
          for (int k = 0; k < K0 * K1; k++) u[r][j][k] *= scale;
        }
      }
      const int n_rows = ((nsb - si0) < R ? (nsb - si0) : R) * K0;  // matrix rows of this round
      const int64_t row0 = (tile + si0) * K0;
      if constexpr (DIRECT) {
        // K1 == 1: a thread's T == V consecutive targets of a row.  Every thread's first column is a multiple of V past
        // the CTA's first, so all of them sit `pad` elements past a 16-byte boundary in this row (pad depends on the row
        // only).  pad == 0: the thread's values are one aligned vector.  Otherwise the aligned vector that starts inside
        // the thread's columns is its last pad values followed by the first mis = V - pad values of the next lane
        // (__shfl_down): lane i stores that vector, lane 0 also stores its own first mis values one by one and lane 31 its
        // last pad values (the vector between two warps is written from both sides).
#pragma unroll
        for (int rr = 0; rr < R * K0; rr++) {
          if (rr < n_rows) {
            Real* p = M + (row0 + rr) * ldM + t0;
            const int pad = MODE == KMAT_DIRECT ? 0 : (a0 + ((int)(row0 + rr) & (V - 1)) * lmod) & (V - 1);
            Real own[V];
#pragma unroll
            for (int j = 0; j < V; j++) own[j] = u[rr / K0][j][rr % K0];
            if (MODE == KMAT_DIRECT || pad == 0) {
              if (t0 + T <= Nt) {
                *reinterpret_cast<Vec*>(p) = *reinterpret_cast<const Vec*>(own);
              } else {
#pragma unroll
                for (int j = 0; j < T; j++)
                  if (t0 + j < Nt) p[j] = own[j];
              }
            } else {
              const int mis = V - pad;
              Real nb[V - 1];
#pragma unroll
              for (int j = 0; j < V - 1; j++) nb[j] = __shfl_down_sync(0xffffffffu, own[j], 1);
              Real vec[V];
#pragma unroll
              for (int m = 1; m < V; m++)
                if (mis == m) {
#pragma unroll
                  for (int e = 0; e < V; e++) vec[e] = (e < V - m) ? own[m + e] : nb[e - (V - m)];
                }
              const int64_t c = t0 + mis;  // first column of the vector
              if (lane == 0) {
#pragma unroll
                for (int j = 0; j < V - 1; j++)
                  if (j < mis && t0 + j < Nt) p[j] = own[j];
              }
              if (lane < 31) {
                if (c + V <= Nt) {
                  *reinterpret_cast<Vec*>(p + mis) = *reinterpret_cast<const Vec*>(vec);
                } else {
#pragma unroll
                  for (int e = 0; e < V; e++)
                    if (c + e < Nt) p[mis + e] = vec[e];
                }
              } else {
#pragma unroll
                for (int e = 0; e < V - 1; e++)
                  if (e < pad && c + e < Nt) p[mis + e] = vec[e];
              }
            }
          }
        }
      } else {
        // This warp's part of matrix row `row` starts at wrow = M + row*ldM + c_warp, `pad` elements past a 16-byte
        // boundary.  Column c of the part is staged at index pad + c: the stage rows are 16-byte aligned, so global and
        // shared addresses of a column are congruent modulo 16 and the columns from mis = (V - pad) % V on move as
        // aligned vectors on both sides.  pad = (a0 + row * ldM) mod V needs only the low bits of its terms.
        Real* wrow = M + row0 * ldM + c_warp;
        __syncwarp();  // this warp's reads of the previous round are done
        if (simple) {
          // every row of the part starts on a 16-byte boundary and the part is full: vectors only, no index arithmetic
#pragma unroll
          for (int rr = 0; rr < R * K0; rr++) {
            Real* st = &stage[warp][rr][lane * (T * K1)];
#pragma unroll
            for (int j = 0; j < T; j++)
#pragma unroll
              for (int k1 = 0; k1 < K1; k1++) st[j * K1 + k1] = u[rr / K0][j][(rr % K0) * K1 + k1];
          }
          __syncwarp();
#pragma unroll
          for (int rr = 0; rr < R * K0; rr++) {
            if (rr < n_rows) {
              const Vec* sv = reinterpret_cast<const Vec*>(&stage[warp][rr][0]);
              Vec* gv = reinterpret_cast<Vec*>(wrow);
#pragma unroll
              for (int p = 0; p < (NV + 31) / 32; p++)
                if (p * 32 + 32 <= NV || lane < NV - p * 32) gv[p * 32 + lane] = sv[p * 32 + lane];
            }
            wrow += ldM;
          }
        } else {
          int pads[R * K0];
#pragma unroll
          for (int rr = 0; rr < R * K0; rr++) {
            pads[rr] = (a0 + ((int)(row0 + rr) & (V - 1)) * lmod) & (V - 1);
            Real* st = &stage[warp][rr][pads[rr] + lane * (T * K1)];
#pragma unroll
            for (int j = 0; j < T; j++)
#pragma unroll
              for (int k1 = 0; k1 < K1; k1++) st[j * K1 + k1] = u[rr / K0][j][(rr % K0) * K1 + k1];
          }
          __syncwarp();
#pragma unroll
          for (int rr = 0; rr < R * K0; rr++) {
            if (rr < n_rows) {
              const int mis = (V - pads[rr]) & (V - 1);
              const int head = mis < n_col ? mis : n_col;
              const int tail0 = head + ((n_col - head) & ~(V - 1));
              const Real* st = &stage[warp][rr][pads[rr]];
              if (lane < head) wrow[lane] = st[lane];
#pragma unroll
              for (int p = 0; p < (NV + 31) / 32; p++) {
                const int e = head + (p * 32 + lane) * V;
                if (e + V <= tail0) *reinterpret_cast<Vec*>(wrow + e) = *reinterpret_cast<const Vec*>(st + e);
              }
              if (tail0 + lane < n_col) wrow[tail0 + lane] = st[tail0 + lane];
            }
            wrow += ldM;
          }
        }
      }
    }
  }
}

}  // namespace sctl_cuda
