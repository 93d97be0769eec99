// KernelMatrix on one B200:  M[(s*K0+k0)*ldM + t*K1+k1] = scale * u_{k0,k1}(x_t - x_s, n_s), overwriting.
// Replaces GenericKernel<uKernel>::KernelMatrix (reference: include/sctl/generic-kernel.txx:213-329; layout
// :268,306).  Store-bound: 8*K0*K1 bytes written per pair against ~25 flop, so the kernel is shaped for the
// write stream -- one CTA owns BT consecutive targets of a few source rows, the K0 x K1 blocks of a row are
// staged through shared memory and written out as contiguous, fully coalesced row segments.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pair_math.cuh"

namespace sctl_cuda {

template <class Real, int KID, int RSQ, int BT>
__global__ void __launch_bounds__(BT) kernel_matrix_kernel(int64_t Ns, int64_t Nt, const Real* __restrict__ r_src, const Real* __restrict__ n_src,
                                                         const Real* __restrict__ r_trg, Real* __restrict__ M, int64_t ldM, Real scale, Real mu) {
  constexpr int K0 = KerTraits<KID>::K0, K1 = KerTraits<KID>::K1, ND = KerTraits<KID>::ND;
  __shared__ Real stage[K0][BT * K1 + 1];
  const int64_t t_blk = (int64_t)blockIdx.x * BT;
  const int64_t t = t_blk + threadIdx.x;
  const int nt_blk = (int)((Nt - t_blk) < BT ? (Nt - t_blk) : BT);
  Real xt[3] = {0, 0, 0};
  if (t < Nt) {
#pragma unroll
    for (int k = 0; k < 3; k++) xt[k] = r_trg[t * 3 + k];
  }
  for (int64_t s = blockIdx.y; s < Ns; s += gridDim.y) {
    Real dx[3], n[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; k++) dx[k] = xt[k] - r_src[s * 3 + k];
    if (ND) {
#pragma unroll
      for (int k = 0; k < 3; k++) n[k] = n_src[s * 3 + k];
    }
    const Real r2 = fma(dx[2], dx[2], fma(dx[1], dx[1], dx[0] * dx[0]));
    const Real rinv = rsqrt_masked<RSQ>(r2);
    Real u[K0 * K1];
    if constexpr (KerNeedsCtx<KID>::value) PairMatrix<KID, Real>::run_ctx(u, dx, rinv, n, r2, mu);
    else PairMatrix<KID, Real>::run(u, dx, rinv, n);
    if (K1 == 1) {
      // one entry per (row, target): the direct store is already coalesced
      if (t < Nt) {
#pragma unroll
        for (int k0 = 0; k0 < K0; k0++) M[(s * K0 + k0) * ldM + t] = u[k0] * scale;
      }
    } else {
      __syncthreads();  // previous row's readers are done
#pragma unroll
      for (int k0 = 0; k0 < K0; k0++)
#pragma unroll
        for (int k1 = 0; k1 < K1; k1++) stage[k0][threadIdx.x * K1 + k1] = u[k0 * K1 + k1] * scale;
      __syncthreads();
#pragma unroll
      for (int k0 = 0; k0 < K0; k0++) {
        Real* row = M + (s * K0 + k0) * ldM + t_blk * K1;
        for (int c = threadIdx.x; c < nt_blk * K1; c += BT) row[c] = stage[k0][c];
      }
    }
  }
}

}  // namespace sctl_cuda
