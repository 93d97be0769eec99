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
// of CONST_PACK_BYTES (cudaMemcpyToSymbolAsync, device to device) and launches this kernel once per chunk; every
// launch adds scale * (chunk's partial sum) to v_trg.  Targets stay register-resident for the whole chunk, T per
// thread.  When there are too few target blocks to balance 148 SMs, the chunk is split CL ways along grid.y and the
// CL CTAs that own the same targets form a thread-block cluster: partial sums are combined through distributed
// shared memory in rank order (deterministic, no atomics, no scratch traffic).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pair_math.cuh"

namespace sctl_cuda {

constexpr int CONST_PACK_BYTES = 64 * 1024 - 1024;  // leave room for the compiler's own bank-3 constants
__constant__ double c_src_pack[CONST_PACK_BYTES / 8];
// [0]: the coincidence probe's verdict for the call in flight (0 = no source/target pair can have r2 < 2^-1022, so the
// unmasked loop is exact).  Copied device-to-device from the probe's flag, so the choice needs no host round trip, and
// read with LDCU, so the branch on it is uniform and the loops keep their uniform-register operands.
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
  int allow_unmasked; // this call ran the coincidence probe: c_eval_must_mask is meaningful
  Real scale;
  Real ctx0;          // kernel context by value (Helmholtz wavenumber)
};

template <class Real, int KID> constexpr int const_chunk_records() { return CONST_PACK_BYTES / (RecTraits<Real, KID>::REC * (int)sizeof(Real)); }

// Multi-bank streaming: the buffer is cut into `banks` equal slices and chunk i goes through slice i % banks on its
// own stream, so the launch for chunk i+1 is already queued (its slice filled) while chunk i runs and its CTAs take
// over the SMs as chunk i's last wave drains -- no idle tail and no wave quantisation per launch.  Launches that
// share a slice are ordered by their stream; launches of different slices accumulate into different buffers.
constexpr int CONST_MAX_BANKS = 4;
template <class Real, int KID> inline int const_bank_records(int banks) {
  return (CONST_PACK_BYTES / banks / 16 * 16) / (RecTraits<Real, KID>::REC * (int)sizeof(Real));
}

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
  if (a.allow_unmasked && c_eval_must_mask[0] == 0) {
#pragma unroll U
    for (int s = 0; s < n; s++) {
      const Real* rec = src + s * REC;  // uniform address: LDCU into uniform registers
#pragma unroll
      for (int j = 0; j < T; j++) pair_update<KID, RSQ, false>(acc[j], xt[j], rec, a.ctx0);
    }
  } else {
#pragma unroll U
    for (int s = 0; s < n; s++) {
      const Real* rec = src + s * REC;
#pragma unroll
      for (int j = 0; j < T; j++) pair_update<KID, RSQ, true>(acc[j], xt[j], rec, a.ctx0);
    }
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
