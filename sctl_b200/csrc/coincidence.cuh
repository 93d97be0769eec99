// Can any (source, target) pair have r^2 below the smallest normal number?
//
// The pair kernels mask r^2 == 0 (and subnormal r^2) pairs with a compare + select per pair -- two register-reading
// instructions that cost ~3 % of a Stokes3D_FxU pair on B200.  The mask is only ever needed when a source and a
// target (nearly) coincide, so for big problems the host first runs this O(Ns + Nt) check: every source is inserted
// into an open-addressing hash set keyed on its coordinate bit patterns, every target probes it, and a hit raises a
// device flag that the pair kernels read to choose between their masked and unmasked loops.  No host round trip.
//
// Completeness: r^2 = dx^2 + dy^2 + dz^2 < 2^-1022 needs |d_k| < 1.5e-154 in every coordinate; two DISTINCT doubles
// that close are both smaller than 1.5e-154 * 2^52 = 6.7e-139.  Coordinates below TINY (1e-130; FP32: r^2 < 2^-126
// needs |d_k| < 1.1e-19, distinct floats that close are below 1.9e-12, TINY = 1e-11) are therefore flushed to zero in
// the KEY, so any such pair has equal keys and is flagged.  A false positive (hash collision, tiny coordinates)
// only selects the masked loop, which is always correct.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sctl_cuda {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

template <class Real> __device__ __forceinline__ uint64_t coord_bits(Real c);
template <> __device__ __forceinline__ uint64_t coord_bits<double>(double c) { return fabs(c) < 1e-130 ? 0ULL : (uint64_t)__double_as_longlong(c); }
template <> __device__ __forceinline__ uint64_t coord_bits<float>(float c) { return fabsf(c) < 1e-11f ? 0ULL : (uint64_t)__float_as_uint(c); }

template <class Real> __device__ __forceinline__ uint64_t point_key(const Real* p) {
  uint64_t k = mix64(coord_bits(p[0]) + 0x9E3779B97F4A7C15ULL);
  k = mix64(k ^ (coord_bits(p[1]) + 0xD1B54A32D192ED03ULL));
  k = mix64(k ^ (coord_bits(p[2]) + 0x8CB92BA72F3D8DD7ULL));
  return k ? k : 1;  // 0 marks an empty slot
}

// Per-block refinement (constant-bank path).  Coincident pairs are rare even when they exist -- targets == sources has
// one per target -- so the probe can also say WHERE they are: the insert pass records, per hash slot, the range of
// source CHUNKS (constant-bank launches) holding that point, and a target that hits a slot raises one bit per
// (chunk, target block) in `bits` (chunk-major rows of `row_words` 32-bit words, at `bits + chunk * row_stride`).  A
// CTA of launch c then takes the unmasked loop unless its own bit in row c is set.  A point duplicated over more than
// MAX_CHUNK_SPAN chunks gives up on the refinement (flag[1]) and the global flag alone decides.
struct CoincidenceBits {
  unsigned* bits;       // nullptr: no refinement
  int64_t row_stride;   // 32-bit words between the rows of consecutive chunks
  int64_t chunk_len;    // sources per chunk
  int64_t trg_per_blk;  // targets per CTA (one bit each)
  unsigned* cmin;       // [slots] first / last chunk holding the slot's point
  unsigned* cmax;
};
constexpr unsigned MAX_CHUNK_SPAN = 8;

// table: (mask + 1) zero-initialised 64-bit slots, mask + 1 = power of two >= 2 * Ns; cmin initialised to 0xFFFFFFFF, cmax to 0
template <class Real> __global__ void coincidence_insert_kernel(int64_t Ns, const Real* __restrict__ r_src, unsigned long long* table, uint64_t mask, const CoincidenceBits cb) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < Ns; s += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = point_key(r_src + s * 3);
    uint64_t slot = key & mask;
    while (true) {
      const unsigned long long prev = atomicCAS(&table[slot], 0ULL, (unsigned long long)key);
      if (prev == 0ULL || prev == key) break;
      slot = (slot + 1) & mask;
    }
    if (cb.bits) {
      const unsigned c = (unsigned)(s / cb.chunk_len);
      atomicMin(&cb.cmin[slot], c);
      atomicMax(&cb.cmax[slot], c);
    }
  }
}

// flag[0] becomes 1 when some target's key is in the table; flag[1] when the per-block refinement cannot be trusted
template <class Real>
__global__ void coincidence_probe_kernel(int64_t Nt, const Real* __restrict__ r_trg, const unsigned long long* __restrict__ table, uint64_t mask, int* flag,
                                         const CoincidenceBits cb) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < Nt; t += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = point_key(r_trg + t * 3);
    uint64_t slot = key & mask;
    while (true) {
      const unsigned long long v = table[slot];
      if (v == 0ULL) break;
      if (v == key) {
        flag[0] = 1;
        if (cb.bits) {
          const unsigned c0 = cb.cmin[slot], c1 = cb.cmax[slot];
          if (c1 - c0 >= MAX_CHUNK_SPAN) {
            flag[1] = 1;
          } else {
            const int64_t b = t / cb.trg_per_blk;
            for (unsigned c = c0; c <= c1; c++) atomicOr(cb.bits + (int64_t)c * cb.row_stride + (b >> 5), 1u << (b & 31));
          }
        }
        break;
      }
      slot = (slot + 1) & mask;
    }
  }
}

}  // namespace sctl_cuda
