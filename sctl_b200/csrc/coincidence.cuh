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

// table: (mask + 1) zero-initialised 64-bit slots, mask + 1 = power of two >= 2 * Ns
template <class Real> __global__ void coincidence_insert_kernel(int64_t Ns, const Real* __restrict__ r_src, unsigned long long* table, uint64_t mask) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < Ns; s += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = point_key(r_src + s * 3);
    uint64_t slot = key & mask;
    while (true) {
      const unsigned long long prev = atomicCAS(&table[slot], 0ULL, (unsigned long long)key);
      if (prev == 0ULL || prev == key) break;
      slot = (slot + 1) & mask;
    }
  }
}

// must_mask[0] becomes 1 when some target's key is in the table
template <class Real>
__global__ void coincidence_probe_kernel(int64_t Nt, const Real* __restrict__ r_trg, const unsigned long long* __restrict__ table, uint64_t mask, int* must_mask) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < Nt; t += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = point_key(r_trg + t * 3);
    uint64_t slot = key & mask;
    while (true) {
      const unsigned long long v = table[slot];
      if (v == 0ULL) break;
      if (v == key) {
        must_mask[0] = 1;
        break;
      }
      slot = (slot + 1) & mask;
    }
  }
}

}  // namespace sctl_cuda
