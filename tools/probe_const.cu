// tools/probe_const.cu -- constant-cache behaviour for the chunk-streaming design (tuning aid, not product code).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
constexpr int ND = (64 * 1024 - 1024) / 8;
__constant__ double cbuf[ND];

// pass structure per warp: [optional parallel warm-up of a 1/nw slice] then 2 full sequential passes
template <int WARM> __global__ void k(long long* out, double* sink, int stride) {
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double s = 0;
  long long t0 = clock64();
  if (WARM == 1) {  // each warp touches its own slice, one double per `stride` doubles
    const int per = ND / nw;
    for (int i = warp * per; i < (warp + 1) * per; i += stride) s += cbuf[i];
  }
  if (WARM == 2) {  // vector LDC: lane-distinct addresses, one per 64-byte line
    for (int i = (threadIdx.x) * 8; i < ND; i += blockDim.x * 8) s += cbuf[i];
  }
  __syncthreads();
  long long t1 = clock64();
  for (int i = 0; i < ND; i += 6) s += cbuf[i] * cbuf[i + 1] + cbuf[i + 2] * cbuf[i + 3] + cbuf[i + 4] * cbuf[i + 5];
  long long t2 = clock64();
  for (int i = 0; i < ND; i += 6) s += cbuf[i] * cbuf[i + 1] + cbuf[i + 2] * cbuf[i + 3] + cbuf[i + 4] * cbuf[i + 5];
  long long t3 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; }
  if (s == -1.234) sink[0] = s;
}

int main() {
  double* h = (double*)malloc(ND * 8);
  for (int i = 0; i < ND; i++) h[i] = 1.0 + i * 1e-6;
  long long* out; double* sink; double* dsrc;
  CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&sink, 8)); CK(cudaMalloc(&dsrc, ND * 8));
  CK(cudaMemcpy(dsrc, h, ND * 8, cudaMemcpyHostToDevice));
  long long r[3];
  for (int threads : {128, 512}) {
    for (int mode = 0; mode < 5; mode++) {
      for (int rep = 0; rep < 2; rep++) {
        CK(cudaMemcpyToSymbolAsync(cbuf, dsrc, ND * 8, 0, cudaMemcpyDeviceToDevice));  // new contents => cold constant caches
        if (mode == 0) k<0><<<148, threads>>>(out, sink, 8);
        if (mode == 1) k<1><<<148, threads>>>(out, sink, 8);   // one touch per 64 B
        if (mode == 2) k<1><<<148, threads>>>(out, sink, 16);  // one touch per 128 B
        if (mode == 3) k<1><<<148, threads>>>(out, sink, 32);  // one touch per 256 B
        if (mode == 4) k<2><<<148, threads>>>(out, sink, 8);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(r, out, 24, cudaMemcpyDeviceToHost));
      }
      printf("threads=%d mode=%d  warmup %lld cyc | first full pass %lld cyc | second full pass %lld cyc  (%d doubles; %.1f / %.1f cyc per 48-byte record)\n", threads, mode,
             r[0], r[1], r[2], ND, r[1] / (ND / 6.0), r[2] / (ND / 6.0));
    }
  }
  // back-to-back launches without rewriting the constant bank: does the cache stay warm across launches?
  CK(cudaMemcpyToSymbol(cbuf, dsrc, ND * 8, 0, cudaMemcpyDeviceToDevice));
  for (int rep = 0; rep < 3; rep++) {
    k<0><<<148, 128>>>(out, sink, 8);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(r, out, 24, cudaMemcpyDeviceToHost));
    printf("no rewrite, launch %d: first pass %lld second pass %lld\n", rep, r[1], r[2]);
  }
  return 0;
}
