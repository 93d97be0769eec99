// tools/probe.cu -- FP64-pipe operand/issue probes for B200 (tuning aid, not product code).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NC = 8;  // independent chains per thread

// MODE 0: a = fma(a, m, c)           shared m, c
// MODE 1: a_i = fma(b_i, c_i, a_i)   3 distinct register operands, no sharing
// MODE 2: a_i = fma(b, c_i, a_i)     first multiplicand shared across consecutive instructions (reuse-able)
// MODE 3: a_i = fma(b_i, b_i, a_i)   2 distinct
// MODE 4: a_i = a_i * b_i            DMUL 2 distinct
// MODE 5: a_i = a_i + b_i            DADD 2 distinct
// MODE 6: a_i = fma(b_i, c, a_i)     second multiplicand shared
// MODE 7: mode 1 + one IADD3-class integer op per DFMA
// MODE 8: mode 1 + one LDS.128 broadcast per 4 DFMA
// MODE 9: a_i = fma(b_i, c_i, d_i) -> rotating destinations (no accumulate dependence on same reg)
template <int MODE> __global__ void __launch_bounds__(256) probe(int iters, const double* __restrict__ in, double* out) {
  __shared__ double2 sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = make_double2(in[threadIdx.x], in[threadIdx.x + 1]);
  __syncthreads();
  double a[NC], b[NC], c[NC];
  int k[NC];
#pragma unroll
  for (int i = 0; i < NC; i++) { a[i] = in[i] + threadIdx.x; b[i] = in[NC + i] * 0.999 + threadIdx.x * 1e-9; c[i] = in[2 * NC + i] * 1e-3 + threadIdx.x * 1e-10; k[i] = threadIdx.x + i; }
  const double m = in[40] + threadIdx.x * 1e-11, cc = in[41] + threadIdx.x * 1e-12;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < NC; i++) {
        if (MODE == 0) a[i] = fma(a[i], m, cc);
        if (MODE == 1 || MODE == 7 || MODE == 8) a[i] = fma(b[i], c[i], a[i]);
        if (MODE == 2) a[i] = fma(m, c[i], a[i]);
        if (MODE == 3) a[i] = fma(b[i], b[i], a[i]);
        if (MODE == 4) a[i] = a[i] * b[i];
        if (MODE == 5) a[i] = a[i] + b[i];
        if (MODE == 6) a[i] = fma(b[i], m, a[i]);
        if (MODE == 7) k[i] = (k[i] ^ it) + u;
        if (MODE == 8 && (i & 3) == 0) { double2 v = sm[(it + i) & 63]; b[i] = v.x; }
        if (MODE == 9) { double t = fma(b[i], c[i], a[i]); a[i] = b[i]; b[i] = t; }
      }
    }
  }
  double s = 0; int ks = 0;
#pragma unroll
  for (int i = 0; i < NC; i++) { s += a[i] + b[i] + c[i]; ks += k[i]; }
  if (s == -1.0 || ks == 123456789) out[0] = s + ks;
}

template <int MODE> void run(const char* what, int sms, const double* in, double* out, int threads, int blocks_per_sm) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 20000, blocks = sms * blocks_per_sm;
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    probe<MODE><<<blocks, threads>>>(iters, in, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double n = (double)iters * 4 * NC * blocks * threads;
  printf("probe %-44s thr=%3d blk/SM=%d  %8.3f ms  %6.2f FP64-instr-lanes/clk/SM (@1.965GHz)  %.2f Tinstr-flop/s\n", what, threads, blocks_per_sm, best,
         n / (best * 1e-3) / sms / 1.965e9, 2 * n / (best * 1e-3) * 1e-12);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 0.01 * i;
  double *in, *out; CK(cudaMalloc(&in, sizeof(h) + 16)); CK(cudaMalloc(&out, 8));
  CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
  for (int cfg = 0; cfg < 3; cfg++) {
    const int threads = cfg == 0 ? 256 : cfg == 1 ? 128 : 256, bps = cfg == 0 ? 4 : cfg == 1 ? 1 : 2;
    run<0>("fma(a,m,c) shared m,c", sms, in, out, threads, bps);
    run<1>("fma(b_i,c_i,a_i) 3 distinct", sms, in, out, threads, bps);
    run<2>("fma(m,c_i,a_i) slot-A shared", sms, in, out, threads, bps);
    run<6>("fma(b_i,m,a_i) slot-B shared", sms, in, out, threads, bps);
    run<3>("fma(b_i,b_i,a_i) 2 distinct", sms, in, out, threads, bps);
    run<4>("mul(a_i,b_i)", sms, in, out, threads, bps);
    run<5>("add(a_i,b_i)", sms, in, out, threads, bps);
    run<7>("3 distinct + 2 int ops per DFMA", sms, in, out, threads, bps);
    run<8>("3 distinct + LDS.128 per 4 DFMA", sms, in, out, threads, bps);
    run<9>("fma(b,c,a) rotating dest", sms, in, out, threads, bps);
  }
  return 0;
}
