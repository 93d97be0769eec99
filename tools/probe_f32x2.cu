// tools/probe_f32x2.cu -- does the packed FP32 pipe form (FFMA2 / FADD2 / FMUL2, sm_100a) free issue slots for the
// auxiliary instructions of the FP32 pair loop?  Tuning aid, not product code.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__constant__ float2 cpr[512];
__constant__ float cs[64];

// MODE 0: scalar FFMA, 3 distinct registers            (16 per iteration-unit)
// MODE 1: FFMA2, 3 distinct 64-bit registers           (8 per unit = same flops)
// MODE 2: FFMA2, 2 distinct (b, b, a)
// MODE 3: FFMA2 a = fma(a, m, c) shared m, c
// MODE 4: scalar FFMA x16 + 1 MUFU.RSQ + 4 ALU
// MODE 5: FFMA2 x8 + 1 MUFU.RSQ + 4 ALU
// MODE 6: FFMA2 with one operand from the constant bank (uniform address)
// MODE 7: FFMA2 x8 + 2 MUFU.RSQ + 8 ALU
// MODE 8: scalar FFMA x16 + 2 MUFU.RSQ + 8 ALU
// Round 2: the form the FP32 pair loops use -- one operand a SCALAR broadcast to both halves
// MODE 9:  FFMA2 a = fma(b, s, a), s a 32-bit value from the constant bank at a uniform address (SASS: UR.F32 operand)
// MODE 10: FFMA2 a = fma(b, s, a), s a 32-bit vector register (SASS: R.F32 operand)
// MODE 11: FFMA2 a = fma(a, c, s), the scalar in the addend slot, from the constant bank
template <int MODE> __global__ void __launch_bounds__(256) probe(int iters, const float* __restrict__ in, float* out) {
  float2 a[8], b[8], c[8];
  float r[2];
  int k[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = make_float2(in[i] + threadIdx.x, in[i + 1] - threadIdx.x);
    b[i] = make_float2(in[8 + i] * 0.999f + threadIdx.x * 1e-6f, in[9 + i] * 0.998f);
    c[i] = make_float2(in[16 + i] * 1e-3f + threadIdx.x * 1e-7f, in[17 + i] * 1e-3f);
    k[i] = threadIdx.x + i;
  }
  r[0] = in[30] + threadIdx.x; r[1] = in[31] + threadIdx.x;
  const float2 m = make_float2(in[40] + threadIdx.x * 1e-6f, in[41]), cc = make_float2(in[42] + threadIdx.x * 1e-7f, in[43]);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (MODE == 0 || MODE == 4 || MODE == 8) { a[i].x = fmaf(b[i].x, c[i].x, a[i].x); a[i].y = fmaf(b[i].y, c[i].y, a[i].y); }
        if (MODE == 1 || MODE == 5 || MODE == 7) a[i] = __ffma2_rn(b[i], c[i], a[i]);
        if (MODE == 2) a[i] = __ffma2_rn(b[i], b[i], a[i]);
        if (MODE == 3) a[i] = __ffma2_rn(a[i], m, cc);
        if (MODE == 6) a[i] = __ffma2_rn(b[i], cpr[(u * 8 + i)], a[i]);
        if (MODE == 9) { const float sc = cs[(it & 1) * 32 + u * 8 + i]; a[i] = __ffma2_rn(b[i], make_float2(sc, sc), a[i]); }
        if (MODE == 10) { const float sc = r[i & 1]; a[i] = __ffma2_rn(b[i], make_float2(sc, sc), a[i]); }
        if (MODE == 11) { const float sc = cs[(it & 1) * 32 + u * 8 + i]; a[i] = __ffma2_rn(a[i], c[i], make_float2(sc, sc)); }
      }
      if (MODE == 4 || MODE == 5 || MODE == 7 || MODE == 8) {
        constexpr int NA = (MODE >= 7) ? 2 : 1;
#pragma unroll
        for (int q = 0; q < NA; q++) {
          float y;
          asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(r[q]));
          r[q] = y + 1.5f;
#pragma unroll
          for (int i = 0; i < 4; i++) k[q * 4 + i] = (k[q * 4 + i] ^ it) + u;
        }
      }
    }
  }
  float s = r[0] + r[1]; int ks = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { s += a[i].x + a[i].y + b[i].x + c[i].y; ks += k[i]; }
  if (s == -1.0f || ks == 123456789) out[0] = s + ks;
}

template <int MODE> void run(const char* what, int sms, const float* in, float* out) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 20000, threads = 256, blocks = sms * 4;
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    probe<MODE><<<blocks, threads>>>(iters, in, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double n = (double)iters * 4 * 16 * blocks * threads;  // scalar-FMA equivalents
  printf("f32x2 %-52s %8.3f ms  %7.2f FMA-lanes/clk/SM (@1.965GHz)  %.2f TFLOP/s\n", what, best, n / (best * 1e-3) / sms / 1.965e9, 2 * n / (best * 1e-3) * 1e-12);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  float h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0f + 0.01f * i;
  float *in, *out; CK(cudaMalloc(&in, sizeof(h))); CK(cudaMalloc(&out, 64)); CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
  float2 hc[512]; for (int i = 0; i < 512; i++) hc[i] = make_float2(1e-3f * i, 2e-3f * i);
  CK(cudaMemcpyToSymbol(cpr, hc, sizeof(hc)));
  float hs[64]; for (int i = 0; i < 64; i++) hs[i] = 1e-3f * (i + 1);
  CK(cudaMemcpyToSymbol(cs, hs, sizeof(hs)));
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  run<0>("scalar FFMA 3 distinct", p.multiProcessorCount, in, out);
  run<1>("FFMA2 3 distinct", p.multiProcessorCount, in, out);
  run<2>("FFMA2 2 distinct", p.multiProcessorCount, in, out);
  run<3>("FFMA2 shared m,c", p.multiProcessorCount, in, out);
  run<6>("FFMA2 one operand from the constant bank", p.multiProcessorCount, in, out);
  run<9>("FFMA2 scalar operand UR.F32 (constant bank, uniform)", p.multiProcessorCount, in, out);
  run<10>("FFMA2 scalar operand R.F32 (vector register)", p.multiProcessorCount, in, out);
  run<11>("FFMA2 scalar addend UR.F32", p.multiProcessorCount, in, out);
  run<4>("scalar FFMA x16 + 1 MUFU + 4 ALU", p.multiProcessorCount, in, out);
  run<5>("FFMA2 x8 + 1 MUFU + 4 ALU", p.multiProcessorCount, in, out);
  run<8>("scalar FFMA x16 + 2 MUFU + 8 ALU", p.multiProcessorCount, in, out);
  run<7>("FFMA2 x8 + 2 MUFU + 8 ALU", p.multiProcessorCount, in, out);
  return 0;
}
