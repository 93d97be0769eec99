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

// FP32 operand probes: same modes as probe<> (0 shared m,c | 1 three distinct | 2 slot-A shared | 3 two distinct | 4 mul | 5 add | 6 slot-B shared)
template <int MODE> __global__ void __launch_bounds__(256) probef(int iters, const double* __restrict__ in, double* out) {
  constexpr int N = 12;
  float a[N], b[N], c[N];
#pragma unroll
  for (int i = 0; i < N; i++) { a[i] = (float)in[i] + threadIdx.x; b[i] = (float)in[N + i] * 0.999f + threadIdx.x * 1e-6f; c[i] = (float)in[2 * N + i] * 1e-3f + threadIdx.x * 1e-7f; }
  const float m = (float)in[40] + threadIdx.x * 1e-6f, cc = (float)in[41] + threadIdx.x * 1e-7f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < N; i++) {
        if (MODE == 0) a[i] = fmaf(a[i], m, cc);
        if (MODE == 1) a[i] = fmaf(b[i], c[i], a[i]);
        if (MODE == 2) a[i] = fmaf(m, c[i], a[i]);
        if (MODE == 3) a[i] = fmaf(b[i], b[i], a[i]);
        if (MODE == 4) a[i] = a[i] * b[i];
        if (MODE == 5) a[i] = a[i] + b[i];
        if (MODE == 6) a[i] = fmaf(b[i], m, a[i]);
        if (MODE == 7) a[i] = fmaf(a[i], 0.999f, 1e-6f);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < N; i++) s += a[i] + b[i] + c[i];
  if (s == -1.0f) out[0] = s;
}
template <int MODE> void runf(const char* what, int sms, const double* in, double* out, int threads, int blocks_per_sm) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 40000, blocks = sms * blocks_per_sm;
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    probef<MODE><<<blocks, threads>>>(iters, in, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double n = (double)iters * 4 * 12 * blocks * threads;
  printf("probef %-40s thr=%3d blk/SM=%d  %8.3f ms  %7.2f FP32-instr-lanes/clk/SM (@1.965GHz)\n", what, threads, blocks_per_sm, best, n / (best * 1e-3) / sms / 1.965e9);
}

// What does each auxiliary instruction of the pair loop cost the FP64 pipe?  24 two-operand DFMA per iteration plus
// EXTRA: 0 nothing | 1 one MUFU.RSQ64H | 2 one MUFU.RSQ (f32) | 3 four ALU ops (add/compare/select/mov) | 4 two LDCU.64 | 5 all
__constant__ double cprobe[1024];
template <int EXTRA> __global__ void __launch_bounds__(256) probe4(int iters, const double* __restrict__ in, double* out) {
  double a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = in[i] + threadIdx.x; b[i] = in[8 + i] * 1e-3 + threadIdx.x * 1e-9; }
  double x = 2.0 + threadIdx.x;
  float xf = 2.0f + threadIdx.x;
  int k = threadIdx.x;
  double cs = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 3; u++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = fma(b[i], b[i], a[i]);
    if (EXTRA == 1 || EXTRA == 5) {
      double y;
      asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
      x = __hiloint2double(__double2hiint(x) ^ (__double2hiint(y) & 1), __double2loint(x));
    }
    if (EXTRA == 2) {
      float y;
      asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(xf));
      xf = __uint_as_float(__float_as_uint(xf) ^ (__float_as_uint(y) & 1));
    }
    if (EXTRA == 3 || EXTRA == 5) {
      int t = k + 0x00100000;
      k = ((unsigned)t < 0x7fe00000u) ? (t ^ it) : 0;
    }
    if (EXTRA == 4 || EXTRA == 5) cs += cprobe[(it * 2) & 1023] + cprobe[(it * 2 + 1) & 1023];
  }
  double s = x + xf + k + cs;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i];
  if (s == -1.0) out[0] = s;
}
template <int EXTRA> void run4(const char* what, int sms, const double* in, double* out) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 40000, blocks = sms * 2, threads = 256;
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    probe4<EXTRA><<<blocks, threads>>>(iters, in, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double n = (double)iters * 24 * blocks * threads;
  printf("probe4 24 DFMA + %-40s %8.3f ms  %6.2f DFMA-lanes/clk/SM (@1.965GHz)\n", what, best, n / (best * 1e-3) / sms / 1.965e9);
}

// Triples that share one multiplicand (the accumulate pattern acc_i = fma(rinv, w_i, acc_i)):
// MODE 0: rinv in slot A; MODE 1: rinv in slot B; MODE 2: slot A plus a 2-operand DMUL between groups
template <int MODE> __global__ void __launch_bounds__(256) probe3(int iters, const double* __restrict__ in, double* out) {
  constexpr int G = 4;
  double a[3 * G], w[3 * G], r[G];
#pragma unroll
  for (int i = 0; i < 3 * G; i++) { a[i] = in[i] + threadIdx.x; w[i] = in[12 + i] * 1e-3 + threadIdx.x * 1e-9; }
#pragma unroll
  for (int g = 0; g < G; g++) r[g] = in[30 + g] * 0.999 + threadIdx.x * 1e-10;
  const double m = in[40] + threadIdx.x * 1e-11;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int g = 0; g < G; g++) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
          if (MODE == 1) a[3 * g + i] = fma(w[3 * g + i], r[g], a[3 * g + i]);
          else a[3 * g + i] = fma(r[g], w[3 * g + i], a[3 * g + i]);
        }
        if (MODE == 2) r[g] = r[g] * m;
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 3 * G; i++) s += a[i] + w[i];
#pragma unroll
  for (int g = 0; g < G; g++) s += r[g];
  if (s == -1.0) out[0] = s;
}

template <int MODE> void run3(const char* what, int sms, const double* in, double* out, int threads, int blocks_per_sm) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 20000, blocks = sms * blocks_per_sm;
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    probe3<MODE><<<blocks, threads>>>(iters, in, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double n = (double)iters * 4 * (MODE == 2 ? 16 : 12) * blocks * threads;
  printf("probe3 %-43s thr=%3d blk/SM=%d  %8.3f ms  %6.2f FP64-instr-lanes/clk/SM (@1.965GHz)\n", what, threads, blocks_per_sm, best, n / (best * 1e-3) / sms / 1.965e9);
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
  for (int bps : {2, 4}) {
    runf<7>("ffma(a, imm, imm)", sms, in, out, 256, bps);
    runf<0>("ffma(a, m, c) shared m,c", sms, in, out, 256, bps);
    runf<1>("ffma(b_i, c_i, a_i) 3 distinct", sms, in, out, 256, bps);
    runf<2>("ffma(m, c_i, a_i) slot-A shared", sms, in, out, 256, bps);
    runf<6>("ffma(b_i, m, a_i) slot-B shared", sms, in, out, 256, bps);
    runf<3>("ffma(b_i, b_i, a_i) 2 distinct", sms, in, out, 256, bps);
    runf<4>("fmul(a_i, b_i)", sms, in, out, 256, bps);
    runf<5>("fadd(a_i, b_i)", sms, in, out, 256, bps);
  }
  run4<0>("nothing", sms, in, out);
  run4<1>("1 MUFU.RSQ64H", sms, in, out);
  run4<2>("1 MUFU.RSQ (f32)", sms, in, out);
  run4<3>("4 ALU ops", sms, in, out);
  run4<4>("2 LDCU.64 (+1 DADD)", sms, in, out);
  run4<5>("MUFU.RSQ64H + 4 ALU + 2 LDCU", sms, in, out);
  for (int cfg = 0; cfg < 4; cfg++) {
    const int threads = 128, bps = cfg == 0 ? 1 : cfg == 1 ? 2 : cfg == 2 ? 4 : 8;  // 1, 2, 4, 8 warps per sub-partition
    run3<0>("triples, shared operand in slot A", sms, in, out, threads, bps);
    run3<1>("triples, shared operand in slot B", sms, in, out, threads, bps);
    run3<2>("triples (slot A) + DMUL between groups", sms, in, out, threads, bps);
  }
  return 0;
}
