// tools/sweep.cu -- tuning harness (not part of the product): times eval_pairs_kernel variants on one B200 and
// prints pairs/s, fraction of the DFMA peak measured in the same process, and the error against a reference variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I sctl_b200/csrc tools/sweep.cu -o gpurun_out/sweep
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "eval_kernel.cuh"
#include "eval_const_kernel.cuh"

using namespace sctl_cuda;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);       \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

template <class Real> __global__ void fill_kernel(Real* p, int64_t n, uint64_t seed, Real off) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = (seed << 32) + (uint64_t)i + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    p[i] = (Real)((double)(z >> 11) * 0x1.0p-53) + off;
  }
}

template <class Real> __global__ void __launch_bounds__(256) peak_kernel(int iters, Real* out, Real seed) {
  Real a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const Real m = (Real)0.999999, c = (Real)1e-6;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const Real s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == (Real)-1) out[0] = s;
}

// XU-pipe probes: how fast are MUFU.RSQ64H / MUFU.RSQ / F2F alone and next to a saturating DFMA stream?
template <int MODE> __global__ void __launch_bounds__(256) xu_probe_kernel(int iters, double* out) {
  double a0 = 1.0 + threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double x0 = 2.0 + threadIdx.x, x1 = x0 + 1;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (MODE == 0 || MODE >= 2) {  // 8 DFMA
        a0 = fma(a0, 0.999999, 1e-6); a1 = fma(a1, 0.999999, 1e-6); a2 = fma(a2, 0.999999, 1e-6); a3 = fma(a3, 0.999999, 1e-6);
        a0 = fma(a0, 0.999999, 1e-6); a1 = fma(a1, 0.999999, 1e-6); a2 = fma(a2, 0.999999, 1e-6); a3 = fma(a3, 0.999999, 1e-6);
      }
      if (MODE == 1 || MODE == 2) {  // 2 MUFU.RSQ64H (integer feedback keeps them live without FP64 work)
        double y0, y1;
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x0));
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y1) : "d"(x1));
        x0 = __hiloint2double(__double2hiint(x0) ^ (__double2hiint(y0) & 1), __double2loint(x0));
        x1 = __hiloint2double(__double2hiint(x1) ^ (__double2hiint(y1) & 1), __double2loint(x1));
      }
      if (MODE == 3) {  // 8 DFMA + 2x (F2F.F32.F64, MUFU.RSQ, F2F.F64.F32)
        float f0 = (float)x0, f1 = (float)x1, r0, r1;
        asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(f0));
        asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(f1));
        double y0 = (double)r0, y1 = (double)r1;
        x0 = __hiloint2double(__double2hiint(x0) ^ (__double2hiint(y0) & 1), __double2loint(x0));
        x1 = __hiloint2double(__double2hiint(x1) ^ (__double2hiint(y1) & 1), __double2loint(x1));
      }
    }
  }
  const double s = (a0 + a1) + (a2 + a3) + x0 + x1;
  if (s == -1.0) out[0] = s;
}

struct Ctx {
  int sms;
  double peak64, peak32;
  int64_t Ns, Nt;
  void *r_src, *n_src, *v_src, *r_trg, *v_trg, *v_ref, *pack, *partial;
  cudaEvent_t e0, e1;
};

template <class Real> double run_peak(Ctx& c) {
  Real* out;
  CK(cudaMalloc(&out, sizeof(Real)));
  const int blocks = c.sms * 8, iters = 40000;
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(c.e0));
    peak_kernel<Real><<<blocks, 256>>>(iters, out, (Real)0.5);
    CK(cudaEventRecord(c.e1));
    CK(cudaEventSynchronize(c.e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, c.e0, c.e1));
    best = std::max(best, 2.0 * 64 * iters * (double)blocks * 256 / (ms * 1e-3) * 1e-12);
  }
  CK(cudaFree(out));
  return best;
}

template <int MODE> void run_xu(Ctx& c, const char* what, double dfma_per_iter, double xu_per_iter) {
  double* out;
  CK(cudaMalloc(&out, 8));
  const int blocks = c.sms * 8, iters = 20000;
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(c.e0));
    xu_probe_kernel<MODE><<<blocks, 256>>>(iters, out);
    CK(cudaEventRecord(c.e1));
    CK(cudaEventSynchronize(c.e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, c.e0, c.e1));
    best = std::min(best, ms);
  }
  const double n = (double)iters * 8 * blocks * 256;
  printf("xu_probe %-46s %8.3f ms  DFMA %.2f TFLOP/s  XU-ops %.1f G/s (%.2f /clk/SM @1.965GHz)\n", what, best, 2 * dfma_per_iter * n / (best * 1e-3) * 1e-12,
         xu_per_iter * n / (best * 1e-3) * 1e-9, xu_per_iter * n / (best * 1e-3) / c.sms / 1.965e9);
  CK(cudaFree(out));
}

template <class Real, int KID, int T, int NW, int R, int RSQ, int TILE, int STAGES, int MINB, int U>
void run_variant(Ctx& c, bool is_ref, int force_split = 0) {
  using Cfg = EvalConfig<Real, KID, T, NW, R, RSQ, TILE, STAGES>;
  auto kern = eval_pairs_kernel<Real, KID, T, NW, R, RSQ, TILE, STAGES, MINB, U>;
  constexpr int K1 = KerTraits<KID>::K1;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
  int nb = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::THREADS, Cfg::SMEM_BYTES));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  EvalPlan plan = plan_eval(c.Ns, c.Nt, Cfg::TPB, c.sms * nb, R > 1 ? TILE : 8 * TILE, R > 1 ? 1024 : 64);
  if (force_split > 0) {
    plan.n_split = force_split;
    plan.chunk = (c.Ns + force_split - 1) / force_split;
  }
  EvalArgs<Real> a;
  a.pack = (const Real*)c.pack;
  a.r_trg = (const Real*)c.r_trg;
  a.v_trg = (Real*)(is_ref ? c.v_ref : c.v_trg);
  a.partial = (Real*)c.partial;
  a.Ns = c.Ns;
  a.Nt = c.Nt;
  a.chunk = plan.chunk;
  a.n_split = plan.n_split;
  a.scale = scale_factor<Real>(KID);
  dim3 grid((unsigned)plan.n_tblocks, (unsigned)plan.n_split);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaMemsetAsync(a.v_trg, 0, sizeof(Real) * c.Nt * K1));
    CK(cudaEventRecord(c.e0));
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES>>>(a);
    if (plan.n_split > 1) reduce_partials_kernel<Real><<<1024, 256>>>(c.Nt * K1, plan.n_split, a.partial, a.scale, a.v_trg);
    CK(cudaEventRecord(c.e1));
    CK(cudaEventSynchronize(c.e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, c.e0, c.e1));
    if (rep > 0) best = std::min(best, ms);
  }
  // error vs the reference variant
  std::vector<Real> h((size_t)c.Nt * K1), href((size_t)c.Nt * K1);
  CK(cudaMemcpy(h.data(), a.v_trg, sizeof(Real) * c.Nt * K1, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(href.data(), c.v_ref, sizeof(Real) * c.Nt * K1, cudaMemcpyDeviceToHost));
  double vmax = 0, dmax = 0;
  for (size_t i = 0; i < h.size(); i++) {
    vmax = std::max(vmax, std::fabs((double)href[i]));
    dmax = std::max(dmax, std::fabs((double)h[i] - (double)href[i]));
  }
  const double pairs = (double)c.Ns * c.Nt;
  const double rate = pairs / (best * 1e-3);
  const double peak = sizeof(Real) == 8 ? c.peak64 : c.peak32;
  printf("%s kid=%d T=%d NW=%d R=%d RSQ=%d TILE=%4d ST=%d MINB=%d U=%d | regs=%3d occ=%d smem=%6zu grid=%5ux%-3u eff=%.3f | %8.3f ms %7.1f Gpair/s  F_alg-frac=%.3f  err=%.2e\n",
         sizeof(Real) == 8 ? "f64" : "f32", KID, T, NW, R, RSQ, TILE, STAGES, MINB, U, fa.numRegs, nb, Cfg::SMEM_BYTES, grid.x, grid.y, plan.wave_eff, best, rate * 1e-9,
         rate * KerTraits<KID>::F_ALG * 1e-12 / peak, vmax > 0 ? dmax / vmax : 0.0);
  fflush(stdout);
}

// constant-bank variant: time the whole host loop (copy chunk to the constant bank + launch) over all Ns sources
template <class Real, int KID, int T, int NT, int RSQ, int U, int MINB>
void run_const(Ctx& c, int n_sub = 1, int wpc = 0) {
  if (wpc == 0) wpc = NT / 32;
  constexpr int REC = RecTraits<Real, KID>::REC;
  constexpr int K1 = KerTraits<KID>::K1;
  constexpr int CH = const_chunk_records<Real, KID>();
  auto kern = eval_pairs_const_kernel<Real, KID, T, NT, RSQ, U, MINB>;
  int nb = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, NT, 0));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  ConstEvalArgs<Real> a = {};
  a.r_trg = (const Real*)c.r_trg;
  a.v_trg = (Real*)c.v_trg;
  a.Nt = c.Nt;
  a.scale = scale_factor<Real>(KID);
  a.rsq_k = (Real)rsq_poly_k<RSQ>();
  const int tpb = 32 * wpc * T;
  dim3 grid((unsigned)((c.Nt + tpb - 1) / tpb), (unsigned)n_sub);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(32 * wpc);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = n_sub;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = n_sub > 1 ? 1 : 0;
  float best = 1e30f;
  int launches = 0;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaMemsetAsync(a.v_trg, 0, sizeof(Real) * c.Nt * K1));
    CK(cudaEventRecord(c.e0));
    launches = 0;
    for (int64_t s0 = 0; s0 < c.Ns; s0 += CH) {
      const int n = (int)std::min<int64_t>(CH, c.Ns - s0);
      CK(cudaMemcpyToSymbolAsync(c_src_pack, (const Real*)c.pack + s0 * REC, sizeof(Real) * n * REC, 0, cudaMemcpyDeviceToDevice));
      a.n_rec = n;
      a.sub_len = (n + n_sub - 1) / n_sub;
      a.full_ranks = n / a.sub_len;
      a.tail_len = n % a.sub_len;
      CK(cudaLaunchKernelEx(&cfg, kern, a));
      launches++;
    }
    CK(cudaEventRecord(c.e1));
    CK(cudaEventSynchronize(c.e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, c.e0, c.e1));
    if (rep > 0) best = std::min(best, ms);
  }
  std::vector<Real> h((size_t)c.Nt * K1), href((size_t)c.Nt * K1);
  CK(cudaMemcpy(h.data(), a.v_trg, sizeof(Real) * c.Nt * K1, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(href.data(), c.v_ref, sizeof(Real) * c.Nt * K1, cudaMemcpyDeviceToHost));
  double vmax = 0, dmax = 0;
  for (size_t i = 0; i < h.size(); i++) {
    vmax = std::max(vmax, std::fabs((double)href[i]));
    dmax = std::max(dmax, std::fabs((double)h[i] - (double)href[i]));
  }
  const double rate = (double)c.Ns * c.Nt / (best * 1e-3);
  const double peak = sizeof(Real) == 8 ? c.peak64 : c.peak32;
  printf("%s CONST kid=%d T=%d NT=%d wpc=%d RSQ=%d U=%d MINB=%d CL=%d | regs=%3d occ=%d CH=%d launches=%d grid=%ux%u | %8.3f ms %7.1f Gpair/s  F_alg-frac=%.3f  err=%.2e\n",
         sizeof(Real) == 8 ? "f64" : "f32", KID, T, NT, wpc, RSQ, U, MINB, n_sub, fa.numRegs, nb, CH, launches, grid.x, grid.y, best, rate * 1e-9,
         rate * KerTraits<KID>::F_ALG * 1e-12 / peak, vmax > 0 ? dmax / vmax : 0.0);
  fflush(stdout);
}

// Where does the per-launch cost of the constant-bank path go?  Same launches, three ways.
template <class Real, int KID, int T, int NT, int RSQ, int U, int MINB>
void run_const_breakdown(Ctx& c) {
  constexpr int REC = RecTraits<Real, KID>::REC;
  constexpr int CH = const_chunk_records<Real, KID>();
  auto kern = eval_pairs_const_kernel<Real, KID, T, NT, RSQ, U, MINB>;
  ConstEvalArgs<Real> a = {};
  a.r_trg = (const Real*)c.r_trg;
  a.v_trg = (Real*)c.v_trg;
  a.Nt = c.Nt;
  a.scale = scale_factor<Real>(KID);
  a.rsq_k = (Real)rsq_poly_k<RSQ>();
  a.n_rec = CH;
  a.sub_len = CH;
  a.full_ranks = 1;
  a.tail_len = 0;
  const int tpb = NT * T;
  dim3 grid((unsigned)((c.Nt + tpb - 1) / tpb), 1);
  const int n_launch = (int)((c.Ns + CH - 1) / CH);
  for (int mode = 0; mode < 4; mode++) {
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaEventRecord(c.e0));
      for (int i = 0; i < n_launch; i++) {
        if (mode == 1 || mode == 2) CK(cudaMemcpyToSymbolAsync(c_src_pack, (const Real*)c.pack + (int64_t)(i % 64) * CH * REC, sizeof(Real) * CH * REC, 0, cudaMemcpyDeviceToDevice));
        if (mode == 3) CK(cudaMemcpyAsync(c.partial, (const Real*)c.pack + (int64_t)(i % 64) * CH * REC, sizeof(Real) * CH * REC, cudaMemcpyDeviceToDevice));
        if (mode != 1) kern<<<grid, NT>>>(a);
      }
      CK(cudaEventRecord(c.e1));
      CK(cudaEventSynchronize(c.e1));
      CK(cudaGetLastError());
      float ms;
      CK(cudaEventElapsedTime(&ms, c.e0, c.e1));
      best = std::min(best, ms);
    }
    const char* what[4] = {"kernel only (constant bank untouched)", "memcpyToSymbol only", "memcpyToSymbol + kernel", "plain D2D memcpy (not to the bank) + kernel"};
    printf("breakdown Nt=%lld grid=%u x %d launches: %-46s %9.3f ms total, %8.2f us per launch\n", (long long)c.Nt, grid.x, n_launch, what[mode], best, best * 1e3 / n_launch);
  }
  fflush(stdout);
}

template <class Real, int KID> void setup(Ctx& c, int64_t Ns, int64_t Nt) {
  constexpr int REC = RecTraits<Real, KID>::REC;
  constexpr int K0 = KerTraits<KID>::K0, K1 = KerTraits<KID>::K1;
  c.Ns = Ns;
  c.Nt = Nt;
  CK(cudaMalloc(&c.r_src, sizeof(Real) * Ns * 3));
  CK(cudaMalloc(&c.n_src, sizeof(Real) * Ns * 3));
  CK(cudaMalloc(&c.v_src, sizeof(Real) * Ns * K0));
  CK(cudaMalloc(&c.r_trg, sizeof(Real) * Nt * 3));
  CK(cudaMalloc(&c.v_trg, sizeof(Real) * Nt * K1));
  CK(cudaMalloc(&c.v_ref, sizeof(Real) * Nt * K1));
  CK(cudaMalloc(&c.pack, sizeof(Real) * Ns * REC));
  CK(cudaMalloc(&c.partial, sizeof(Real) * Nt * K1 * 1024));
  fill_kernel<Real><<<1024, 256>>>((Real*)c.r_src, Ns * 3, 1, (Real)0);
  fill_kernel<Real><<<1024, 256>>>((Real*)c.r_trg, Nt * 3, 2, (Real)0);
  fill_kernel<Real><<<1024, 256>>>((Real*)c.n_src, Ns * 3, 3, (Real)-0.5);
  fill_kernel<Real><<<1024, 256>>>((Real*)c.v_src, Ns * K0, 4, (Real)-0.5);
  pack_sources_kernel<Real, KID><<<1024, 256>>>(Ns, (const Real*)c.r_src, (const Real*)c.n_src, (const Real*)c.v_src, (Real*)c.pack, (int64_t)Ns, (int64_t)0);
  CK(cudaDeviceSynchronize());
}
void teardown(Ctx& c) {
  for (void* p : {c.r_src, c.n_src, c.v_src, c.r_trg, c.v_trg, c.v_ref, c.pack, c.partial}) CK(cudaFree(p));
}

#ifndef SWEEP_LIST
#define SWEEP_LIST "sweep_list.inc"
#endif

int main(int argc, char** argv) {
  Ctx c{};
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  c.sms = prop.multiProcessorCount;
  CK(cudaEventCreate(&c.e0));
  CK(cudaEventCreate(&c.e1));
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, c.sms, prop.clockRate);
  c.peak64 = run_peak<double>(c);
  c.peak32 = run_peak<float>(c);
  printf("peak DFMA %.2f TFLOP/s, peak FFMA %.2f TFLOP/s  (lanes/SM/clk @1.965GHz: fp64 %.1f, fp32 %.1f)\n", c.peak64, c.peak32,
         c.peak64 * 1e12 / 2 / c.sms / 1.965e9, c.peak32 * 1e12 / 2 / c.sms / 1.965e9);
  run_xu<0>(c, "8 DFMA", 8, 0);
  run_xu<1>(c, "2 MUFU.RSQ64H", 0, 2);
  run_xu<2>(c, "8 DFMA + 2 MUFU.RSQ64H", 8, 2);
  run_xu<3>(c, "8 DFMA + 2x(F2F.F32.F64, MUFU.RSQ, F2F.F64.F32)", 8, 6);
  const int64_t Ns = argc > 1 ? atoll(argv[1]) : (1 << 18), Nt = argc > 2 ? atoll(argv[2]) : (1 << 17);
  printf("Ns=%lld Nt=%lld\n", (long long)Ns, (long long)Nt);
#include SWEEP_LIST
  c.peak64 = run_peak<double>(c);
  printf("peak DFMA after sweep %.2f TFLOP/s\n", c.peak64);
  return 0;
}
