// Device twins of the reference micro-kernels and of approx_rsqrt<digits>.
//
// Reference semantics (paths relative to /root/reference/include/sctl/):
//   kernel_impl::Laplace3D_FxU/DxU/FxdU, Stokes3D_FxU/DxU      kernel_functions.hpp:32-218
//   approx_rsqrt<digits>(r2, r2 > 0): seed + Newton, 0 on masked lanes   vec.txx:375-382, intrin-wrapper.hpp:540-559
//   pair update dX = xt - xs; fused uKerApply or uKerMatrix + FMA          generic-kernel.txx:107-121
//
// This is not a translation of those bodies: the arithmetic is re-associated for the B200 FP64 pipe
// (fewest DADD/DMUL/DFMA per pair; everything that is not FP64 goes to the ALU/XU pipes that issue in the
// FP64 pipe's shadow), and sources arrive as packed records (see source_pack.cuh) instead of three AoS arrays.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// 1: FP32 pair loops run on packed values (f32x2 below: two targets per 64-bit register pair); 0: scalar loops
// (tools/ab_build.sh builds both to measure the difference)
#ifndef SCTL_CUDA_F32X2
#define SCTL_CUDA_F32X2 1
#endif

namespace sctl_cuda {

// 0-4: the north-star kernels.  5-11: the other non-Helmholtz kernels of kernel_functions.hpp.
enum : int {
  LAPLACE3D_FXU = 0, LAPLACE3D_DXU = 1, LAPLACE3D_FXDU = 2, STOKES3D_FXU = 3, STOKES3D_DXU = 4,
  STOKES3D_FXT = 5, STOKES3D_FSXU = 6, STOKES3D_FXUP = 7, LAPLACE3D_FXD2U = 8, LAPLACE3D_DXDU = 9, BIOTSAVART3D_FXU = 10, BIOTSAVART3D_FXDU = 11,
  // 12-15: the Helmholtz kernels; their context (ctx_ptr) is the wavenumber, passed to the device by value
  HELMHOLTZ3D_FXU = 12, HELMHOLTZ3D_DXU = 13, HELMHOLTZ3D_FXDU = 14, HELMHOLTZDIFF3D_FXDU = 15,
  NUM_KERNELS = 16
};
template <int KID> struct KerNeedsCtx { static constexpr bool value = (KID >= HELMHOLTZ3D_FXU); };

// How far the seed is refined; chosen on the host from `digits` (see rsq_mode_for_digits()).
enum : int { RSQ_SEED = 0, RSQ_NEWTON = 1, RSQ_CUBIC = 2 };

// ------------------------------------------------------------------------------------------------
// Two FP32 values in one 64-bit register pair (FP32 pair loops, two targets per value)
// ------------------------------------------------------------------------------------------------
// sm_100a has packed FP32 instructions, FADD2 / FMUL2 / FFMA2: one issue slot, two IEEE operations, each half rounded
// exactly like the scalar instruction, so a loop written on f32x2 gives the same bits as the scalar loop.  The FP32
// pair loops are issue-bound (one FFMA per cycle per SM sub-partition leaves no slot for the loads, MUFU and loop
// control), and the packed form halves the FP32 issue count.  What makes it usable here: a packed instruction takes a
// SCALAR 32-bit operand broadcast to both halves (SASS `FFMA2 R4, R6.F32x2.HI_LO, UR9.F32, R4.F32x2.HI_LO`, also
// `R9.F32`), so one source value -- from a uniform register on the constant-bank path, from an LDS result on the ring
// path -- serves the two targets of a value without being duplicated (ptxas folds f32x2(s) below into that operand).
// The pair bodies (PairApply) are written once on `Real` and instantiate for float, double and f32x2.
struct f32x2 {
  float2 v;
  __device__ __forceinline__ f32x2() {}
  __device__ __forceinline__ f32x2(float s) : v(make_float2(s, s)) {}  // broadcast (implicit: source values, literals)
  __device__ __forceinline__ f32x2(float lo, float hi) : v(make_float2(lo, hi)) {}
  __device__ __forceinline__ explicit f32x2(float2 p) : v(p) {}
};
__device__ __forceinline__ f32x2 operator-(f32x2 a) { return f32x2(-a.v.x, -a.v.y); }
__device__ __forceinline__ f32x2 operator+(f32x2 a, f32x2 b) { return f32x2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ f32x2 operator-(f32x2 a, f32x2 b) { return f32x2(__fadd2_rn(a.v, (-b).v)); }
__device__ __forceinline__ f32x2 operator*(f32x2 a, f32x2 b) { return f32x2(__fmul2_rn(a.v, b.v)); }
using ::fma;  // the overload below must not hide fma(double, double, double) / fma(float, float, float) in this namespace
__device__ __forceinline__ f32x2 fma(f32x2 a, f32x2 b, f32x2 c) { return f32x2(__ffma2_rn(a.v, b.v, c.v)); }
__device__ __forceinline__ f32x2& operator+=(f32x2& a, f32x2 b) { return a = a + b; }
__device__ __forceinline__ f32x2& operator-=(f32x2& a, f32x2 b) { return a = a - b; }
__device__ __forceinline__ f32x2& operator*=(f32x2& a, f32x2 b) { return a = a * b; }

// ------------------------------------------------------------------------------------------------
// Masked reciprocal square root:  r2 > 0 ? r2^(-1/2) : 0     (approx_rsqrt<digits>(r2, r2 > 0))
// ------------------------------------------------------------------------------------------------

// FP64.  Seed = MUFU.RSQ64H on the high word (one XU-pipe instruction, valid over the whole normal
// double range, ~2^-20 relative error because the low mantissa word is ignored).  Refinement runs on
// the FP64 pipe with the residual e = 1 - x*y0^2; the exact factor is (1 - e)^(-1/2) = 1 + e/2 + 3/8 e^2 + ...:
//   RSQ_NEWTON: y = y0 * (A + B e)                    4 FP64 instr, |error| < 9e-13, zero mean
//   RSQ_CUBIC : y = y0 * (1 + e*(1/2 + 3/8 e))        5 FP64 instr, error ~ 5/16 e^3 (~1 ulp)
// SCTL's AVX-512 recipe (14-bit seed, 2 Newton steps, deferred scaling) costs 7.
// RSQ_NEWTON's line is not the Taylor line 1 + e/2 (error 3/8 e^2: always positive, up to 1.3e-12) but the tangent of
// the quadratic at the seed's MEAN residual, shifted so that the error has zero mean over the seed's measured residual
// distribution (tools/seed_error.py on B200: e in [-1.13e-6, 1.85e-6], mean m = 2.65e-7, variance v = 2.3e-13):
//   B = 1/2 + 3/4 m,  A = 1 + 3/8 (v - m^2)   =>   error = 3/8 ((e - m)^2 - v)  in [-8.6e-14, 8.6e-13].
// Same instruction count; sums of many pairs no longer carry a systematic 1.1e-13 bias (profiles/r2_seed_error.md).
constexpr double RSQ_NEWTON_B = 0.5 + 0.75 * 2.65e-7;
constexpr double RSQ_NEWTON_A = 1.0 + 0.375 * (2.3e-13 - 2.65e-7 * 2.65e-7);
// Mask: like _mm512_maskz_rsqrt14_pd in the reference (intrin-wrapper.hpp:3403) the SEED is zeroed and the
// refinement then keeps it zero (e = 1, y0 * (1 + ...) = 0).  Subnormal r2 (|dx| < 1.5e-154) is treated as
// coincident; the reference overflows to Inf/NaN there for every kernel but Laplace FxU.
// MASK = false drops the compare + select.  It is bit-identical to MASK = true whenever r2 >= 2^-1022 (or NaN/Inf),
// and the host only selects it after coincidence_probe_kernel has shown that no (source, target) pair can have a
// smaller r2 (eval_kernel.cuh); the two instructions cost ~3 % of the Stokes pair (profiles/r1_whatif.txt).
// `lo` and `k` are the two loop-invariant operands of the refinement: the seed's low word (always 0: MUFU.RSQ64H only
// writes the high word) and the polynomial coefficient that cannot be an immediate because the same DFMA already
// carries one (3/8 for RSQ_CUBIC, B for RSQ_NEWTON).  Given as literals (the one-argument overload below) ptxas
// re-materialises both inside the pair loop: one IMAD.MOV per pair for the low word, two per iteration for the
// coefficient.  A pair loop that hands in values ptxas cannot fold (kernel arguments: eval_const_kernel.cuh) keeps the
// coefficient in one register pair and lets MUFU.RSQ64H write the high half of a pair whose low half stays zero, so
// none of those moves is left in the loop.  The arithmetic, and so every result bit, is the same.
template <int MODE> __host__ __device__ constexpr double rsq_poly_k() { return MODE == RSQ_NEWTON ? RSQ_NEWTON_B : 0.375; }
template <int MODE, bool MASK = true> __device__ __forceinline__ double rsqrt_masked(double x, int lo, double k) {
  // One signed compare on the high word: false for zero, subnormal and negative inputs, then a select on the seed's high
  // word; its low word is zero anyway.  NaN / Inf inputs are NOT masked and propagate as NaN.  (The reference's r2 > 0
  // compare is false for NaN, so there a NaN pair is dropped by kernels that only multiply by rinv, e.g. Laplace3D_FxU,
  // and still gives NaN wherever the kernel also multiplies by the displacement; inputs with NaN are outside the parity
  // contract and the unmasked loop could not reproduce that distinction anyway.)  Every
  // register-reading instruction competes with the DFMA stream for register-file bandwidth on B200
  // (tools/probe.cu, probe4), so the mask is kept to compare + select.
  double ys;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(ys) : "d"(x));
  const int hs = __double2hiint(ys);
  const double y0 = __hiloint2double(MASK ? (__double2hiint(x) >= 0x00100000 ? hs : 0) : hs, lo);
  if (MODE == RSQ_SEED) return y0;
  const double e = fma(-x, y0 * y0, 1.0);
  // Written as y0 * (1 + e*p) rather than y0 + (y0*e)*p: every instruction then reads at most two distinct
  // vector registers.  A DFMA with three distinct register sources issues every 3 cycles instead of 2 on B200.
  if (MODE == RSQ_NEWTON) return y0 * fma(e, k, RSQ_NEWTON_A);
  return y0 * fma(e, fma(e, k, 0.5), 1.0);
}
template <int MODE, bool MASK = true> __device__ __forceinline__ double rsqrt_masked(double x) {
  return rsqrt_masked<MODE, MASK>(x, 0, rsq_poly_k<MODE>());
}

// FP32.  Seed = MUFU.RSQ (2^-22 relative error): that already meets approx_rsqrt<digits<=6>, which is what the
// reference computes for float by default (digits = (int)(23*0.30103) = 6, vec.txx:376).  digits >= 7 adds one
// Newton step on the FP32 pipe.  Same masking rule as FP64 (subnormal r2 means |dx| < 1.1e-19).
template <int MODE, bool MASK = true> __device__ __forceinline__ float rsqrt_masked(float x, int = 0, float = 0) {
  float ys;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ys) : "f"(x));
  const float y0 = MASK ? (__float_as_int(x) >= 0x00800000 ? ys : 0.0f) : ys;
  if (MODE == RSQ_SEED) return y0;
  const float e = fmaf(-x, y0 * y0, 1.0f);
  return y0 * fmaf(e, 0.5f, 1.0f);
}
// Two FP32 values: one MUFU.RSQ (and one compare + select) per half, the Newton step packed.  Same operations per half as
// the float overload, hence the same bits.
template <int MODE, bool MASK = true> __device__ __forceinline__ f32x2 rsqrt_masked(f32x2 x, int = 0, f32x2 = f32x2()) {
  const f32x2 y0(rsqrt_masked<RSQ_SEED, MASK>(x.v.x), rsqrt_masked<RSQ_SEED, MASK>(x.v.y));
  if (MODE == RSQ_SEED) return y0;
  const f32x2 e = fma(-x, y0 * y0, f32x2(1.0f));
  return y0 * fma(e, f32x2(0.5f), f32x2(1.0f));
}

// ------------------------------------------------------------------------------------------------
// Kernel traits
// ------------------------------------------------------------------------------------------------
// K0/K1/ND: GenericKernel's KDIM0/KDIM1/N_DIM (generic-kernel.hpp:22-41).  FLOPS: uKernel::FLOPS().
// PACK: reals per packed source record before padding (source_pack.cuh).  F_ALG: SURVEY.md section 8(d).
template <int KID> struct KerTraits;
template <> struct KerTraits<LAPLACE3D_FXU>  { static constexpr int K0 = 1, K1 = 1, ND = 0, PACK = 4, FLOPS = 6,  F_ALG = 11; };
template <> struct KerTraits<LAPLACE3D_DXU>  { static constexpr int K0 = 1, K1 = 1, ND = 3, PACK = 6, FLOPS = 14, F_ALG = 19; };
template <> struct KerTraits<LAPLACE3D_FXDU> { static constexpr int K0 = 1, K1 = 3, ND = 0, PACK = 4, FLOPS = 11, F_ALG = 18; };
template <> struct KerTraits<STOKES3D_FXU>   { static constexpr int K0 = 3, K1 = 3, ND = 0, PACK = 6, FLOPS = 23, F_ALG = 29; };
template <> struct KerTraits<STOKES3D_DXU>   { static constexpr int K0 = 3, K1 = 3, ND = 3, PACK = 9, FLOPS = 26, F_ALG = 30; };
// kernel_functions.hpp:220-483.  F_ALG here = flop count of the fused device formulation below (FMA = 2, rsqrt = 1).
template <> struct KerTraits<STOKES3D_FXT>      { static constexpr int K0 = 3, K1 = 9, ND = 0, PACK = 6, FLOPS = 39, F_ALG = 39; };
template <> struct KerTraits<STOKES3D_FSXU>     { static constexpr int K0 = 4, K1 = 3, ND = 0, PACK = 7, FLOPS = 26, F_ALG = 30; };
template <> struct KerTraits<STOKES3D_FXUP>     { static constexpr int K0 = 3, K1 = 4, ND = 0, PACK = 6, FLOPS = 26, F_ALG = 31; };
template <> struct KerTraits<LAPLACE3D_FXD2U>   { static constexpr int K0 = 1, K1 = 9, ND = 0, PACK = 4, FLOPS = 25, F_ALG = 38; };
template <> struct KerTraits<LAPLACE3D_DXDU>    { static constexpr int K0 = 1, K1 = 3, ND = 3, PACK = 6, FLOPS = 22, F_ALG = 30; };
template <> struct KerTraits<BIOTSAVART3D_FXU>  { static constexpr int K0 = 3, K1 = 3, ND = 0, PACK = 6, FLOPS = 15, F_ALG = 26; };
template <> struct KerTraits<BIOTSAVART3D_FXDU> { static constexpr int K0 = 3, K1 = 9, ND = 0, PACK = 6, FLOPS = 51, F_ALG = 52; };
// kernel_functions.hpp:490-659; complex values as adjacent (re, im) pairs; sincos counted as 2 flop in F_ALG
template <> struct KerTraits<HELMHOLTZ3D_FXU>      { static constexpr int K0 = 2, K1 = 2, ND = 0, PACK = 5, FLOPS = 20, F_ALG = 24; };
template <> struct KerTraits<HELMHOLTZ3D_DXU>      { static constexpr int K0 = 2, K1 = 2, ND = 3, PACK = 8, FLOPS = 29, F_ALG = 38; };
template <> struct KerTraits<HELMHOLTZ3D_FXDU>     { static constexpr int K0 = 2, K1 = 6, ND = 0, PACK = 5, FLOPS = 35, F_ALG = 44; };
template <> struct KerTraits<HELMHOLTZDIFF3D_FXDU> { static constexpr int K0 = 2, K1 = 6, ND = 0, PACK = 5, FLOPS = 41, F_ALG = 50; };

// Reals per record, padded so that every record is a whole number of 16-byte words (LDS.128 / bulk-copy unit).
template <class Real, int KID> struct RecTraits {
  static constexpr int W = 16 / (int)sizeof(Real);
  static constexpr int REC = (KerTraits<KID>::PACK + W - 1) / W * W;
};

struct KernelInfo { int k0, k1, nd, flops, f_alg; const char* name; };
__host__ inline KernelInfo kernel_info(int kid) {
  switch (kid) {
    case LAPLACE3D_FXU:  return {1, 1, 0, 6, 11, "Laplace3D-FxU"};
    case LAPLACE3D_DXU:  return {1, 1, 3, 14, 19, "Laplace3D-DxU"};
    case LAPLACE3D_FXDU: return {1, 3, 0, 11, 18, "Laplace3D-FxdU"};
    case STOKES3D_FXU:   return {3, 3, 0, 23, 29, "Stokes3D-FxU"};
    case STOKES3D_DXU:   return {3, 3, 3, 26, 30, "Stokes3D-DxU"};
    case STOKES3D_FXT:      return {3, 9, 0, 39, 39, "Stokes3D-FxT"};
    case STOKES3D_FSXU:     return {4, 3, 0, 26, 30, "Stokes3D-FSxU"};
    case STOKES3D_FXUP:     return {3, 4, 0, 26, 31, "Stokes3D-FxUP"};
    case LAPLACE3D_FXD2U:   return {1, 9, 0, 25, 38, "Laplace3D-Fxd2U"};
    case LAPLACE3D_DXDU:    return {1, 3, 3, 22, 30, "Laplace3D-DxdU"};
    case BIOTSAVART3D_FXU:  return {3, 3, 0, 15, 26, "BiotSavart3D-FxU"};
    case BIOTSAVART3D_FXDU: return {3, 9, 0, 51, 52, "BiotSavart3D-FxdU"};
    case HELMHOLTZ3D_FXU:      return {2, 2, 0, 20, 24, "Helmholtz3D-FxU"};
    case HELMHOLTZ3D_DXU:      return {2, 2, 3, 29, 38, "Helmholtz3D-DxU"};
    case HELMHOLTZ3D_FXDU:     return {2, 6, 0, 35, 44, "Helmholtz3D-FxdU"};
    case HELMHOLTZDIFF3D_FXDU: return {2, 6, 0, 41, 50, "HelmholtzDiff3D-FxdU"};
    default:             return {0, 0, 0, 0, 0, nullptr};
  }
}

// uKerScaleFactor<Real>() (kernel_functions.hpp:40-42, 69-71, 89-91, 133-135, 188-190), evaluated in Real.
template <class Real> __host__ __device__ inline Real scale_factor(int kid) {
  const Real pi = (Real)3.14159265358979323846264338327950288L;
  switch (kid) {
    case LAPLACE3D_FXU: case LAPLACE3D_DXU: case LAPLACE3D_FXD2U: case LAPLACE3D_DXDU: case BIOTSAVART3D_FXU:
    case HELMHOLTZ3D_FXU: case HELMHOLTZ3D_FXDU: case HELMHOLTZDIFF3D_FXDU: return 1 / (4 * pi);
    case LAPLACE3D_FXDU: case BIOTSAVART3D_FXDU: case HELMHOLTZ3D_DXU: return -1 / (4 * pi);
    case STOKES3D_FXU: case STOKES3D_FSXU: case STOKES3D_FXUP: return 1 / (8 * pi);
    case STOKES3D_FXT: return -3 / (4 * pi);
    default: return 3 / (4 * pi);  // Stokes3D_DxU
  }
}

// ------------------------------------------------------------------------------------------------
// Fused pair update:  acc[K1] += K(dx; n) f   (unscaled), given dx = xt - xs and rinv.
// `s` points at the packed source record held in registers: s[0..2] = xs, s[3..] = payload.
// FP64-pipe instruction counts per pair (after the 6-instruction prefix dx, r2 and the rsqrt):
//   Laplace FxU 1 | Laplace DxU 6 | Laplace FxdU 6 | Stokes FxU 11 | Stokes DxU 14
// ------------------------------------------------------------------------------------------------
template <int KID, class Real> struct PairApply;

// u = rinv                                   (kernel_functions.hpp:43-47)
template <class Real> struct PairApply<LAPLACE3D_FXU, Real> {
  template <class S> static __device__ __forceinline__ void run(Real (&acc)[1], const Real (&dx)[3], Real rinv, const S* s) {
    acc[0] = fma(rinv, s[3], acc[0]);
  }
};
// u = (r.n) rinv^3; the record carries g = f*n so the pair costs one dot product      (:72-78)
template <class Real> struct PairApply<LAPLACE3D_DXU, Real> {
  template <class S> static __device__ __forceinline__ void run(Real (&acc)[1], const Real (&dx)[3], Real rinv, const S* s) {
    const Real rinv3 = rinv * rinv * rinv;
    const Real d = fma(dx[2], s[5], fma(dx[1], s[4], dx[0] * s[3]));
    acc[0] = fma(d, rinv3, acc[0]);
  }
};
// v_i += r_i (f rinv^3)                      (fused apply, :102-110)
template <class Real> struct PairApply<LAPLACE3D_FXDU, Real> {
  template <class S> static __device__ __forceinline__ void run(Real (&acc)[3], const Real (&dx)[3], Real rinv, const S* s) {
    const Real t = (rinv * rinv) * (rinv * s[3]);
    acc[0] = fma(dx[0], t, acc[0]);
    acc[1] = fma(dx[1], t, acc[1]);
    acc[2] = fma(dx[2], t, acc[2]);
  }
};
// v_i += f_i rinv + r_i (f.r) rinv^3  =  rinv (f_i + r_i (f.r) rinv^2)      (fused apply, :166-176)
template <class Real> struct PairApply<STOKES3D_FXU, Real> {
  template <class S> static __device__ __forceinline__ void run(Real (&acc)[3], const Real (&dx)[3], Real rinv, const S* s) {
    const Real rinv2 = rinv * rinv;
    const Real d = fma(dx[2], s[5], fma(dx[1], s[4], dx[0] * s[3]));
    const Real c = d * rinv2;
    acc[0] = fma(rinv, fma(dx[0], c, s[3]), acc[0]);
    acc[1] = fma(rinv, fma(dx[1], c, s[4]), acc[1]);
    acc[2] = fma(rinv, fma(dx[2], c, s[5]), acc[2]);
  }
};
// v_i += r_i (f.r)(r.n) rinv^5              (fused apply, :205-217)
template <class Real> struct PairApply<STOKES3D_DXU, Real> {
  template <class S> static __device__ __forceinline__ void run(Real (&acc)[3], const Real (&dx)[3], Real rinv, const S* s) {
    const Real rinv2 = rinv * rinv;
    const Real rinv5 = (rinv2 * rinv2) * rinv;
    const Real df = fma(dx[2], s[5], fma(dx[1], s[4], dx[0] * s[3]));
    const Real dn = fma(dx[2], s[8], fma(dx[1], s[7], dx[0] * s[6]));
    const Real t = (df * dn) * rinv5;
    acc[0] = fma(dx[0], t, acc[0]);
    acc[1] = fma(dx[1], t, acc[1]);
    acc[2] = fma(dx[2], t, acc[2]);
  }
};


// ---- the other non-Helmholtz kernels (kernel_functions.hpp:220-483) ----
// Stokes3D_FxT: v_ij += r_i r_j (f.r) rinv^5                    (fused apply, :241-252)
template <class Real> struct PairApply<STOKES3D_FXT, Real> {
  template <class S> static __device__ __forceinline__ void run(Real (&acc)[9], const Real (&dx)[3], Real rinv, const S* s) {
    const Real rinv2 = rinv * rinv;
    const Real rinv5 = (rinv2 * rinv2) * rinv;
    const Real t = fma(dx[2], s[5], fma(dx[1], s[4], dx[0] * s[3])) * rinv5;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const Real a = dx[i] * t;
#pragma unroll
      for (int j = 0; j < 3; j++) acc[i * 3 + j] = fma(a, dx[j], acc[i * 3 + j]);
    }
  }
};
// Stokes3D_FSxU: v_i += f_i rinv + r_i (f.r + f_3) rinv^3        (fused apply, :311-321)
template <class Real> struct PairApply<STOKES3D_FSXU, Real> {
  template <class S> static __device__ __forceinline__ void run(Real (&acc)[3], const Real (&dx)[3], Real rinv, const S* s) {
    const Real rinv2 = rinv * rinv;
    // one source value per instruction: a DFMA can take a single uniform-register operand (constant-bank path)
    const Real c = (fma(dx[2], s[5], fma(dx[1], s[4], dx[0] * s[3])) + s[6]) * rinv2;
    acc[0] = fma(rinv, fma(dx[0], c, s[3]), acc[0]);
    acc[1] = fma(rinv, fma(dx[1], c, s[4]), acc[1]);
    acc[2] = fma(rinv, fma(dx[2], c, s[5]), acc[2]);
  }
};
// Stokes3D_FxUP: velocity as Stokes3D_FxU, pressure v_3 += (f.r) rinv^3   (fused apply, :351-362)
template <class Real> struct PairApply<STOKES3D_FXUP, Real> {
  template <class S> static __device__ __forceinline__ void run(Real (&acc)[4], const Real (&dx)[3], Real rinv, const S* s) {
    const Real rinv2 = rinv * rinv;
    const Real c = fma(dx[2], s[5], fma(dx[1], s[4], dx[0] * s[3])) * rinv2;
    acc[0] = fma(rinv, fma(dx[0], c, s[3]), acc[0]);
    acc[1] = fma(rinv, fma(dx[1], c, s[4]), acc[1]);
    acc[2] = fma(rinv, fma(dx[2], c, s[5]), acc[2]);
    acc[3] = fma(rinv, c, acc[3]);
  }
};
// Laplace3D_Fxd2U: v_ij += f (3 r_i r_j rinv^5 - delta_ij rinv^3)        (uKerMatrix + FMA, :376-388)
template <class Real> struct PairApply<LAPLACE3D_FXD2U, Real> {
  template <class S> static __device__ __forceinline__ void run(Real (&acc)[9], const Real (&dx)[3], Real rinv, const S* s) {
    const Real rinv2 = rinv * rinv;
    const Real g = (rinv2 * rinv) * s[3];
    const Real h = (g * rinv2) * (Real)3;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const Real a = dx[i] * h;
#pragma unroll
      for (int j = 0; j < 3; j++) acc[i * 3 + j] = fma(a, dx[j], acc[i * 3 + j]);
      acc[i * 3 + i] -= g;
    }
  }
};
// Laplace3D_DxdU: v_i += f (n_i rinv^3 - 3 r_i (r.n) rinv^5); the record carries g = f*n     (:402-411)
template <class Real> struct PairApply<LAPLACE3D_DXDU, Real> {
  template <class S> static __device__ __forceinline__ void run(Real (&acc)[3], const Real (&dx)[3], Real rinv, const S* s) {
    const Real rinv2 = rinv * rinv;
    const Real rinv3 = rinv2 * rinv;
    const Real c = (fma(dx[2], s[5], fma(dx[1], s[4], dx[0] * s[3])) * rinv2) * (Real)3;
    acc[0] = fma(rinv3, fma(-dx[0], c, s[3]), acc[0]);
    acc[1] = fma(rinv3, fma(-dx[1], c, s[4]), acc[1]);
    acc[2] = fma(rinv3, fma(-dx[2], c, s[5]), acc[2]);
  }
};
// BiotSavart3D_FxU: v += (f x r) rinv^3                                  (fused apply, :436-447)
template <class Real> struct PairApply<BIOTSAVART3D_FXU, Real> {
  template <class S> static __device__ __forceinline__ void run(Real (&acc)[3], const Real (&dx)[3], Real rinv, const S* s) {
    const Real rinv3 = rinv * rinv * rinv;
    acc[0] = fma(fma(s[4], dx[2], -(dx[1] * s[5])), rinv3, acc[0]);
    acc[1] = fma(fma(s[5], dx[0], -(dx[2] * s[3])), rinv3, acc[1]);
    acc[2] = fma(fma(s[3], dx[1], -(dx[0] * s[4])), rinv3, acc[2]);
  }
};
// BiotSavart3D_FxdU: v_ab += eps_abk f_k rinv^3 + 3 (f x r)_a r_b rinv^5   (uKerMatrix + FMA, :460-482; the 3 x 9
// block is the negated gradient of the FxU block, its sign lives in the scale factor)
template <class Real> struct PairApply<BIOTSAVART3D_FXDU, Real> {
  template <class S> static __device__ __forceinline__ void run(Real (&acc)[9], const Real (&dx)[3], Real rinv, const S* s) {
    const Real rinv2 = rinv * rinv;
    const Real rinv3 = rinv2 * rinv;
    const Real t5 = (rinv3 * rinv2) * (Real)3;
    Real c[3];
    c[0] = fma(s[4], dx[2], -(dx[1] * s[5]));
    c[1] = fma(s[5], dx[0], -(dx[2] * s[3]));
    c[2] = fma(s[3], dx[1], -(dx[0] * s[4]));
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const Real ca = c[a] * t5;
#pragma unroll
      for (int b = 0; b < 3; b++) acc[a * 3 + b] = fma(ca, dx[b], acc[a * 3 + b]);
    }
    const Real g0 = s[3] * rinv3, g1 = s[4] * rinv3, g2 = s[5] * rinv3;
    acc[1] += g2; acc[2] -= g1;
    acc[3] -= g2; acc[5] += g0;
    acc[6] += g1; acc[7] -= g0;
  }
};

// ---- Helmholtz kernels (kernel_functions.hpp:490-659): G = exp(i mu r) / r and derivatives; record = [x y z f_re f_im (n)] ----
// FP64 sin and cos of the same argument.  CUDA's sincos() spends ~21 FP64 and ~45 auxiliary instructions per call in
// these loops (conversions on the XU pipe, a slow-path call, coefficient moves); next to a saturated FP64 pipe the
// auxiliary ones cost almost as much as the arithmetic (profiles/r1_whatif_and_aux_cost.txt).  This version keeps the
// arithmetic -- round-to-nearest quadrant by the 1.5*2^52 trick, two-term Cody-Waite reduction with FMA, the fdlibm
// minimax polynomials on [-pi/4, pi/4] -- takes its coefficients straight from the constant bank, and selects the
// quadrant with integer ops on the high words.  Max abs error 1.3e-16 for |x| <= 1e5 (checked against long double on
// 2e7 points); larger arguments (mu*r > 1e5) take the library path.
__constant__ double c_sincos[16] = {
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, 2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11,
    6.36619772367581382433e-01 /* 2/pi */, -1.5707963267948966 /* -pi/2 hi */, -6.123233995736766e-17 /* -pi/2 lo */, 6755399441055744.0 /* 1.5 * 2^52 */};
__device__ __noinline__ void sincos_large(double x, double* s, double* c) { sincos(x, s, c); }  // out of line: keeps the loops small
__device__ __forceinline__ void sincos_real(double x, double* s, double* c) {
  if (!(fabs(x) <= 1.0e5)) {  // also NaN / Inf
    sincos_large(x, s, c);
    return;
  }
  const double* k = c_sincos;
  const double t = fma(x, k[12], k[15]);
  const int q = __double2loint(t);  // nearest integer to x * 2/pi, low bits
  const double n = t - k[15];
  double r = fma(n, k[13], x);
  r = fma(n, k[14], r);
  const double z = r * r;
  double ps = fma(z, k[5], k[4]);
  double pc = fma(z, k[11], k[10]);
  ps = fma(z, ps, k[3]);
  pc = fma(z, pc, k[9]);
  ps = fma(z, ps, k[2]);
  pc = fma(z, pc, k[8]);
  ps = fma(z, ps, k[1]);
  pc = fma(z, pc, k[7]);
  ps = fma(z, ps, k[0]);
  pc = fma(z, pc, k[6]);
  const double sr = fma(r * z, ps, r);
  const double cr = fma(z, fma(z, pc, -0.5), 1.0);
  const bool swap = q & 1;
  const double s0 = swap ? cr : sr, c0 = swap ? sr : cr;
  *s = __hiloint2double(__double2hiint(s0) ^ ((q & 2) << 30), __double2loint(s0));
  *c = __hiloint2double(__double2hiint(c0) ^ (((q + 1) & 2) << 30), __double2loint(c0));
}
// FP32 twin: three-term Cody-Waite reduction with FMA, cephes minimax polynomials; max abs error 7.1e-8 for
// |x| <= 4.5e4 (checked against double on 2e7 points); larger arguments take the library path.
__device__ __noinline__ void sincos_large(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_real(float x, float* s, float* c) {
  if (!(fabsf(x) <= 4.5e4f)) {
    sincos_large(x, s, c);
    return;
  }
  const float t = fmaf(x, 0.636619772367581f, 12582912.0f);  // 1.5 * 2^23: round to nearest integer
  const int q = __float_as_int(t);
  const float n = t - 12582912.0f;
  float r = fmaf(n, -1.57079637050628662109375f, x);
  r = fmaf(n, 4.37113882867379237e-8f, r);
  r = fmaf(n, 1.71512451e-15f, r);
  const float z = r * r;
  float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
  float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  ps = fmaf(z, ps, -1.6666654611e-1f);
  pc = fmaf(z, pc, 4.166664568298827e-2f);
  const float sr = fmaf(r * z, ps, r);
  const float cr = fmaf(z, fmaf(z, pc, -0.5f), 1.0f);
  const bool swap = q & 1;
  const float s0 = swap ? cr : sr, c0 = swap ? sr : cr;
  *s = __int_as_float(__float_as_int(s0) ^ ((q & 2) << 30));
  *c = __int_as_float(__float_as_int(c0) ^ (((q + 1) & 2) << 30));
}

// Two FP32 arguments at once: the reduction and the polynomials run on the packed FP32 instructions, the quadrant fix-up
// per half on the integer pipe.  Same operations per half as the float version (same bits); a value with either half out
// of range takes the scalar path for both.
__device__ __forceinline__ void sincos_real(f32x2 x, f32x2* s, f32x2* c) {
  if (!(fabsf(x.v.x) <= 4.5e4f && fabsf(x.v.y) <= 4.5e4f)) {
    sincos_real(x.v.x, &s->v.x, &c->v.x);
    sincos_real(x.v.y, &s->v.y, &c->v.y);
    return;
  }
  const f32x2 t = fma(x, f32x2(0.636619772367581f), f32x2(12582912.0f));
  const f32x2 n = t - f32x2(12582912.0f);
  f32x2 r = fma(n, f32x2(-1.57079637050628662109375f), x);
  r = fma(n, f32x2(4.37113882867379237e-8f), r);
  r = fma(n, f32x2(1.71512451e-15f), r);
  const f32x2 z = r * r;
  f32x2 ps = fma(z, f32x2(-1.9515295891e-4f), f32x2(8.3321608736e-3f));
  f32x2 pc = fma(z, f32x2(2.443315711809948e-5f), f32x2(-1.388731625493765e-3f));
  ps = fma(z, ps, f32x2(-1.6666654611e-1f));
  pc = fma(z, pc, f32x2(4.166664568298827e-2f));
  const f32x2 sr = fma(r * z, ps, r);
  const f32x2 cr = fma(z, fma(z, pc, f32x2(-0.5f)), f32x2(1.0f));
  const int q0 = __float_as_int(t.v.x), q1 = __float_as_int(t.v.y);
  const float s0 = (q0 & 1) ? cr.v.x : sr.v.x, c0 = (q0 & 1) ? sr.v.x : cr.v.x;
  const float s1 = (q1 & 1) ? cr.v.y : sr.v.y, c1 = (q1 & 1) ? sr.v.y : cr.v.y;
  *s = f32x2(__int_as_float(__float_as_int(s0) ^ ((q0 & 2) << 30)), __int_as_float(__float_as_int(s1) ^ ((q1 & 2) << 30)));
  *c = f32x2(__int_as_float(__float_as_int(c0) ^ (((q0 + 1) & 2) << 30)), __int_as_float(__float_as_int(c1) ^ (((q1 + 1) & 2) << 30)));
}

// (G0, G1) of the gradient-type kernels: (-mu sin - cos/r, mu cos - sin/r) / r^2;  DIFF subtracts the Laplace part
// through the half-angle identity cos(x) - 1 = -2 sin^2(x/2) (:640-647)
template <bool DIFF, class Real> __device__ __forceinline__ void helmholtz_grad_G(Real& G0, Real& G1, Real r2, Real rinv, Real mu) {
  const Real rinv2 = rinv * rinv;
  Real sn, cs;
  if (DIFF) {
    Real sh, ch;
    sincos_real((r2 * rinv) * mu * (Real)0.5, &sh, &ch);
    sn = (Real)2 * sh * ch;
    const Real cm1 = (Real)-2 * sh * sh;
    cs = (Real)1 + cm1;
    G0 = (-mu * sn - cm1 * rinv) * rinv2;
  } else {
    sincos_real((r2 * rinv) * mu, &sn, &cs);
    G0 = (-mu * sn - cs * rinv) * rinv2;
  }
  G1 = (mu * cs - sn * rinv) * rinv2;
}
template <class Real> struct PairApply<HELMHOLTZ3D_FXU, Real> {  // :519-532
  template <class S> static __device__ __forceinline__ void run_ctx(Real (&acc)[2], const Real (&dx)[3], Real rinv, const S* s, Real r2, Real mu) {
    Real sn, cs;
    sincos_real((r2 * rinv) * mu, &sn, &cs);
    const Real G0 = cs * rinv, G1 = sn * rinv;
    acc[0] = fma(s[3], G0, fma(-s[4], G1, acc[0]));
    acc[1] = fma(s[3], G1, fma(s[4], G0, acc[1]));
  }
};
template <class Real> struct PairApply<HELMHOLTZ3D_DXU, Real> {  // :561-577
  template <class S> static __device__ __forceinline__ void run_ctx(Real (&acc)[2], const Real (&dx)[3], Real rinv, const S* s, Real r2, Real mu) {
    Real G0, G1;
    helmholtz_grad_G<false>(G0, G1, r2, rinv, mu);
    const Real ndotr = fma(dx[2], s[7], fma(dx[1], s[6], dx[0] * s[5]));
    G0 *= ndotr;
    G1 *= ndotr;
    acc[0] = fma(s[3], G0, fma(-s[4], G1, acc[0]));
    acc[1] = fma(s[3], G1, fma(s[4], G0, acc[1]));
  }
};
template <bool DIFF, class Real, class S> __device__ __forceinline__ void helmholtz_fxdu_apply(Real (&acc)[6], const Real (&dx)[3], Real rinv, const S* s, Real r2, Real mu) {
  Real G0, G1;
  helmholtz_grad_G<DIFF>(G0, G1, r2, rinv, mu);
  const Real fG0 = fma(s[3], G0, -(s[4] * G1));
  const Real fG1 = fma(s[3], G1, s[4] * G0);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    acc[i * 2 + 0] = fma(fG0, dx[i], acc[i * 2 + 0]);
    acc[i * 2 + 1] = fma(fG1, dx[i], acc[i * 2 + 1]);
  }
}
template <class Real> struct PairApply<HELMHOLTZ3D_FXDU, Real> {  // :607-626
  template <class S> static __device__ __forceinline__ void run_ctx(Real (&acc)[6], const Real (&dx)[3], Real rinv, const S* s, Real r2, Real mu) {
    helmholtz_fxdu_apply<false>(acc, dx, rinv, s, r2, mu);
  }
};
template <class Real> struct PairApply<HELMHOLTZDIFF3D_FXDU, Real> {  // uKerMatrix + FMA, :629-657
  template <class S> static __device__ __forceinline__ void run_ctx(Real (&acc)[6], const Real (&dx)[3], Real rinv, const S* s, Real r2, Real mu) {
    helmholtz_fxdu_apply<true>(acc, dx, rinv, s, r2, mu);
  }
};

// One source-target pair, starting from coordinates.
// `mu` is the kernel context (the Helmholtz wavenumber); the other kernels ignore it.
// `lo`, `k`: the refinement's loop-invariant operands (rsqrt_masked above); callers that do not hoist them leave the
// defaults, which are the literals.
// Real is the type the pair is computed in (double, float, or f32x2 = two targets at once); S the scalar type of the
// record (Real itself, or float for f32x2).
template <int KID, int RSQ, bool MASK = true, class Real, class S>
__device__ __forceinline__ void pair_update(Real (&acc)[KerTraits<KID>::K1], const Real (&xt)[3], const S* s, Real mu = Real(0), int lo = 0,
                                            Real k = Real((S)rsq_poly_k<RSQ>())) {
  Real dx[3];
  dx[0] = xt[0] - s[0];
  dx[1] = xt[1] - s[1];
  dx[2] = xt[2] - s[2];
  const Real r2 = fma(dx[2], dx[2], fma(dx[1], dx[1], dx[0] * dx[0]));
  const Real rinv = rsqrt_masked<RSQ, MASK>(r2, lo, k);
  if constexpr (KerNeedsCtx<KID>::value) PairApply<KID, Real>::run_ctx(acc, dx, rinv, s, r2, mu);
  else PairApply<KID, Real>::run(acc, dx, rinv, s);
}

// The same pair in two halves, for loops that software-pipeline across sources (eval_const_kernel.cuh): the front half
// ends at the seed (everything up to the MUFU), the back half refines it and applies the kernel.  pair_front + pair_back
// perform exactly the operations of pair_update, in the same order per pair.
template <int RSQ, bool MASK, class S>
__device__ __forceinline__ void pair_front(double (&dx)[3], double& r2, double& y0, const double (&xt)[3], const S* s) {
  dx[0] = xt[0] - s[0];
  dx[1] = xt[1] - s[1];
  dx[2] = xt[2] - s[2];
  r2 = fma(dx[2], dx[2], fma(dx[1], dx[1], dx[0] * dx[0]));
  y0 = rsqrt_masked<RSQ_SEED, MASK>(r2, 0, 0.0);
}
template <int KID, int RSQ, class S>
__device__ __forceinline__ void pair_back(double (&acc)[KerTraits<KID>::K1], const double (&dx)[3], double r2, double y0, const S* s) {
  double rinv = y0;
  if (RSQ != RSQ_SEED) {
    const double e = fma(-r2, y0 * y0, 1.0);
    rinv = (RSQ == RSQ_NEWTON) ? y0 * fma(e, RSQ_NEWTON_B, RSQ_NEWTON_A) : y0 * fma(e, fma(e, 0.375, 0.5), 1.0);
  }
  PairApply<KID, double>::run(acc, dx, rinv, s);
}

// uKerMatrix twin: u[K0][K1] (unscaled) for one displacement, with the raw normal (KernelMatrix path,
// generic-kernel.txx:213-329, never uses the fused apply).
template <int KID, class Real> struct PairMatrix;
template <class Real> struct PairMatrix<LAPLACE3D_FXU, Real> {
  static __device__ __forceinline__ void run(Real* u, const Real (&dx)[3], Real rinv, const Real* n) { u[0] = rinv; }
};
template <class Real> struct PairMatrix<LAPLACE3D_DXU, Real> {
  static __device__ __forceinline__ void run(Real* u, const Real (&dx)[3], Real rinv, const Real* n) {
    const Real rdotn = fma(dx[2], n[2], fma(dx[1], n[1], dx[0] * n[0]));
    u[0] = rdotn * (rinv * rinv * rinv);
  }
};
template <class Real> struct PairMatrix<LAPLACE3D_FXDU, Real> {
  static __device__ __forceinline__ void run(Real* u, const Real (&dx)[3], Real rinv, const Real* n) {
    const Real rinv3 = rinv * rinv * rinv;
    u[0] = dx[0] * rinv3; u[1] = dx[1] * rinv3; u[2] = dx[2] * rinv3;
  }
};
template <class Real> struct PairMatrix<STOKES3D_FXU, Real> {
  static __device__ __forceinline__ void run(Real* u, const Real (&dx)[3], Real rinv, const Real* n) {
    const Real rinv3 = rinv * rinv * rinv;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const Real a = dx[i] * rinv3;
#pragma unroll
      for (int j = 0; j < 3; j++) u[i * 3 + j] = (i == j) ? fma(a, dx[j], rinv) : a * dx[j];
    }
  }
};
template <class Real> struct PairMatrix<STOKES3D_DXU, Real> {
  static __device__ __forceinline__ void run(Real* u, const Real (&dx)[3], Real rinv, const Real* n) {
    const Real rinv2 = rinv * rinv;
    const Real rinv5 = (rinv2 * rinv2) * rinv;
    const Real c = fma(dx[2], n[2], fma(dx[1], n[1], dx[0] * n[0])) * rinv5;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const Real a = dx[i] * c;
#pragma unroll
      for (int j = 0; j < 3; j++) u[i * 3 + j] = a * dx[j];
    }
  }
};


template <class Real> struct PairMatrix<STOKES3D_FXT, Real> {
  static __device__ __forceinline__ void run(Real* u, const Real (&dx)[3], Real rinv, const Real* n) {
    const Real rinv2 = rinv * rinv;
    const Real rinv5 = (rinv2 * rinv2) * rinv;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const Real a = (dx[i] * dx[j]) * rinv5;
#pragma unroll
        for (int k = 0; k < 3; k++) u[i * 9 + j * 3 + k] = a * dx[k];
      }
  }
};
template <class Real> struct PairMatrix<STOKES3D_FSXU, Real> {
  static __device__ __forceinline__ void run(Real* u, const Real (&dx)[3], Real rinv, const Real* n) {
    const Real rinv3 = rinv * rinv * rinv;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const Real a = dx[i] * rinv3;
#pragma unroll
      for (int j = 0; j < 3; j++) u[i * 3 + j] = (i == j) ? fma(a, dx[j], rinv) : a * dx[j];
      u[9 + i] = a;
    }
  }
};
template <class Real> struct PairMatrix<STOKES3D_FXUP, Real> {
  static __device__ __forceinline__ void run(Real* u, const Real (&dx)[3], Real rinv, const Real* n) {
    const Real rinv3 = rinv * rinv * rinv;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const Real a = dx[i] * rinv3;
#pragma unroll
      for (int j = 0; j < 3; j++) u[i * 4 + j] = (i == j) ? fma(a, dx[j], rinv) : a * dx[j];
      u[i * 4 + 3] = a;
    }
  }
};
template <class Real> struct PairMatrix<LAPLACE3D_FXD2U, Real> {
  static __device__ __forceinline__ void run(Real* u, const Real (&dx)[3], Real rinv, const Real* n) {
    const Real rinv2 = rinv * rinv;
    const Real rinv3 = rinv2 * rinv;
    const Real t5 = (rinv3 * rinv2) * (Real)3;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const Real a = dx[i] * t5;
#pragma unroll
      for (int j = 0; j < 3; j++) u[i * 3 + j] = (i == j) ? fma(a, dx[j], -rinv3) : a * dx[j];
    }
  }
};
template <class Real> struct PairMatrix<LAPLACE3D_DXDU, Real> {
  static __device__ __forceinline__ void run(Real* u, const Real (&dx)[3], Real rinv, const Real* n) {
    const Real rinv2 = rinv * rinv;
    const Real rinv3 = rinv2 * rinv;
    const Real t5 = ((fma(dx[2], n[2], fma(dx[1], n[1], dx[0] * n[0])) * rinv2) * rinv3) * (Real)3;
#pragma unroll
    for (int i = 0; i < 3; i++) u[i] = fma(-dx[i], t5, n[i] * rinv3);
  }
};
template <class Real> struct PairMatrix<BIOTSAVART3D_FXU, Real> {
  static __device__ __forceinline__ void run(Real* u, const Real (&dx)[3], Real rinv, const Real* n) {
    const Real rinv3 = rinv * rinv * rinv;
    const Real a0 = dx[0] * rinv3, a1 = dx[1] * rinv3, a2 = dx[2] * rinv3;
    u[0] = 0;   u[1] = -a2; u[2] = a1;
    u[3] = a2;  u[4] = 0;   u[5] = -a0;
    u[6] = -a1; u[7] = a0;  u[8] = 0;
  }
};
template <class Real> struct PairMatrix<BIOTSAVART3D_FXDU, Real> {
  static __device__ __forceinline__ void run(Real* u, const Real (&dx)[3], Real rinv, const Real* n) {
    const Real rinv2 = rinv * rinv;
    const Real rinv3 = rinv2 * rinv;
    const Real t5 = (rinv3 * rinv2) * (Real)3;
    // u[k0][3a+b] = -eps_{a k0 b} rinv3 + t5 (eps_{a k0 c} r_c) r_b
#pragma unroll
    for (int k0 = 0; k0 < 3; k0++)
#pragma unroll
      for (int a = 0; a < 3; a++) {
        // e_c = eps_{a k0 c} r_c : nonzero only for c = 3 - a - k0 when a != k0
        const int c = 3 - a - k0;
        const bool nz = (a != k0);
        const Real sgn = nz ? ((((k0 - a + 3) % 3) == 1) ? (Real)1 : (Real)-1) : (Real)0;  // eps_{a k0 c}
        const Real e = nz ? sgn * dx[c < 3 ? c : 0] : (Real)0;
#pragma unroll
        for (int b = 0; b < 3; b++) {
          Real val = (t5 * e) * dx[b];
          if (nz && b == c) val -= sgn * rinv3;  // -eps_{a k0 b} rinv3
          u[k0 * 9 + a * 3 + b] = val;
        }
      }
  }
};

// Helmholtz uKerMatrix twins: 2 x 2K complex-as-real blocks, u[1][*] is the rotated copy of u[0][*]
template <class Real> struct PairMatrix<HELMHOLTZ3D_FXU, Real> {
  static __device__ __forceinline__ void run_ctx(Real* u, const Real (&dx)[3], Real rinv, const Real* n, Real r2, Real mu) {
    Real sn, cs;
    sincos_real((r2 * rinv) * mu, &sn, &cs);
    u[0] = cs * rinv; u[1] = sn * rinv;
    u[2] = -u[1];     u[3] = u[0];
  }
};
template <class Real> struct PairMatrix<HELMHOLTZ3D_DXU, Real> {
  static __device__ __forceinline__ void run_ctx(Real* u, const Real (&dx)[3], Real rinv, const Real* n, Real r2, Real mu) {
    Real G0, G1;
    helmholtz_grad_G<false>(G0, G1, r2, rinv, mu);
    const Real ndotr = fma(dx[2], n[2], fma(dx[1], n[1], dx[0] * n[0]));
    u[0] = G0 * ndotr; u[1] = G1 * ndotr;
    u[2] = -u[1];      u[3] = u[0];
  }
};
template <bool DIFF, class Real> __device__ __forceinline__ void helmholtz_fxdu_matrix(Real* u, const Real (&dx)[3], Real rinv, Real r2, Real mu) {
  Real G0, G1;
  helmholtz_grad_G<DIFF>(G0, G1, r2, rinv, mu);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    u[i * 2 + 0] = G0 * dx[i];         u[i * 2 + 1] = G1 * dx[i];
    u[6 + i * 2 + 0] = -u[i * 2 + 1];  u[6 + i * 2 + 1] = u[i * 2 + 0];
  }
}
template <class Real> struct PairMatrix<HELMHOLTZ3D_FXDU, Real> {
  static __device__ __forceinline__ void run_ctx(Real* u, const Real (&dx)[3], Real rinv, const Real* n, Real r2, Real mu) { helmholtz_fxdu_matrix<false>(u, dx, rinv, r2, mu); }
};
template <class Real> struct PairMatrix<HELMHOLTZDIFF3D_FXDU, Real> {
  static __device__ __forceinline__ void run_ctx(Real* u, const Real (&dx)[3], Real rinv, const Real* n, Real r2, Real mu) { helmholtz_fxdu_matrix<true>(u, dx, rinv, r2, mu); }
};

}  // namespace sctl_cuda
