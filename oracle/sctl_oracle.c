/* oracle/sctl_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's (dmalhotra/SCTL) dense kernel-summation path:
 *   GenericKernel<uKernel>::Eval / KernelMatrix   include/sctl/generic-kernel.txx:102-329
 *   kernel_impl::Laplace3D_FxU/DxU/FxdU, Stokes3D_FxU/DxU   include/sctl/kernel_functions.hpp:32-218
 *   approx_rsqrt<digits> (masked)                 include/sctl/vec.txx:375-382, intrin-wrapper.hpp:540-579
 * It exists so the `-m gpu` parity tests, __graft_entry__.smoke() and bench.py's cpu_baseline leg have a
 * checker that travels to the GPU box.  Only tests/, smoke() and bench.py's CPU-baseline / --impl reference
 * legs may load it; the product (libsctl_cuda.so, include/, sctl_b200/) never does and has no CPU fallback.
 *
 * Parity pin: tests/test_oracle.py checks this file against (a) every analytic fixture the reference's own
 * tests hold for the path (src/test-generic-kernel.cpp:57-103, src/test-kernel_functions.cpp:36-131,488-526,
 * include/sctl/vec-test.hpp:431-465) and (b) golden vectors produced by running the UNMODIFIED reference
 * (oracle/_ref/libsctl_ref_*.so, built by oracle/Makefile from /root/reference/include; generator script
 * tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* (KDIM0, KDIM1, N_DIM) deduced by generic-kernel.hpp:22-41 from each uKerMatrix signature. */
#define NKER 16
static const int KDIMS[NKER][3] = {{1,1,0}, {1,1,3}, {1,3,0}, {3,3,0}, {3,3,3}, {3,9,0}, {4,3,0}, {3,4,0}, {1,9,0}, {1,3,3}, {3,3,0}, {3,9,0},
                                    {2,2,0}, {2,2,3}, {2,6,0}, {2,6,0}};
/* uKernel::FLOPS(): kernel_functions.hpp:37-39, 66-68, 86-88, 130-132, 185-187. */
static const int KFLOPS[NKER] = {6, 14, 11, 23, 26, 39, 26, 26, 25, 22, 15, 51, 20, 29, 35, 41};  /* + :225-227, 261-263, 331-333, 371-373, 396-398, 419-421, 455-457 */
static const char* KNAMES[NKER] = {"Laplace3D-FxU", "Laplace3D-DxU", "Laplace3D-FxdU", "Stokes3D-FxU", "Stokes3D-DxU", "Stokes3D-FxT", "Stokes3D-FSxU",
                                   "Stokes3D-FxUP", "Laplace3D-Fxd2U", "Laplace3D-DxdU", "BiotSavart3D-FxU", "BiotSavart3D-FxdU",
                                   "Helmholtz3D-FxU", "Helmholtz3D-DxU", "Helmholtz3D-FxdU", "HelmholtzDiff3D-FxdU"};

#define ORACLE_PI 3.14159265358979323846264338327950288L

/* uKerScaleFactor<Real>(): kernel_functions.hpp:40-42, 69-71, 89-91, 133-135, 188-190; evaluated in Real. */
static double scale_factor_f64(int kid) {
  const double pi = (double)ORACLE_PI;
  switch (kid) {
    case 0: case 1: case 8: case 9: case 10: case 12: case 14: case 15: return 1 / (4 * pi);
    case 2: case 11: case 13: return -1 / (4 * pi);  /* Helmholtz3D_DxU: kernel_functions.hpp:543-545 */
    case 3: case 6: case 7: return 1 / (8 * pi);
    case 5: return -3 / (4 * pi);  /* kernel_functions.hpp:228-230 */
    default: return 3 / (4 * pi);
  }
}
static float scale_factor_f32(int kid) {
  const float pi = (float)ORACLE_PI;
  switch (kid) {
    case 0: case 1: case 8: case 9: case 10: case 12: case 14: case 15: return 1 / (4 * pi);
    case 2: case 11: case 13: return -1 / (4 * pi);  /* Helmholtz3D_DxU: kernel_functions.hpp:543-545 */
    case 3: case 6: case 7: return 1 / (8 * pi);
    case 5: return -3 / (4 * pi);  /* kernel_functions.hpp:228-230 */
    default: return 3 / (4 * pi);
  }
}

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL double
#define SQRT sqrt
#define SIN sin
#define COS cos
#define FN(name) CAT(name, _f64)
#include "sctl_oracle_impl.h"
#undef REAL
#undef SQRT
#undef FN

#undef SIN
#undef COS
#define REAL float
#define SQRT sqrtf
#define SIN sinf
#define COS cosf
#define FN(name) CAT(name, _f32)
#include "sctl_oracle_impl.h"
#undef REAL
#undef SQRT
#undef FN

/* mu: kernel context (Helmholtz wavenumber, kernel_functions.hpp:484-488); ignored by the other kernels */
int oracle_eval(int kernel_id, int real_bytes, int omp, int64_t Ns, int64_t Nt, const void* r_src, const void* n_src, const void* v_src, const void* r_trg, void* v_trg, double mu) {
  if (kernel_id < 0 || kernel_id >= NKER) return 1;
  if (real_bytes == 8) eval_f64(kernel_id, omp, Ns, Nt, (const double*)r_src, (const double*)n_src, (const double*)v_src, (const double*)r_trg, (double*)v_trg, mu);
  else if (real_bytes == 4) eval_f32(kernel_id, omp, Ns, Nt, (const float*)r_src, (const float*)n_src, (const float*)v_src, (const float*)r_trg, (float*)v_trg, (float)mu);
  else return 3;
  return 0;
}

int oracle_kernel_matrix(int kernel_id, int real_bytes, int omp, int64_t Ns, int64_t Nt, const void* r_src, const void* n_src, const void* r_trg, void* M, int64_t ldM, double mu) {
  if (kernel_id < 0 || kernel_id >= NKER) return 1;
  if (real_bytes == 8) kernel_matrix_f64(kernel_id, omp, Ns, Nt, (const double*)r_src, (const double*)n_src, (const double*)r_trg, (double*)M, ldM, mu);
  else if (real_bytes == 4) kernel_matrix_f32(kernel_id, omp, Ns, Nt, (const float*)r_src, (const float*)n_src, (const float*)r_trg, (float*)M, ldM, (float)mu);
  else return 3;
  return 0;
}

int oracle_rsqrt_newton(int real_bytes, int seed_bits, int max_iter, int64_t n, const void* x, void* y) {
  if (real_bytes == 8) rsqrt_newton_f64(seed_bits, max_iter, n, (const double*)x, (double*)y);
  else if (real_bytes == 4) rsqrt_newton_f32(seed_bits, max_iter, n, (const float*)x, (float*)y);
  else return 3;
  return 0;
}

/* Newton count the reference picks for a 14-bit seed: mylog2((Integer)(digits/4.2144199393)),
 * intrin-wrapper.hpp:560-562, 3384; digits == -1 -> (Integer)(SigBits*0.3010299957), vec.txx:376. */
int oracle_newton_iters(int real_bytes, int digits) {
  if (digits == -1) digits = (int)((real_bytes == 8 ? 52 : 23) * 0.3010299957);
  int x = (int)(digits / 4.2144199393), l = 0;
  while (x >= 1) { l++; x /= 2; }
  return l;
}

int oracle_dims(int kernel_id, int* k0, int* k1, int* nd, int* flops, double* scale64, float* scale32, char* name, int name_len) {
  if (kernel_id < 0 || kernel_id >= NKER) return 1;
  *k0 = KDIMS[kernel_id][0]; *k1 = KDIMS[kernel_id][1]; *nd = KDIMS[kernel_id][2]; *flops = KFLOPS[kernel_id];
  *scale64 = scale_factor_f64(kernel_id); *scale32 = scale_factor_f32(kernel_id);
  strncpy(name, KNAMES[kernel_id], (size_t)name_len - 1); name[name_len - 1] = 0;
  return 0;
}

/* SURVEY.md section 8(d) synthetic-input generator: counter-based SplitMix64, u(seed,i) in [0,1). */
void oracle_splitmix_fill(uint64_t seed, int64_t first, int64_t n, double offset, double* out) {
  for (int64_t i = 0; i < n; i++) {
    uint64_t z = (seed << 32) + (uint64_t)(first + i) + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    out[i] = (double)(z >> 11) * 0x1.0p-53 + offset;
  }
}

int oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
