/* TEST INFRASTRUCTURE ONLY (see sctl_oracle.c).  Included twice: REAL=double / REAL=float.
 * Every function cites the reference lines (relative to /root/reference/include/sctl/) it restates. */

/* approx_rsqrt<digits>(r2, r2 > 0): vec.txx:379-382.  Scalar/generic lowering = exact 1/sqrt, zero on
 * masked lanes (intrin-wrapper.hpp:563-579); the SIMD lowerings (seed + Newton, :540-559, :3383-3408)
 * agree with it to <= 10^-digits, which is all the contract (vec-test.hpp:431-465) promises. */
static inline REAL FN(rsqrt_masked)(REAL r2) { return r2 > 0 ? (REAL)1 / SQRT(r2) : (REAL)0; }

/* One source-target pair: dX = xt - xs (generic-kernel.txx:108-109), then the micro-kernel.
 * vt is the UNSCALED accumulator (scale applied once at the end, generic-kernel.txx:204-208). */
static inline void FN(pair_matrix)(int kid, REAL* u, const REAL* r, const REAL* ns, REAL mu);
static inline void FN(pair_apply)(int kid, REAL* vt, const REAL* xt, const REAL* xs, const REAL* ns, const REAL* f, REAL mu) {
  REAL r[3];
  for (int k = 0; k < 3; k++) r[k] = xt[k] - xs[k];
  const REAL r2 = r[0]*r[0] + r[1]*r[1] + r[2]*r[2];
  const REAL rinv = FN(rsqrt_masked)(r2);
  switch (kid) {
    case 0: { /* Laplace3D_FxU: uKerMatrix kernel_functions.hpp:43-47, apply generic-kernel.txx:113-119 */
      vt[0] = rinv * f[0] + vt[0];
    } break;
    case 1: { /* Laplace3D_DxU: kernel_functions.hpp:72-78 */
      const REAL rdotn = r[0]*ns[0] + r[1]*ns[1] + r[2]*ns[2];
      const REAL rinv3 = rinv * rinv * rinv;
      vt[0] = (rdotn * rinv3) * f[0] + vt[0];
    } break;
    case 2: { /* Laplace3D_FxdU fused apply: kernel_functions.hpp:102-110 */
      const REAL rinv3 = rinv * rinv * rinv;
      const REAL f_rinv3 = f[0] * rinv3;
      for (int i = 0; i < 3; i++) vt[i] += r[i] * f_rinv3;
    } break;
    case 3: { /* Stokes3D_FxU fused apply: kernel_functions.hpp:166-176 */
      const REAL rinv3 = rinv * rinv * rinv;
      const REAL fdotr_rinv3 = (f[0]*r[0] + f[1]*r[1] + f[2]*r[2]) * rinv3;
      for (int i = 0; i < 3; i++) vt[i] += f[i]*rinv + r[i]*fdotr_rinv3;
    } break;
    case 4: { /* Stokes3D_DxU fused apply: kernel_functions.hpp:205-217 */
      const REAL rinv2 = rinv * rinv;
      const REAL rinv5 = rinv2 * rinv2 * rinv;
      const REAL rdotn_rinv5 = (r[0]*ns[0] + r[1]*ns[1] + r[2]*ns[2]) * rinv5;
      const REAL t = (f[0]*r[0] + f[1]*r[1] + f[2]*r[2]) * rdotn_rinv5;
      for (int i = 0; i < 3; i++) vt[i] += r[i] * t;
    } break;
    case 5: { /* Stokes3D_FxT fused apply: kernel_functions.hpp:241-252 */
      const REAL rinv2 = rinv * rinv;
      const REAL rinv5 = rinv2 * rinv2 * rinv;
      const REAL fdotr_rinv5 = (f[0]*r[0] + f[1]*r[1] + f[2]*r[2]) * rinv5;
      for (int i = 0; i < 3; i++) {
        const REAL ri = r[i] * fdotr_rinv5;
        for (int j = 0; j < 3; j++) vt[i*3+j] += ri * r[j];
      }
    } break;
    case 6: { /* Stokes3D_FSxU fused apply: kernel_functions.hpp:311-321 */
      const REAL rinv3 = rinv * rinv * rinv;
      const REAL t = (f[0]*r[0] + f[1]*r[1] + f[2]*r[2] + f[3]) * rinv3;
      for (int i = 0; i < 3; i++) vt[i] += f[i]*rinv + r[i]*t;
    } break;
    case 7: { /* Stokes3D_FxUP fused apply: kernel_functions.hpp:351-362 */
      const REAL rinv3 = rinv * rinv * rinv;
      const REAL t = (f[0]*r[0] + f[1]*r[1] + f[2]*r[2]) * rinv3;
      for (int i = 0; i < 3; i++) vt[i] += f[i]*rinv + r[i]*t;
      vt[3] += t;
    } break;
    case 8: case 9: case 11: case 15: { /* no fused apply: uKerMatrix then FMA, generic-kernel.txx:113-119 */
      REAL u[27];
      const int K0 = KDIMS[kid][0], K1 = KDIMS[kid][1];
      FN(pair_matrix)(kid, u, r, ns, mu);
      for (int k0 = 0; k0 < K0; k0++) for (int k1 = 0; k1 < K1; k1++) vt[k1] = u[k0*K1+k1] * f[k0] + vt[k1];
    } break;
    case 10: { /* BiotSavart3D_FxU fused apply: kernel_functions.hpp:436-447 */
      const REAL rinv3 = rinv * rinv * rinv;
      vt[0] += (f[1]*r[2] - r[1]*f[2]) * rinv3;
      vt[1] += (f[2]*r[0] - r[2]*f[0]) * rinv3;
      vt[2] += (f[0]*r[1] - r[0]*f[1]) * rinv3;
    } break;
    case 12: { /* Helmholtz3D_FxU fused apply: kernel_functions.hpp:519-532 */
      const REAL arg = r2 * rinv * mu, sn = SIN(arg), cs = COS(arg);
      const REAL G0 = cs * rinv, G1 = sn * rinv;
      vt[0] += f[0]*G0 - f[1]*G1;
      vt[1] += f[0]*G1 + f[1]*G0;
    } break;
    case 13: { /* Helmholtz3D_DxU fused apply: kernel_functions.hpp:561-577 */
      const REAL rinv2 = rinv * rinv;
      const REAL ndotr = ns[0]*r[0] + ns[1]*r[1] + ns[2]*r[2];
      const REAL arg = r2 * rinv * mu, sn = SIN(arg), cs = COS(arg);
      const REAL G0 = (-mu*sn - cs * rinv) * rinv2 * ndotr;
      const REAL G1 = ( mu*cs - sn * rinv) * rinv2 * ndotr;
      vt[0] += f[0]*G0 - f[1]*G1;
      vt[1] += f[0]*G1 + f[1]*G0;
    } break;
    case 14: { /* Helmholtz3D_FxdU fused apply: kernel_functions.hpp:607-626 */
      const REAL rinv2 = rinv * rinv;
      const REAL arg = r2 * rinv * mu, sn = SIN(arg), cs = COS(arg);
      const REAL G0 = (-mu*sn - cs * rinv) * rinv2;
      const REAL G1 = ( mu*cs - sn * rinv) * rinv2;
      const REAL fG0 = f[0]*G0 - f[1]*G1, fG1 = f[0]*G1 + f[1]*G0;
      for (int i = 0; i < 3; i++) { vt[i*2+0] += fG0 * r[i]; vt[i*2+1] += fG1 * r[i]; }
    } break;
  }
}

/* uKerMatrix for one displacement r (= xt - xs): kernel_functions.hpp:43-47, 72-78, 92-99, 136-145, 191-202.
 * u is K0 x K1 row-major, unscaled. */
static inline void FN(pair_matrix)(int kid, REAL* u, const REAL* r, const REAL* ns, REAL mu) {
  const REAL r2 = r[0]*r[0] + r[1]*r[1] + r[2]*r[2];
  const REAL rinv = FN(rsqrt_masked)(r2);
  switch (kid) {
    case 0: u[0] = rinv; break;
    case 1: {
      const REAL rdotn = r[0]*ns[0] + r[1]*ns[1] + r[2]*ns[2];
      const REAL rinv3 = rinv * rinv * rinv;
      u[0] = rdotn * rinv3;
    } break;
    case 2: {
      const REAL rinv3 = rinv * rinv * rinv;
      for (int i = 0; i < 3; i++) u[i] = r[i] * rinv3;
    } break;
    case 3: {
      const REAL rinv3 = rinv * rinv * rinv;
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) u[i*3+j] = (i == j ? rinv : (REAL)0) + r[i]*r[j]*rinv3;
    } break;
    case 4: {
      const REAL rinv2 = rinv * rinv;
      const REAL rinv5 = rinv2 * rinv2 * rinv;
      const REAL rdotn_rinv5 = (r[0]*ns[0] + r[1]*ns[1] + r[2]*ns[2]) * rinv5;
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) u[i*3+j] = r[i]*r[j]*rdotn_rinv5;
    } break;
    case 5: { /* Stokes3D_FxT: kernel_functions.hpp:230-239 */
      const REAL rinv2 = rinv * rinv;
      const REAL rinv5 = rinv2 * rinv2 * rinv;
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) u[i*9+j*3+k] = r[i]*r[j]*r[k]*rinv5;
    } break;
    case 6: { /* Stokes3D_FSxU: kernel_functions.hpp:267-279 */
      const REAL rinv3 = rinv * rinv * rinv;
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) u[i*3+j] = (i == j ? rinv : (REAL)0) + r[i]*r[j]*rinv3;
      for (int j = 0; j < 3; j++) u[9+j] = r[j]*rinv3;
    } break;
    case 7: { /* Stokes3D_FxUP: kernel_functions.hpp:337-349 */
      const REAL rinv3 = rinv * rinv * rinv;
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) u[i*4+j] = (i == j ? rinv : (REAL)0) + r[i]*r[j]*rinv3;
      for (int i = 0; i < 3; i++) u[i*4+3] = r[i]*rinv3;
    } break;
    case 8: { /* Laplace3D_Fxd2U: kernel_functions.hpp:376-388 */
      const REAL rinv2 = rinv * rinv;
      const REAL rinv3 = rinv * rinv2;
      const REAL t5 = (REAL)3 * rinv3 * rinv2;
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) u[i*3+j] = t5 * r[i] * r[j] - (i == j ? rinv3 : (REAL)0);
    } break;
    case 9: { /* Laplace3D_DxdU: kernel_functions.hpp:402-411 */
      const REAL rinv2 = rinv * rinv;
      const REAL rinv3 = rinv * rinv2;
      const REAL rdotn = r[0]*ns[0] + r[1]*ns[1] + r[2]*ns[2];
      const REAL t5 = (REAL)3 * rdotn * rinv3 * rinv2;
      for (int i = 0; i < 3; i++) u[i] = ns[i] * rinv3 - r[i] * t5;
    } break;
    case 10: { /* BiotSavart3D_FxU: kernel_functions.hpp:425-433 */
      const REAL rinv3 = rinv * rinv * rinv;
      u[0] = 0;           u[1] = -r[2]*rinv3; u[2] =  r[1]*rinv3;
      u[3] =  r[2]*rinv3; u[4] = 0;           u[5] = -r[0]*rinv3;
      u[6] = -r[1]*rinv3; u[7] =  r[0]*rinv3; u[8] = 0;
    } break;
    case 11: { /* BiotSavart3D_FxdU: kernel_functions.hpp:460-482; u[k0][k1] at u[k0*9+k1] */
      const REAL rinv2 = rinv * rinv;
      const REAL rinv3 = rinv * rinv2;
      const REAL t5 = (REAL)3 * rinv3 * rinv2;
      const REAL Z = 0;
      REAL (*U)[9] = (REAL (*)[9])u;
      U[0][0] =                    Z; U[1][0] =         t5*r[2]*r[0]; U[2][0] =        -t5*r[1]*r[0];
      U[0][1] =                    Z; U[1][1] =         t5*r[2]*r[1]; U[2][1] =  rinv3 -t5*r[1]*r[1];
      U[0][2] =                    Z; U[1][2] = -rinv3 +t5*r[2]*r[2]; U[2][2] =        -t5*r[1]*r[2];
      U[0][3] =        -t5*r[2]*r[0]; U[1][3] =                    Z; U[2][3] = -rinv3 +t5*r[0]*r[0];
      U[0][4] =        -t5*r[2]*r[1]; U[1][4] =                    Z; U[2][4] =         t5*r[0]*r[1];
      U[0][5] =  rinv3 -t5*r[2]*r[2]; U[1][5] =                    Z; U[2][5] =         t5*r[0]*r[2];
      U[0][6] =         t5*r[1]*r[0]; U[1][6] =  rinv3 -t5*r[0]*r[0]; U[2][6] =                    Z;
      U[0][7] = -rinv3 +t5*r[1]*r[1]; U[1][7] =        -t5*r[0]*r[1]; U[2][7] =                    Z;
      U[0][8] =         t5*r[1]*r[2]; U[1][8] =        -t5*r[0]*r[2]; U[2][8] =                    Z;
    } break;
    case 12: { /* Helmholtz3D_FxU: kernel_functions.hpp:506-516 */
      const REAL arg = r2 * rinv * mu, sn = SIN(arg), cs = COS(arg);
      u[0] = cs * rinv; u[1] = sn * rinv; u[2] = -u[1]; u[3] = u[0];
    } break;
    case 13: { /* Helmholtz3D_DxU: kernel_functions.hpp:546-558 */
      const REAL rinv2 = rinv * rinv;
      const REAL ndotr = ns[0]*r[0] + ns[1]*r[1] + ns[2]*r[2];
      const REAL arg = r2 * rinv * mu, sn = SIN(arg), cs = COS(arg);
      u[0] = (-mu*sn - cs * rinv) * rinv2 * ndotr;
      u[1] = ( mu*cs - sn * rinv) * rinv2 * ndotr;
      u[2] = -u[1]; u[3] = u[0];
    } break;
    case 14: { /* Helmholtz3D_FxdU: kernel_functions.hpp:591-604 */
      const REAL rinv2 = rinv * rinv;
      const REAL arg = r2 * rinv * mu, sn = SIN(arg), cs = COS(arg);
      const REAL G0 = (-mu*sn - cs * rinv) * rinv2;
      const REAL G1 = ( mu*cs - sn * rinv) * rinv2;
      for (int i = 0; i < 3; i++) { u[i*2+0] = G0 * r[i]; u[i*2+1] = G1 * r[i]; u[6+i*2+0] = -u[i*2+1]; u[6+i*2+1] = u[i*2+0]; }
    } break;
    case 15: { /* HelmholtzDiff3D_FxdU: kernel_functions.hpp:640-656 (half-angle form) */
      const REAL rinv2 = rinv * rinv;
      const REAL arg = r2 * rinv * mu * (REAL)0.5, sh = SIN(arg), ch = COS(arg);
      const REAL sn = (REAL)2 * sh * ch;
      const REAL cm1 = -(REAL)2 * sh * sh;
      const REAL cs = (REAL)1 + cm1;
      const REAL G0 = (-mu*sn - cm1 * rinv) * rinv2;
      const REAL G1 = ( mu*cs - sn  * rinv) * rinv2;
      for (int i = 0; i < 3; i++) { u[i*2+0] = G0 * r[i]; u[i*2+1] = G1 * r[i]; u[6+i*2+0] = -u[i*2+1]; u[6+i*2+1] = u[i*2+0]; }
    } break;
  }
}

/* GenericKernel::Eval, generic-kernel.txx:102-211.  Per target: zero accumulator (:179), sources in
 * index order (:181-186), then v_trg[t*K1+k] += acc[k] * scale (:204-208) -- ACCUMULATES into v_trg
 * (the caller-side ReInit+SetZero on a size mismatch, :129-132, is the wrapper's job). */
static void FN(eval)(int kid, int omp, int64_t Ns, int64_t Nt, const REAL* r_src, const REAL* n_src, const REAL* v_src, const REAL* r_trg, REAL* v_trg, REAL mu) {
  const int K0 = KDIMS[kid][0], K1 = KDIMS[kid][1], ND = KDIMS[kid][2];
  const REAL scale = (REAL)FN(scale_factor)(kid);
  #pragma omp parallel for schedule(static) if (omp)
  for (int64_t t = 0; t < Nt; t++) {
    REAL acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t s = 0; s < Ns; s++)
      FN(pair_apply)(kid, acc, r_trg + t*3, r_src + s*3, ND ? n_src + s*ND : (const REAL*)0, v_src + s*K0, mu);
    for (int k = 0; k < K1; k++) v_trg[t*K1 + k] += acc[k] * scale;
  }
}

/* GenericKernel::KernelMatrix, generic-kernel.txx:213-329: M[(s*K0+k0)][(t*K1+k1)] = u[k0][k1] * scale,
 * OVERWRITES (:268,306), row-major (Ns*K0) x (Nt*K1) with leading dimension ldM. */
static void FN(kernel_matrix)(int kid, int omp, int64_t Ns, int64_t Nt, const REAL* r_src, const REAL* n_src, const REAL* r_trg, REAL* M, int64_t ldM, REAL mu) {
  const int K0 = KDIMS[kid][0], K1 = KDIMS[kid][1], ND = KDIMS[kid][2];
  const REAL scale = (REAL)FN(scale_factor)(kid);
  #pragma omp parallel for schedule(static) if (omp)
  for (int64_t s = 0; s < Ns; s++) {
    for (int64_t t = 0; t < Nt; t++) {
      REAL r[3], u[27];
      for (int k = 0; k < 3; k++) r[k] = r_trg[t*3+k] - r_src[s*3+k];
      FN(pair_matrix)(kid, u, r, ND ? n_src + s*ND : (const REAL*)0, mu);
      for (int k0 = 0; k0 < K0; k0++) for (int k1 = 0; k1 < K1; k1++) M[(s*K0+k0)*ldM + t*K1+k1] = u[k0*K1+k1] * scale;
    }
  }
}

/* rsqrt_newton_iter<MAX,ITER> (intrin-wrapper.hpp:540-559) on a seed truncated to seed_bits mantissa bits:
 * y <- y*(x*y*y + c1_k), c1_k = -3*2^(3^k-1) for k = 0..MAX-1, one final multiply by
 * c2 = (-1/2)^((3^(MAX-1)-1)*3/2+1); MAX = mylog2((int)(digits/4.2144199393)) for the 14-bit seed (:3384). */
static void FN(rsqrt_newton)(int seed_bits, int max_iter, int64_t n, const REAL* x, REAL* y) {
  for (int64_t i = 0; i < n; i++) {
    if (!(x[i] > 0)) { y[i] = 0; continue; }
    int e; double m = frexp(1.0 / sqrt((double)x[i]), &e);
    m = ldexp(floor(ldexp(m, seed_bits)), -seed_bits);  /* truncated seed, relative error < 2^-(seed_bits-1) */
    REAL yy = (REAL)ldexp(m, e);
    int64_t p3 = 1;  /* 3^k */
    for (int k = 0; k < max_iter; k++) {
      const REAL c1 = (REAL)(-3.0 * ldexp(1.0, (int)(p3 - 1)));
      yy = yy * (x[i] * (yy * yy) + c1);
      if (k == max_iter - 1) {
        const int64_t ex = (p3 - 1) * 3 / 2 + 1;
        yy = yy * (REAL)((ex & 1 ? -1.0 : 1.0) * ldexp(1.0, (int)-ex));
      }
      p3 *= 3;
    }
    y[i] = yy;
  }
}
