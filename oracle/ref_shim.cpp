// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// extern "C" wrapper around the UNMODIFIED reference (SCTL) CPU implementation of the hot path.
// This translation unit contains no SCTL code: it only #includes the reference headers where they
// lie (-I/root/reference/include, see oracle/Makefile) and instantiates
//   sctl::GenericKernel<uKernel>::Eval<Real,omp,digits>        (generic-kernel.txx:102-211)
//   sctl::GenericKernel<uKernel>::KernelMatrix<Real,omp,digits> (generic-kernel.txx:213-329)
//   sctl::approx_rsqrt<digits>(x, mask)                         (vec.txx:379-382)
// on non-owning Vector/Matrix views (Vector(dim, ptr, own_data=false), vector.hpp:51) of the caller's
// buffers.  Output goes to oracle/_ref/libsctl_ref_<isa>.so (git-ignored, travels to the GPU box).
//
// It is a separate shared object on purpose: the drop-in headers of this repo specialise members of
// the same sctl::GenericKernel<> templates, so both cannot live in one image (ODR).
#include <cstdint>
#include <cstring>
#include <omp.h>
#include <sctl.hpp>

namespace {

template <class Real, class Ker, bool omp, sctl::Integer digits>
void eval_t(int64_t Ns, int64_t Nt, const void* r_src, const void* n_src, const void* v_src, const void* r_trg, void* v_trg, const void* ctx) {
  using namespace sctl;
  Ker ker;
  ker.SetCtxPtr((void*)ctx);  // Helmholtz: pointer to the wavenumber, of type Real (kernel_functions.hpp:484-488)
  constexpr Integer K0 = Ker::SrcDim(), K1 = Ker::TrgDim(), ND = Ker::NormalDim();
  const Vector<Real> Xs(Ns * 3, (Iterator<Real>)(Real*)r_src, false);
  const Vector<Real> Xn(n_src ? Ns * ND : 0, (Iterator<Real>)(Real*)n_src, false);
  const Vector<Real> Vs(Ns * K0, (Iterator<Real>)(Real*)v_src, false);
  const Vector<Real> Xt(Nt * 3, (Iterator<Real>)(Real*)r_trg, false);
  Vector<Real> Vt(Nt * K1, (Iterator<Real>)(Real*)v_trg, false);  // right size => Eval accumulates in place
  ker.template Eval<Real, omp, digits>(Vt, Xt, Xs, Xn, Vs);
}

template <class Real, class Ker, bool omp, sctl::Integer digits>
void kmat_t(int64_t Ns, int64_t Nt, const void* r_src, const void* n_src, const void* r_trg, void* M, const void* ctx) {
  using namespace sctl;
  Ker ker;
  ker.SetCtxPtr((void*)ctx);
  constexpr Integer K0 = Ker::SrcDim(), K1 = Ker::TrgDim(), ND = Ker::NormalDim();
  const Vector<Real> Xs(Ns * 3, (Iterator<Real>)(Real*)r_src, false);
  const Vector<Real> Xn(n_src ? Ns * ND : 0, (Iterator<Real>)(Real*)n_src, false);
  const Vector<Real> Xt(Nt * 3, (Iterator<Real>)(Real*)r_trg, false);
  Matrix<Real> M_(Ns * K0, Nt * K1, (Iterator<Real>)(Real*)M, false);
  ker.template KernelMatrix<Real, omp, digits>(M_, Xt, Xs, Xn);
}

template <class Real, class Ker, bool omp, class... A> int eval_digits(int digits, A... a) {
  switch (digits) {  // the digit classes that differ in Newton count on any ISA plus the callers' 10/11
    case -1: eval_t<Real, Ker, omp, -1>(a...); return 0;
    case 3: eval_t<Real, Ker, omp, 3>(a...); return 0;
    case 6: eval_t<Real, Ker, omp, 6>(a...); return 0;
    case 10: eval_t<Real, Ker, omp, 10>(a...); return 0;
    case 11: eval_t<Real, Ker, omp, 11>(a...); return 0;
    case 15: eval_t<Real, Ker, omp, 15>(a...); return 0;
    default: return 2;
  }
}
template <class Real, class Ker, class... A> int eval_omp(int omp, int digits, A... a) {
  return omp ? eval_digits<Real, Ker, true>(digits, a...) : eval_digits<Real, Ker, false>(digits, a...);
}
template <class Real, class... A> int eval_ker(int kernel_id, int omp, int digits, A... a) {
  switch (kernel_id) {
    case 0: return eval_omp<Real, sctl::Laplace3D_FxU>(omp, digits, a...);
    case 1: return eval_omp<Real, sctl::Laplace3D_DxU>(omp, digits, a...);
    case 2: return eval_omp<Real, sctl::Laplace3D_FxdU>(omp, digits, a...);
    case 3: return eval_omp<Real, sctl::Stokes3D_FxU>(omp, digits, a...);
    case 4: return eval_omp<Real, sctl::Stokes3D_DxU>(omp, digits, a...);
    case 5: return eval_omp<Real, sctl::Stokes3D_FxT>(omp, digits, a...);
    case 6: return eval_omp<Real, sctl::Stokes3D_FSxU>(omp, digits, a...);
    case 7: return eval_omp<Real, sctl::Stokes3D_FxUP>(omp, digits, a...);
    case 8: return eval_omp<Real, sctl::Laplace3D_Fxd2U>(omp, digits, a...);
    case 9: return eval_omp<Real, sctl::Laplace3D_DxdU>(omp, digits, a...);
    case 10: return eval_omp<Real, sctl::BiotSavart3D_FxU>(omp, digits, a...);
    case 11: return eval_omp<Real, sctl::BiotSavart3D_FxdU>(omp, digits, a...);
    case 12: return eval_omp<Real, sctl::Helmholtz3D_FxU>(omp, digits, a...);
    case 13: return eval_omp<Real, sctl::Helmholtz3D_DxU>(omp, digits, a...);
    case 14: return eval_omp<Real, sctl::Helmholtz3D_FxdU>(omp, digits, a...);
    case 15: return eval_omp<Real, sctl::HelmholtzDiff3D_FxdU>(omp, digits, a...);
    default: return 1;
  }
}

template <class Real, class Ker, class... A> int kmat_omp(int omp, int digits, A... a) {
  if (digits != -1) return 2;
  if (omp) kmat_t<Real, Ker, true, -1>(a...); else kmat_t<Real, Ker, false, -1>(a...);
  return 0;
}
template <class Real, class... A> int kmat_ker(int kernel_id, int omp, int digits, A... a) {
  switch (kernel_id) {
    case 0: return kmat_omp<Real, sctl::Laplace3D_FxU>(omp, digits, a...);
    case 1: return kmat_omp<Real, sctl::Laplace3D_DxU>(omp, digits, a...);
    case 2: return kmat_omp<Real, sctl::Laplace3D_FxdU>(omp, digits, a...);
    case 3: return kmat_omp<Real, sctl::Stokes3D_FxU>(omp, digits, a...);
    case 4: return kmat_omp<Real, sctl::Stokes3D_DxU>(omp, digits, a...);
    case 5: return kmat_omp<Real, sctl::Stokes3D_FxT>(omp, digits, a...);
    case 6: return kmat_omp<Real, sctl::Stokes3D_FSxU>(omp, digits, a...);
    case 7: return kmat_omp<Real, sctl::Stokes3D_FxUP>(omp, digits, a...);
    case 8: return kmat_omp<Real, sctl::Laplace3D_Fxd2U>(omp, digits, a...);
    case 9: return kmat_omp<Real, sctl::Laplace3D_DxdU>(omp, digits, a...);
    case 10: return kmat_omp<Real, sctl::BiotSavart3D_FxU>(omp, digits, a...);
    case 11: return kmat_omp<Real, sctl::BiotSavart3D_FxdU>(omp, digits, a...);
    case 12: return kmat_omp<Real, sctl::Helmholtz3D_FxU>(omp, digits, a...);
    case 13: return kmat_omp<Real, sctl::Helmholtz3D_DxU>(omp, digits, a...);
    case 14: return kmat_omp<Real, sctl::Helmholtz3D_FxdU>(omp, digits, a...);
    case 15: return kmat_omp<Real, sctl::HelmholtzDiff3D_FxdU>(omp, digits, a...);
    default: return 1;
  }
}

template <class Real, sctl::Integer digits> void rsqrt_t(int64_t n, const Real* x, Real* y) {
  using namespace sctl;
  constexpr Integer N = DefaultVecLen<Real>();
  using V = Vec<Real, N>;
  for (int64_t i = 0; i < n; i += N) {
    Real xi[N], yi[N];
    for (Integer k = 0; k < N; k++) xi[k] = (i + k < n ? x[i + k] : (Real)1);
    V vx = V::Load(xi);
    V vy = approx_rsqrt<digits>(vx, vx > V::Zero());
    vy.Store(yi);
    for (Integer k = 0; k < N && i + k < n; k++) y[i + k] = yi[k];
  }
}

}  // namespace

extern "C" {

// v_trg[Nt*K1] += scale * sum_s K(x_t - x_s, n_s) f_s   via the reference's own Eval.
int sctl_ref_eval(int kernel_id, int real_bytes, int omp, int digits, int64_t Ns, int64_t Nt, const void* r_src, const void* n_src, const void* v_src, const void* r_trg, void* v_trg, const void* ctx) {
  if (real_bytes == 8) return eval_ker<double>(kernel_id, omp, digits, Ns, Nt, r_src, n_src, v_src, r_trg, v_trg, ctx);
  if (real_bytes == 4) return eval_ker<float>(kernel_id, omp, digits, Ns, Nt, r_src, n_src, v_src, r_trg, v_trg, ctx);
  return 3;
}

// M[(Ns*K0) x (Nt*K1)] row-major, overwritten, via the reference's own KernelMatrix.
int sctl_ref_kernel_matrix(int kernel_id, int real_bytes, int omp, int digits, int64_t Ns, int64_t Nt, const void* r_src, const void* n_src, const void* r_trg, void* M, const void* ctx) {
  if (real_bytes == 8) return kmat_ker<double>(kernel_id, omp, digits, Ns, Nt, r_src, n_src, r_trg, M, ctx);
  if (real_bytes == 4) return kmat_ker<float>(kernel_id, omp, digits, Ns, Nt, r_src, n_src, r_trg, M, ctx);
  return 3;
}

// y[i] = approx_rsqrt<digits>(x[i], x[i] > 0) with the reference's default SIMD width.
int sctl_ref_rsqrt(int real_bytes, int digits, int64_t n, const void* x, void* y) {
  if (real_bytes == 8) {
    if (digits == -1) rsqrt_t<double, -1>(n, (const double*)x, (double*)y);
    else if (digits == 4) rsqrt_t<double, 4>(n, (const double*)x, (double*)y);
    else if (digits == 7) rsqrt_t<double, 7>(n, (const double*)x, (double*)y);
    else if (digits == 10) rsqrt_t<double, 10>(n, (const double*)x, (double*)y);
    else return 2;
    return 0;
  }
  if (real_bytes == 4) {
    if (digits == -1) rsqrt_t<float, -1>(n, (const float*)x, (float*)y);
    else if (digits == 4) rsqrt_t<float, 4>(n, (const float*)x, (float*)y);
    else if (digits == 7) rsqrt_t<float, 7>(n, (const float*)x, (float*)y);
    else return 2;
    return 0;
  }
  return 3;
}

int sctl_ref_dims(int kernel_id, int* k0, int* k1, int* nd, int* flops, double* scale64, float* scale32, char* name, int name_len) {
  auto fill = [&](auto ker) {
    using Ker = decltype(ker);
    *k0 = Ker::SrcDim(); *k1 = Ker::TrgDim(); *nd = Ker::NormalDim(); *flops = Ker::FLOPS();
    *scale64 = Ker::template uKerScaleFactor<double>();
    *scale32 = Ker::template uKerScaleFactor<float>();
    std::strncpy(name, Ker::Name().c_str(), name_len - 1); name[name_len - 1] = 0;
  };
  switch (kernel_id) {
    case 0: fill(sctl::Laplace3D_FxU()); return 0;
    case 1: fill(sctl::Laplace3D_DxU()); return 0;
    case 2: fill(sctl::Laplace3D_FxdU()); return 0;
    case 3: fill(sctl::Stokes3D_FxU()); return 0;
    case 4: fill(sctl::Stokes3D_DxU()); return 0;
    case 5: fill(sctl::Stokes3D_FxT()); return 0;
    case 6: fill(sctl::Stokes3D_FSxU()); return 0;
    case 7: fill(sctl::Stokes3D_FxUP()); return 0;
    case 8: fill(sctl::Laplace3D_Fxd2U()); return 0;
    case 9: fill(sctl::Laplace3D_DxdU()); return 0;
    case 10: fill(sctl::BiotSavart3D_FxU()); return 0;
    case 11: fill(sctl::BiotSavart3D_FxdU()); return 0;
    case 12: fill(sctl::Helmholtz3D_FxU()); return 0;
    case 13: fill(sctl::Helmholtz3D_DxU()); return 0;
    case 14: fill(sctl::Helmholtz3D_FxdU()); return 0;
    case 15: fill(sctl::HelmholtzDiff3D_FxdU()); return 0;
    default: return 1;
  }
}

// Vector<double>::Write / Matrix<double>::Write (vector.txx:132-144, matrix.txx) -- fixtures for sctl_b200/io.py
int sctl_ref_vector_write(const char* path, int64_t n, const double* x) {
  sctl::Vector<double> v(n, (sctl::Iterator<double>)(double*)x, false);
  v.Write(path);
  return 0;
}
int sctl_ref_matrix_write(const char* path, int64_t rows, int64_t cols, const double* x) {
  sctl::Matrix<double> m(rows, cols, (sctl::Iterator<double>)(double*)x, false);
  m.Write(path);
  return 0;
}

int sctl_ref_threads() { return omp_get_max_threads(); }
int sctl_ref_veclen(int real_bytes) { return real_bytes == 8 ? sctl::DefaultVecLen<double>() : sctl::DefaultVecLen<float>(); }

}  // extern "C"
