// Host-layer test for include/sctl_cuda/generic_kernel.hpp, written in the style of the reference's own
// src/test-generic-kernel.cpp and src/test-kernel_functions.cpp (CHECK / approx_eq / PASS summary) and pinning
// the same analytic fixtures (test-generic-kernel.cpp:40-103, test-kernel_functions.cpp:36-131, 488-526),
// but through the GPU path.  Needs a B200; compiled by __graft_entry__.build(), run by tests/test_gpu_cpp.py.
#include <cmath>
#include <cstdio>
#include <string>

#include "sctl_cuda/generic_kernel.hpp"
#include "sctl_cuda/mini_containers.hpp"

namespace test_utils {
static int failures = 0;
template <class T> static inline bool approx_eq(T a, T b, T tol = T(1e-12)) {
  const T diff = std::fabs(a - b);
  const T mag = std::fabs(a) > std::fabs(b) ? std::fabs(a) : std::fabs(b);
  return diff <= tol || diff <= tol * mag;
}
}  // namespace test_utils
#define CHECK(...)                                                  \
  do {                                                              \
    if (!(__VA_ARGS__)) {                                           \
      std::printf("  FAIL @%d: %s\n", __LINE__, #__VA_ARGS__);      \
      ++test_utils::failures;                                       \
    }                                                               \
  } while (0)

using sctl_cuda::Long;
using sctl_cuda::Matrix;
using sctl_cuda::Vector;
static const double PI = 3.14159265358979323846;

template <class R> static void generic_kernel_fixtures() {
  using Kernel = sctl_cuda::Laplace3D_FxU;
  Kernel K;
  const R tol = sizeof(R) == 8 ? (R)1e-9 : (R)2e-6;
  CHECK(K.CoordDim() == 3);
  CHECK(K.SrcDim() == 1);
  CHECK(K.TrgDim() == 1);
  CHECK(K.NormalDim() == 0);
  CHECK(K.Name() == std::string("Laplace3D-FxU"));
  const R sf = sctl_cuda::kernel_impl::Laplace3D_FxU::template uKerScaleFactor<R>();
  CHECK(test_utils::approx_eq(sf, (R)(1.0 / (4.0 * PI)), (R)(sizeof(R) == 8 ? 1e-12 : 1e-6)));
  {  // 1 x 1 KernelMatrix = 1/(4 pi)
    Vector<R> Xt({1.0, 0.0, 0.0}), Xs({0.0, 0.0, 0.0}), Xn;
    Matrix<R> M;
    K.template KernelMatrix<R, false>(M, Xt, Xs, Xn);
    CHECK(M.Dim(0) == 1 && M.Dim(1) == 1);
    CHECK(test_utils::approx_eq(M(0, 0), (R)(1.0 / (4.0 * PI)), tol));
  }
  {  // single source, targets at r = 1, 2, 3, 5
    Vector<R> Xs({0.0, 0.0, 0.0}), v_s({1.0}), Xn;
    Vector<R> Xt({1.0, 0.0, 0.0, 2.0, 0.0, 0.0, 3.0, 0.0, 0.0, 5.0, 0.0, 0.0});
    Vector<R> v_t;
    K.template Eval<R, false>(v_t, Xt, Xs, Xn, v_s);
    CHECK(v_t.Dim() == 4);
    const R d[4] = {1, 2, 3, 5};
    for (int i = 0; i < 4; i++) CHECK(test_utils::approx_eq(v_t[i], sf / d[i], tol));
    // Eval accumulates into a correctly sized v_trg (generic-kernel.txx:129-132,206)
    K.template Eval<R, true>(v_t, Xt, Xs, Xn, v_s);
    for (int i = 0; i < 4; i++) CHECK(test_utils::approx_eq(v_t[i], 2 * sf / d[i], tol));
    // ... and resets a wrongly sized one
    Vector<R> v_bad({7.0, 7.0});
    K.template Eval<R, false>(v_bad, Xt, Xs, Xn, v_s);
    CHECK(v_bad.Dim() == 4);
    for (int i = 0; i < 4; i++) CHECK(test_utils::approx_eq(v_bad[i], sf / d[i], tol));
    // runtime-digits trampoline (generic-kernel.hpp:99)
    Vector<R> v_tr;
    Kernel::template Eval<R, false>(v_tr, Xt, Xs, Xn, v_s, 10, (const char*)&K);
    for (int i = 0; i < 4; i++) CHECK(test_utils::approx_eq(v_tr[i], sf / d[i], tol));
  }
  {  // 2-source linearity
    Vector<R> Xs({0.0, 0.0, 0.0, 4.0, 0.0, 0.0}), v_s({1.0, 1.0}), Xn, Xt({2.0, 0.0, 0.0}), v_t;
    K.template Eval<R, false>(v_t, Xt, Xs, Xn, v_s);
    CHECK(v_t.Dim() == 1);
    CHECK(test_utils::approx_eq(v_t[0], sf / 2 + sf / 2, tol));
  }
  {  // coincident source/target contributes exactly 0
    Vector<R> Xs({0.5, 0.25, 0.125, 2.5, 0.25, 0.125}), v_s({3.0, 1.0}), Xn, Xt({0.5, 0.25, 0.125}), v_t;
    K.template Eval<R, false>(v_t, Xt, Xs, Xn, v_s);
    CHECK(test_utils::approx_eq(v_t[0], sf / 2, tol));
  }
}

template <class R> static void kernel_function_fixtures() {
  const R tol = sizeof(R) == 8 ? (R)1e-9 : (R)2e-6;
  {  // Laplace FxU: f = 3, r = 2 -> 3/(8 pi)
    sctl_cuda::Laplace3D_FxU K;
    Vector<R> Xs({0, 0, 0}), Xt({2, 0, 0}), Xn, f({3}), v;
    K.template Eval<R, false>(v, Xt, Xs, Xn, f);
    CHECK(test_utils::approx_eq(v[0], (R)(3.0 / (8.0 * PI)), tol));
  }
  {  // Laplace DxU: n = (1,0,0), r = (2,0,0) -> 1/(16 pi)
    sctl_cuda::Laplace3D_DxU K;
    CHECK(K.SrcDim() == 1 && K.TrgDim() == 1 && K.NormalDim() == 3);
    CHECK(K.Name() == std::string("Laplace3D-DxU"));
    Vector<R> Xs({0, 0, 0}), Xt({2, 0, 0}), Xn({1, 0, 0}), f({1}), v;
    K.template Eval<R, false>(v, Xt, Xs, Xn, f);
    CHECK(test_utils::approx_eq(v[0], (R)(1.0 / (16.0 * PI)), tol));
  }
  {  // Laplace FxdU: r = (3,0,0) -> |v| = 1/(36 pi), transverse components 0
    sctl_cuda::Laplace3D_FxdU K;
    CHECK(K.SrcDim() == 1 && K.TrgDim() == 3 && K.NormalDim() == 0);
    Vector<R> Xs({0, 0, 0}), Xt({3, 0, 0}), Xn, f({1}), v;
    K.template Eval<R, false>(v, Xt, Xs, Xn, f);
    CHECK(v.Dim() == 3);
    CHECK(test_utils::approx_eq((R)std::fabs(v[0]), (R)(1.0 / (36.0 * PI)), tol));
    CHECK(v[0] < 0);  // scale factor is -1/(4 pi) (kernel_functions.hpp:90)
    CHECK(std::fabs(v[1]) < tol && std::fabs(v[2]) < tol);
  }
  {  // Stokes FxU: f = (1,0,0), r = (2,0,0) -> (1/(8 pi), 0, 0)
    sctl_cuda::Stokes3D_FxU K;
    CHECK(K.SrcDim() == 3 && K.TrgDim() == 3 && K.NormalDim() == 0);
    Vector<R> Xs({0, 0, 0}), Xt({2, 0, 0}), Xn, f({1, 0, 0}), v;
    K.template Eval<R, false>(v, Xt, Xs, Xn, f);
    CHECK(test_utils::approx_eq(v[0], (R)(1.0 / (8.0 * PI)), tol));
    CHECK(std::fabs(v[1]) < tol && std::fabs(v[2]) < tol);
  }
  {  // Stokes DxU: dims + analytic value  u = 3/(4 pi) r_i r_j r_k n_k f_j / r^5 ; r=(2,0,0), n=f=(1,0,0) -> 3/(4 pi)/4
    sctl_cuda::Stokes3D_DxU K;
    CHECK(K.SrcDim() == 3 && K.TrgDim() == 3 && K.NormalDim() == 3);
    CHECK(K.Name() == std::string("Stokes3D-DxU"));
    Vector<R> Xs({0, 0, 0}), Xt({2, 0, 0}), Xn({1, 0, 0}), f({1, 0, 0}), v;
    K.template Eval<R, false>(v, Xt, Xs, Xn, f);
    CHECK(test_utils::approx_eq(v[0], (R)(3.0 / (4.0 * PI) / 4.0), tol));
  }
  // Eval == KernelMatrix^T f on 7 sources x 5 targets (test-kernel_functions.cpp:488-526)
  auto check = [&](auto K, Long K0, Long K1, Long ND) {
    const Long Ns = 7, Nt = 5;
    Vector<R> Xs(Ns * 3), Xn(Ns * ND), Vs(Ns * K0), Xt(Nt * 3);
    for (Long i = 0; i < Xs.Dim(); i++) Xs[i] = (R)(0.1 * (i + 1) - 0.35);
    for (Long i = 0; i < Vs.Dim(); i++) Vs[i] = (R)(0.3 * (i % 5) - 0.4);
    for (Long i = 0; i < Xt.Dim(); i++) Xt[i] = (R)(1.7 + 0.2 * i);
    for (Long i = 0; i < Ns && ND; i++) { Xn[i * 3 + 0] = 0; Xn[i * 3 + 1] = 1; Xn[i * 3 + 2] = 0; }
    Vector<R> Vt;
    K.template Eval<R, false>(Vt, Xt, Xs, Xn, Vs);
    Matrix<R> M;
    K.template KernelMatrix<R, false>(M, Xt, Xs, Xn);
    CHECK(M.Dim(0) == Ns * K0 && M.Dim(1) == Nt * K1);
    for (Long t = 0; t < Nt; t++)
      for (Long k1 = 0; k1 < K1; k1++) {
        R acc = 0;
        for (Long is = 0; is < Ns; is++)
          for (Long k0 = 0; k0 < K0; k0++) acc += M[is * K0 + k0][t * K1 + k1] * Vs[is * K0 + k0];
        CHECK(test_utils::approx_eq(Vt[t * K1 + k1], acc, (R)(sizeof(R) == 8 ? 1e-10 : 2e-5)));
      }
  };
  check(sctl_cuda::Laplace3D_FxU(), 1, 1, 0);
  check(sctl_cuda::Laplace3D_DxU(), 1, 1, 3);
  check(sctl_cuda::Laplace3D_FxdU(), 1, 3, 0);
  check(sctl_cuda::Stokes3D_FxU(), 3, 3, 0);
  check(sctl_cuda::Stokes3D_DxU(), 3, 3, 3);
}

int main() {
  std::printf("devices: %d\n", sctl_cuda_device_count());
  generic_kernel_fixtures<double>();
  generic_kernel_fixtures<float>();
  kernel_function_fixtures<double>();
  kernel_function_fixtures<float>();
  std::printf("\n%s (%d check(s) failed)\n", test_utils::failures == 0 ? "PASS" : "FAIL", test_utils::failures);
  return test_utils::failures == 0 ? 0 : 1;
}
