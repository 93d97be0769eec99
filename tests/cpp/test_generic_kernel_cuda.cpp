// Host-layer test for include/sctl_cuda/generic_kernel.hpp, written in the style of the reference's own
// src/test-generic-kernel.cpp and src/test-kernel_functions.cpp (CHECK / approx_eq / PASS summary) and pinning
// the same analytic fixtures (test-generic-kernel.cpp:40-103, test-kernel_functions.cpp:36-131, 488-526),
// but through the GPU path.  Needs a B200; compiled by __graft_entry__.build(), run by tests/test_gpu_cpp.py.
#include <cmath>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>
#include <algorithm>

#include "sctl_cuda/generic_kernel.hpp"
#include "sctl_cuda/mini_containers.hpp"
#include "sctl_cuda/bio_apply.hpp"
#include "sctl_cuda/near_blocks.hpp"
#include "sctl_cuda/particle_direct.hpp"

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

// particle_direct.hpp / near_blocks.hpp with the mini containers: the resident evaluator must reproduce Eval, the
// batched near blocks must reproduce per-block KernelMatrix + GEMV (the same device arithmetic, so tight tolerances).
template <class R> static void session_layers() {
  const R tol = sizeof(R) == 8 ? (R)1e-13 : (R)2e-5;
  const Long Ns = 700, Nt = 450;
  Vector<R> Xs(Ns * 3), Xn(Ns * 3), F(Ns * 3), Xt(Nt * 3), W(Ns);
  unsigned long long z = 88172645463325252ULL;
  auto rnd = [&z]() { z ^= z << 13; z ^= z >> 7; z ^= z << 17; return (R)((z >> 11) * (1.0 / 9007199254740992.0)) - (R)0.5; };
  for (Long i = 0; i < Ns * 3; i++) { Xs[i] = rnd(); Xn[i] = rnd(); F[i] = rnd(); }
  for (Long i = 0; i < Nt * 3; i++) Xt[i] = rnd();
  for (Long i = 0; i < Ns; i++) W[i] = (R)0.01 + (R)0.001 * rnd();
  sctl_cuda::Stokes3D_DxU dl;
  sctl_cuda::Stokes3D_FxU sl;
  {
    sctl_cuda::ParticleDirect<R> pd;
    pd.SetAccuracy(-1);
    pd.AddSrc("S", sl, sl);
    pd.AddSrc("D", dl, dl);
    pd.AddTrg("T", sl, sl);
    pd.SetKernelS2T("S", "T", sl);
    pd.SetKernelS2T("D", "T", dl);
    pd.SetSrcCoord("S", Xs);
    pd.SetSrcCoord("D", Xs, Xn);
    pd.SetSrcDensity("S", F);
    pd.SetSrcDensity("D", F);
    pd.SetTrgCoord("T", Xt);
    Vector<R> U, V, none;
    pd.Eval(U, "T");
    sl.template Eval<R, true>(V, Xt, Xs, none, F);
    dl.template Eval<R, true>(V, Xt, Xs, Xn, F);
    CHECK(U.Dim() == Nt * 3 && V.Dim() == Nt * 3);
    R err = 0, mag = 0;
    for (Long i = 0; i < U.Dim(); i++) { err = std::fmax(err, std::fabs(U[i] - V[i])); mag = std::fmax(mag, std::fabs(V[i])); }
    CHECK(err <= tol * mag);
    // slot numbers are bounded by the C ABI (4096): deleted sources / targets hand theirs back, so long-running callers
    // that add and delete particle types keep working
    for (int it = 0; it < 5000; it++) {
      pd.AddSrc("tmp", sl, sl);
      pd.AddTrg("tmpT", sl, sl);
      if (it == 4999) {
        pd.SetKernelS2T("tmp", "tmpT", sl);
        pd.SetSrcCoord("tmp", Xs);
        pd.SetSrcDensity("tmp", F);
        pd.SetTrgCoord("tmpT", Xt);
        Vector<R> U2, V2;
        pd.Eval(U2, "tmpT");
        sl.template Eval<R, true>(V2, Xt, Xs, none, F);
        R e2 = 0, m2 = 0;
        for (Long i = 0; i < U2.Dim(); i++) { e2 = std::fmax(e2, std::fabs(U2[i] - V2[i])); m2 = std::fmax(m2, std::fabs(V2[i])); }
        CHECK(U2.Dim() == Nt * 3 && e2 <= tol * m2);
      }
      pd.DeleteSrc("tmp");
      pd.DeleteTrg("tmpT");
    }
    pd.Eval(U, "T");  // the original interactions are untouched
    err = 0;
    for (Long i = 0; i < U.Dim(); i++) err = std::fmax(err, std::fabs(U[i] - V[i]));
    CHECK(err <= tol * mag);
  }
  {  // Helmholtz through the session: the wavenumber is read through the kernel's ctx pointer at every Eval, as the
     // reference does (the kernel object is copied with its pointer, fmm-wrapper.txx:442-443)
    R mu = (R)3;
    sctl_cuda::Helmholtz3D_FxU hk;
    hk.SetCtxPtr(&mu);
    Vector<R> F2(Ns * 2), U1, U2, V1, V2, none;
    for (Long i = 0; i < Ns * 2; i++) F2[i] = F[i];
    sctl_cuda::ParticleDirect<R> pd;
    pd.SetAccuracy(-1);
    pd.AddSrc("H", hk, hk);
    pd.AddTrg("W", hk, hk);
    pd.SetKernelS2T("H", "W", hk);
    pd.SetSrcCoord("H", Xs);
    pd.SetSrcDensity("H", F2);
    pd.SetTrgCoord("W", Xt);
    pd.Eval(U1, "W");
    hk.template Eval<R, true>(V1, Xt, Xs, none, F2);
    mu = (R)7.5;  // no new SetKernelS2T
    pd.Eval(U2, "W");
    hk.template Eval<R, true>(V2, Xt, Xs, none, F2);
    R e1 = 0, e2 = 0, d12 = 0, mag = 0;
    for (Long i = 0; i < U1.Dim(); i++) {
      e1 = std::fmax(e1, std::fabs(U1[i] - V1[i]));
      e2 = std::fmax(e2, std::fabs(U2[i] - V2[i]));
      d12 = std::fmax(d12, std::fabs(V2[i] - V1[i]));
      mag = std::fmax(mag, std::fabs(V1[i]));
    }
    CHECK(e1 <= tol * mag && e2 <= tol * mag);
    CHECK(d12 > (R)1e-3 * mag);  // the two wavenumbers really give different fields
  }
  {  // two elements (300 and 400 nodes), near targets cut into blocks of 64
    using LV = Vector<Long>;
    LV elem_nds_cnt({300, 400}), elem_nds_dsp({0, 300}), near_elem_dsp({0, 200}), near_elem_cnt({200, 250}), K_near_dsp({0, 300 * 200});
    std::vector<Long> be, bt, bc;
    for (Long e = 0; e < 2; e++)
      for (Long t0 = 0; t0 < near_elem_cnt[e]; t0 += 64) { be.push_back(e); bt.push_back(t0); bc.push_back(std::min<Long>(64, near_elem_cnt[e] - t0)); }
    LV blk_e((Long)be.size()), blk_t((Long)be.size()), blk_c((Long)be.size());
    for (size_t i = 0; i < be.size(); i++) { blk_e[i] = be[i]; blk_t[i] = bt[i]; blk_c[i] = bc[i]; }
    std::unique_ptr<sctl_cuda::NearBlocks<R>> nbp(new sctl_cuda::NearBlocks<R>(3, 3, blk_e, blk_t, blk_c, elem_nds_cnt, elem_nds_dsp, elem_nds_dsp, near_elem_dsp, K_near_dsp));
    sctl_cuda::NearBlocks<R>& nb = *nbp;
    const Long K_len = (300 * 200 + 400 * 250) * 9;
    Vector<R> K(K_len), K2(K_len), Unear(450 * 3);
    for (Long i = 0; i < K_len; i++) K[i] = rnd();
    nb.SetMatrix(K);
    nb.SubtractDirect(sl, -1, Xs, Xn, W, Xt);
    nb.GetMatrix(K2);
    nb.Apply(Unear, F);
    R err = 0, mag = 0, uerr = 0, umag = 0;
    for (Long e = 0; e < 2; e++) {  // reference: one KernelMatrix per element, weights, transpose, subtract; then K * F
      Vector<R> xs(elem_nds_cnt[e] * 3, &Xs[elem_nds_dsp[e] * 3], false), xt(near_elem_cnt[e] * 3, &Xt[near_elem_dsp[e] * 3], false), none;
      Matrix<R> M;
      sl.template KernelMatrix<R, false>(M, xt, xs, none);
      const Long rows = near_elem_cnt[e] * 3, cols = elem_nds_cnt[e] * 3;
      for (Long i = 0; i < rows; i++) {
        R u = 0;
        for (Long j = 0; j < cols; j++) {
          const R want = K[K_near_dsp[e] * 9 + i * cols + j] - M(j, i) * W[elem_nds_dsp[e] + j / 3];
          const R got = K2[K_near_dsp[e] * 9 + i * cols + j];
          err = std::fmax(err, std::fabs(got - want));
          mag = std::fmax(mag, std::fabs(want));
          u += got * F[elem_nds_dsp[e] * 3 + j];
        }
        uerr = std::fmax(uerr, std::fabs(u - Unear[near_elem_dsp[e] * 3 + i]));
        umag = std::fmax(umag, std::fabs(u));
      }
    }
    CHECK(err <= tol * mag);
    CHECK(uerr <= (sizeof(R) == 8 ? (R)1e-12 : (R)1e-4) * umag);
    // bio_apply.hpp: far field (F * wts through the resident evaluator) + near correction (ApplyAccumulate) behind one call.
    // Every near pair belongs to its own operator target here; the scatter index reverses their order.
    LV cnt(Nt), dsp(Nt), sc(Nt);
    for (Long i = 0; i < Nt; i++) { cnt[i] = 1; dsp[i] = i; sc[i] = Nt - 1 - i; }
    nb.SetTargets(cnt, dsp, sc);
    Vector<R> FW(Ns * 3), Ufar, Ubio, none;
    for (Long i = 0; i < Ns * 3; i++) FW[i] = F[i] * W[i / 3];
    sl.template Eval<R, true>(Ufar, Xt, Xs, none, FW);
    sctl_cuda::BioApply<R> bio(sl, false, -1, Xs, Xn, W, Xt, none);
    CHECK(bio.SrcCount() == Ns && bio.TrgCount() == Nt && bio.TrgDof() == 3);
    bio.ComputePotential(Ubio, F);  // far field alone while no near session is attached
    R e0 = 0, e1 = 0, m1 = 0;
    for (Long i = 0; i < Nt * 3; i++) e0 = std::fmax(e0, std::fabs(Ubio[i] - Ufar[i]));
    bio.SetNear(std::move(nbp));
    bio.ComputePotential(Ubio, F);
    for (Long i = 0; i < Nt; i++)
      for (Long k = 0; k < 3; k++) {
        const R want = Ufar[i * 3 + k] + Unear[(Nt - 1 - i) * 3 + k];
        e1 = std::fmax(e1, std::fabs(Ubio[i * 3 + k] - want));
        m1 = std::fmax(m1, std::fabs(want));
      }
    CHECK(e0 <= tol * m1);
    CHECK(e1 <= tol * m1);
    // the pipeline (one upload, near GEMV beside the far field, one download) against the same apply staged through the two
    // sessions' host-array entries: same arithmetic, same order -> identical bits
    Vector<R> Ustaged;
    bio.ComputePotentialStaged(Ustaged, F);
    Long n_diff = 0;
    for (Long i = 0; i < Nt * 3; i++) n_diff += (Ustaged[i] != Ubio[i]);
    CHECK(n_diff == 0);
    double t_far = 0, t_tail = 0, t_all = 0;
    bio.LastTiming(t_far, t_tail, t_all);
    CHECK(t_all > 0 && t_far > 0);
  }
}

int main() {
  std::printf("devices: %d\n", sctl_cuda_device_count());
  generic_kernel_fixtures<double>();
  generic_kernel_fixtures<float>();
  kernel_function_fixtures<double>();
  kernel_function_fixtures<float>();
  session_layers<double>();
  session_layers<float>();
  std::printf("\n%s (%d check(s) failed)\n", test_utils::failures == 0 ? "PASS" : "FAIL", test_utils::failures);
  return test_utils::failures == 0 ? 0 : 1;
}
