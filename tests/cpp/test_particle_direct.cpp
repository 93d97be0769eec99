// tests/cpp/test_particle_direct.cpp -- sctl_cuda::ParticleDirect (device-resident, GPU) against the reference's own
// ParticleFMM<Real,3>::EvalDirect on the CPU (unmodified SCTL headers, NOT routed through the drop-in), driven through
// the same sequence of calls.  Built only where /root/reference is mounted; the binary travels to the GPU box.
//   usage: test_particle_direct [n_gpus]
#include <sctl.hpp>

#include <cstdio>
#include <cstdlib>

#include "sctl_cuda/particle_direct.hpp"

using namespace sctl;

static int n_fail = 0;
#define CHECK(cond, ...)                      \
  do {                                        \
    if (!(cond)) {                            \
      n_fail++;                               \
      std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
      std::printf(__VA_ARGS__);               \
      std::printf("\n");                      \
    }                                         \
  } while (0)

static uint64_t rng_state = 12345;
static double u01() {  // SplitMix64
  uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
template <class Real> static Vector<Real> rand_vec(Long n) {
  Vector<Real> v(n);
  for (auto& a : v) a = (Real)(u01() - 0.5);
  return v;
}
template <class Real> static double rel_err(const Vector<Real>& a, const Vector<Real>& b) {
  if (a.Dim() != b.Dim()) return 1e300;
  double e = 0, m = 0;
  for (Long i = 0; i < a.Dim(); i++) {
    e = std::max(e, std::fabs((double)a[i] - (double)b[i]));
    m = std::max(m, std::fabs((double)b[i]));
  }
  return m > 0 ? e / m : e;
}

// Two source types (single layer + double layer with normals) into one velocity target set, a second target set fed by
// one source type only, then a new density (the GMRES pattern), then changed coordinates.
template <class Real> static void run(int n_gpus, int digits, double tol) {
  const Stokes3D_FSxU ker_m2l;
  const Stokes3D_FxU ker_sl;
  const Stokes3D_DxU ker_dl;
  const Laplace3D_FxU lap_sl;
  const Laplace3D_FxdU lap_grad;
  const Long Nsl = 3001, Ndl = 1777, Nt1 = 2500, Nt2 = 333;
  Vector<Real> Xsl = rand_vec<Real>(Nsl * 3), Fsl = rand_vec<Real>(Nsl * 3);
  Vector<Real> Xdl = rand_vec<Real>(Ndl * 3), Ndl_ = rand_vec<Real>(Ndl * 3), Fdl = rand_vec<Real>(Ndl * 3);
  Vector<Real> Xt1 = rand_vec<Real>(Nt1 * 3), Xt2 = rand_vec<Real>(Nt2 * 3);

  ParticleFMM<Real, 3> ref;  // CPU checker: the reference's own direct path
  sctl_cuda::ParticleDirect<Real> gpu(n_gpus);
  ref.SetAccuracy(digits);
  gpu.SetAccuracy(digits);
  ref.SetKernels(ker_m2l, ker_m2l, ker_sl);
  gpu.SetKernels(ker_m2l, ker_m2l, ker_sl);
  ref.AddTrg("Velocity", ker_m2l, ker_sl);
  gpu.AddTrg("Velocity", ker_m2l, ker_sl);
  ref.AddTrg("Probe", ker_m2l, ker_sl);
  gpu.AddTrg("Probe", ker_m2l, ker_sl);
  ref.AddSrc("SingleLayer", ker_sl, ker_sl);
  gpu.AddSrc("SingleLayer", ker_sl, ker_sl);
  ref.AddSrc("DoubleLayer", ker_dl, ker_dl);
  gpu.AddSrc("DoubleLayer", ker_dl, ker_dl);
  ref.SetKernelS2T("SingleLayer", "Velocity", ker_sl);
  gpu.SetKernelS2T("SingleLayer", "Velocity", ker_sl);
  ref.SetKernelS2T("DoubleLayer", "Velocity", ker_dl);
  gpu.SetKernelS2T("DoubleLayer", "Velocity", ker_dl);
  ref.SetKernelS2T("DoubleLayer", "Probe", ker_dl);
  gpu.SetKernelS2T("DoubleLayer", "Probe", ker_dl);

  ref.SetTrgCoord("Velocity", Xt1);
  gpu.SetTrgCoord("Velocity", Xt1);
  ref.SetTrgCoord("Probe", Xt2);
  gpu.SetTrgCoord("Probe", Xt2);
  ref.SetSrcCoord("SingleLayer", Xsl);
  gpu.SetSrcCoord("SingleLayer", Xsl);
  ref.SetSrcCoord("DoubleLayer", Xdl, Ndl_);
  gpu.SetSrcCoord("DoubleLayer", Xdl, Ndl_);
  ref.SetSrcDensity("SingleLayer", Fsl);
  gpu.SetSrcDensity("SingleLayer", Fsl);
  ref.SetSrcDensity("DoubleLayer", Fdl);
  gpu.SetSrcDensity("DoubleLayer", Fdl);

  Vector<Real> Uref, Ugpu(7);  // wrong-sized U is resized; contents are overwritten, not accumulated into
  ref.EvalDirect(Uref, "Velocity");
  gpu.EvalDirect(Ugpu, "Velocity");
  CHECK(Ugpu.Dim() == Nt1 * 3, "Velocity size %ld", (long)Ugpu.Dim());
  CHECK(rel_err(Ugpu, Uref) <= tol, "Velocity (two source types) err %.3e", rel_err(Ugpu, Uref));
  gpu.Eval(Ugpu, "Velocity");  // again into the now-filled U: overwrite semantics
  CHECK(rel_err(Ugpu, Uref) <= tol, "Velocity second apply (overwrite) err %.3e", rel_err(Ugpu, Uref));

  Vector<Real> Pref, Pgpu;
  ref.EvalDirect(Pref, "Probe");
  gpu.EvalDirect(Pgpu, "Probe");
  CHECK(rel_err(Pgpu, Pref) <= tol, "Probe err %.3e", rel_err(Pgpu, Pref));

  for (int it = 0; it < 3; it++) {  // new densities, same coordinates
    Fdl = rand_vec<Real>(Ndl * 3);
    ref.SetSrcDensity("DoubleLayer", Fdl);
    gpu.SetSrcDensity("DoubleLayer", Fdl);
    ref.EvalDirect(Uref, "Velocity");
    gpu.EvalDirect(Ugpu, "Velocity");
    CHECK(rel_err(Ugpu, Uref) <= tol, "iteration %d err %.3e", it, rel_err(Ugpu, Uref));
  }

  Xt1 = rand_vec<Real>((Nt1 + 11) * 3);  // moved and resized target set
  ref.SetTrgCoord("Velocity", Xt1);
  gpu.SetTrgCoord("Velocity", Xt1);
  ref.EvalDirect(Uref, "Velocity");
  gpu.EvalDirect(Ugpu, "Velocity");
  CHECK(Ugpu.Dim() == (Nt1 + 11) * 3 && rel_err(Ugpu, Uref) <= tol, "moved targets err %.3e", rel_err(Ugpu, Uref));

  ref.DeleteSrc("SingleLayer");
  gpu.DeleteSrc("SingleLayer");
  ref.EvalDirect(Uref, "Velocity");
  gpu.EvalDirect(Ugpu, "Velocity");
  CHECK(rel_err(Ugpu, Uref) <= tol, "after DeleteSrc err %.3e", rel_err(Ugpu, Uref));

  // Laplace potential + gradient targets from one charge set (different target dof per target type)
  ParticleFMM<Real, 3> ref2;
  sctl_cuda::ParticleDirect<Real> gpu2(n_gpus);
  Vector<Real> Q = rand_vec<Real>(Nsl);
  ref2.SetAccuracy(digits);
  gpu2.SetAccuracy(digits);
  ref2.SetKernels(lap_sl, lap_sl, lap_sl);
  ref2.AddSrc("Charge", lap_sl, lap_sl);
  gpu2.AddSrc("Charge", lap_sl, lap_sl);
  ref2.AddTrg("Potential", lap_sl, lap_sl);
  gpu2.AddTrg("Potential", lap_sl, lap_sl);
  ref2.AddTrg("Gradient", lap_grad, lap_grad);
  gpu2.AddTrg("Gradient", lap_grad, lap_grad);
  ref2.SetKernelS2T("Charge", "Potential", lap_sl);
  gpu2.SetKernelS2T("Charge", "Potential", lap_sl);
  ref2.SetKernelS2T("Charge", "Gradient", lap_grad);
  gpu2.SetKernelS2T("Charge", "Gradient", lap_grad);
  ref2.SetSrcCoord("Charge", Xsl);
  gpu2.SetSrcCoord("Charge", Xsl);
  ref2.SetSrcDensity("Charge", Q);
  gpu2.SetSrcDensity("Charge", Q);
  ref2.SetTrgCoord("Potential", Xsl);  // targets == sources: self pairs contribute 0
  gpu2.SetTrgCoord("Potential", Xsl);
  ref2.SetTrgCoord("Gradient", Xt2);
  gpu2.SetTrgCoord("Gradient", Xt2);
  ref2.EvalDirect(Uref, "Potential");
  gpu2.EvalDirect(Ugpu, "Potential");
  CHECK(rel_err(Ugpu, Uref) <= tol, "Laplace potential (self) err %.3e", rel_err(Ugpu, Uref));
  ref2.EvalDirect(Uref, "Gradient");
  gpu2.EvalDirect(Ugpu, "Gradient");
  CHECK(Ugpu.Dim() == Nt2 * 3 && rel_err(Ugpu, Uref) <= tol, "Laplace gradient err %.3e", rel_err(Ugpu, Uref));
}

int main(int argc, char** argv) {
  const int n_gpus = argc > 1 ? std::atoi(argv[1]) : 1;
  // ParticleFMM's default accuracy is 10 digits (fmm-wrapper.txx:205): the device serves that request with its
  // 4-instruction reciprocal-square-root class (error < 9e-13 per pair, zero mean), the reference's AVX-512 path with full
  // precision, so the two agree to ~1e-11 there (kernels with rinv^5 amplify the per-pair error five-fold); at 15 digits
  // both are at working precision and the north-star gate 1e-12 applies.
  run<double>(n_gpus, 10, 2e-11);
  run<double>(n_gpus, 15, 1e-12);
  run<float>(n_gpus, 10, 1e-5);
  std::printf("%s (%d check(s) failed)\n", n_fail ? "FAIL" : "PASS", n_fail);
  return n_fail ? 1 : 0;
}
