// BASELINE config 5: the reference's UNMODIFIED BoundaryIntegralOp<double, Laplace3D_FxU>::ComputePotential on a
// synthetic torus discretisation: Nu x Nv trapezoid nodes, elements = 8 x 8-node patches,
// far-field nodes = the nodes themselves with trapezoid weights and dist_far = 0, base-class Self/Near defaults
// (boundary_integral.txx:536-546), so the near list is empty and the whole cost is the far-field Eval through
// ParticleFMM::EvalDirect (fmm-wrapper.txx:551-623) and the runtime-digits trampoline, digits = 11.
//
// Built twice by tests/cpp/Makefile from this one source against the reference headers:
//   ref-bio-torus-cpu     plain reference (CPU)                 -- the checker
//   ref-bio-torus-dropin  with -include sctl_cuda/sctl_dropin.hpp -- the same caller, far field on the B200
// Usage: <exe> Nu Nv out.bin   -> writes U (double[Nu*Nv]) and prints the ComputePotential wall time.
#include <sctl.hpp>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

using namespace sctl;

template <class Real> class TorusPatchList : public ElementListBase<Real> {
 public:
  TorusPatchList(Long Nu = 0, Long Nv = 0) : Nu_(Nu), Nv_(Nv) {
    const Real R = 1, a = (Real)0.25, two_pi = 2 * const_pi<Real>();
    const Long N = Nu * Nv;
    X_.ReInit(N * 3);
    Xn_.ReInit(N * 3);
    w_.ReInit(N);
    // element-major node order: patches of 8 x 8 nodes
    Long idx = 0;
    for (Long pu = 0; pu < Nu / 8; pu++)
      for (Long pv = 0; pv < Nv / 8; pv++)
        for (Long iu = 0; iu < 8; iu++)
          for (Long iv = 0; iv < 8; iv++, idx++) {
            const Real u = (pu * 8 + iu) * two_pi / Nu, v = (pv * 8 + iv) * two_pi / Nv;
            const Real rho = R + a * std::cos(v);
            X_[idx * 3 + 0] = rho * std::cos(u); X_[idx * 3 + 1] = rho * std::sin(u); X_[idx * 3 + 2] = a * std::sin(v);
            Xn_[idx * 3 + 0] = std::cos(v) * std::cos(u); Xn_[idx * 3 + 1] = std::cos(v) * std::sin(u); Xn_[idx * 3 + 2] = std::sin(v);
            w_[idx] = a * rho * (two_pi / Nu) * (two_pi / Nv);
          }
  }
  Long Size() const override { return (Nu_ / 8) * (Nv_ / 8); }
  void GetNodeCoord(Vector<Real>* X, Vector<Real>* Xn, Vector<Long>* cnt) const override {
    if (X) *X = X_;
    if (Xn) *Xn = Xn_;
    if (cnt) { cnt->ReInit(Size()); for (Long i = 0; i < Size(); i++) (*cnt)[i] = 64; }
  }
  void GetFarFieldNodes(Vector<Real>& X, Vector<Real>& Xn, Vector<Real>& wts, Vector<Real>& dist_far, Vector<Long>& cnt, const Real tol) const override {
    X = X_; Xn = Xn_; wts = w_;
    dist_far.ReInit(w_.Dim()); dist_far.SetZero();
    cnt.ReInit(Size()); for (Long i = 0; i < Size(); i++) cnt[i] = 64;
  }
  const Vector<Real>& Nodes() const { return X_; }
 private:
  Long Nu_, Nv_;
  Vector<Real> X_, Xn_, w_;
};

int main(int argc, char** argv) {
  const Long Nu = argc > 1 ? atol(argv[1]) : 128, Nv = argc > 2 ? atol(argv[2]) : 64;
  const char* out = argc > 3 ? argv[3] : nullptr;
  TorusPatchList<double> elems(Nu, Nv);
  const Long N = Nu * Nv;
  Vector<double> F(N), U;
  {  // density sigma = 1 + 0.5 sin(3u) cos(2v) in element-major node order, recovered from the coordinates
    const Vector<double>& X = elems.Nodes();
    for (Long i = 0; i < N; i++) {
      const double u = std::atan2(X[i * 3 + 1], X[i * 3 + 0]);
      const double rho = std::sqrt(X[i * 3 + 0] * X[i * 3 + 0] + X[i * 3 + 1] * X[i * 3 + 1]);
      const double v = std::atan2(X[i * 3 + 2], rho - 1.0);
      F[i] = 1 + 0.5 * std::sin(3 * u) * std::cos(2 * v);
    }
  }
  Laplace3D_FxU ker;
  BoundaryIntegralOp<double, Laplace3D_FxU> op(ker);
  op.SetAccuracy(1e-10);  // digits = (Integer)(log(tol)/log(0.1)) + 1 = 11 (boundary_integral.txx:562)
  op.AddElemList(elems);
  op.ComputePotential(U, F);  // includes Setup
  const auto t0 = std::chrono::steady_clock::now();
  op.ComputePotential(U, F);
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  double umax = 0, usum = 0;
  for (Long i = 0; i < U.Dim(); i++) { umax = std::max(umax, std::fabs(U[i])); usum += U[i]; }
  std::printf("nodes %ld  U.Dim %ld  ComputePotential %.6f s  %.3f Gpair/s  max|U| %.15e  sum U %.15e\n", (long)N, (long)U.Dim(), sec, (double)N * N / sec * 1e-9, umax, usum);
  if (out) {
    FILE* f = std::fopen(out, "wb");
    if (!f) return 2;
    std::fwrite(&U[0], sizeof(double), U.Dim(), f);
    std::fclose(f);
  }
  return U.Dim() == N ? 0 : 1;
}
