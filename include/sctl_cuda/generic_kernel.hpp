// sctl_cuda/generic_kernel.hpp -- header-only C++ host layer over the C ABI (include/sctl_cuda.h).
//
// Mirrors the public surface of sctl::GenericKernel<uKernel> (reference: include/sctl/generic-kernel.hpp:48-136)
// and of the five in-scope micro-kernel structs (include/sctl/kernel_functions.hpp:32-218, aliases :669-681):
// same member names, template parameters, argument order and meaning, same AoS Vector<Real> layouts, the same
// "accumulate unless v_trg has the wrong size" rule (generic-kernel.txx:129-132) and the same abort-on-violated-
// precondition behaviour (SCTL_ASSERT, common.hpp:98-105).  It is container-agnostic: any vector type with
// Dim(), ReInit(n), SetZero() and operator[] (sctl::Vector<Real>, or sctl_cuda::Vector<Real> from
// mini_containers.hpp when SCTL itself is not on the include path) and any matrix type with Dim(0/1),
// ReInit(r,c), SetZero() and operator[](row) works.
//
// With SCTL present, include sctl_cuda/sctl_dropin.hpp instead: it routes sctl::Laplace3D_FxU ... themselves here.
// Link with -lsctl_cuda.  No CPU fallback: without a GPU every Eval/KernelMatrix call aborts with the C ABI's message.
#ifndef SCTL_CUDA_GENERIC_KERNEL_HPP_
#define SCTL_CUDA_GENERIC_KERNEL_HPP_

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../sctl_cuda.h"

namespace sctl_cuda {

using Integer = long;     // sctl::Integer (common.hpp:50)
using Long = int64_t;     // sctl::Long    (common.hpp:51)

#define SCTL_CUDA_ASSERT(cond)                                                                                   \
  do {                                                                                                           \
    if (!(cond)) {                                                                                               \
      std::fprintf(stderr, "Error: assertion `%s` failed in %s:%d (%s)\n", #cond, __FILE__, __LINE__, __func__); \
      std::abort();                                                                                              \
    }                                                                                                            \
  } while (0)

namespace detail {

[[noreturn]] inline void die(const char* what) {
  std::fprintf(stderr, "Error: %s: %s\n", what, sctl_cuda_last_error());
  std::abort();
}

template <class Real> struct real_bytes;
template <> struct real_bytes<double> { static constexpr int value = 8; };
template <> struct real_bytes<float> { static constexpr int value = 4; };

template <class V> inline auto data_ptr(V& v) -> decltype(&v[0]) { return v.Dim() ? &v[0] : nullptr; }

// Body of GenericKernel<uKernel>::Eval<Real,omp,digits> (generic-kernel.txx:102-211) on top of the C ABI.
template <class Real, class VecOut, class VecIn>
inline void eval_host(int kernel_id, Integer K0, Integer K1, Integer ND, Integer digits, VecOut& v_trg, const VecIn& r_trg, const VecIn& r_src,
                      const VecIn& n_src, const VecIn& v_src, const void* ctx_ptr) {
  constexpr Integer DIM = 3;
  const Long Ns = r_src.Dim() / DIM;
  const Long Nt = r_trg.Dim() / DIM;
  SCTL_CUDA_ASSERT(r_trg.Dim() == Nt * DIM);
  SCTL_CUDA_ASSERT(r_src.Dim() == Ns * DIM);
  SCTL_CUDA_ASSERT(v_src.Dim() == Ns * K0);
  SCTL_CUDA_ASSERT(n_src.Dim() == Ns * ND || !ND);
  if (v_trg.Dim() != Nt * K1) {
    v_trg.ReInit(Nt * K1);
    v_trg.SetZero();
  }
  if (Ns == 0 || Nt == 0) return;
  const Real* xn = ND ? &n_src[0] : nullptr;
  if (sctl_cuda_eval(kernel_id, real_bytes<Real>::value, (int)(digits >= 16 ? -1 : digits), Ns, Nt, &r_src[0], xn, &v_src[0], &r_trg[0], &v_trg[0],
                     ctx_ptr, ctx_ptr ? (int64_t)sizeof(Real) : 0))
    die("sctl_cuda_eval");
}

// Body of GenericKernel<uKernel>::KernelMatrix<Real,omp,digits> (generic-kernel.txx:213-329).
template <class Real, class Mat, class VecIn>
inline void kernel_matrix_host(int kernel_id, Integer K0, Integer K1, Integer ND, Integer digits, Mat& M, const VecIn& Xt, const VecIn& Xs, const VecIn& Xn,
                               const void* ctx_ptr = nullptr) {
  constexpr Integer DIM = 3;
  const Long Ns = Xs.Dim() / DIM;
  const Long Nt = Xt.Dim() / DIM;
  SCTL_CUDA_ASSERT(Xt.Dim() == Nt * DIM);
  SCTL_CUDA_ASSERT(Xs.Dim() == Ns * DIM);
  SCTL_CUDA_ASSERT(Xn.Dim() == Ns * ND || !ND);
  if (M.Dim(0) != Ns * K0 || M.Dim(1) != Nt * K1) {
    M.ReInit(Ns * K0, Nt * K1);
    M.SetZero();
  }
  if (Ns == 0 || Nt == 0) return;
  const Real* xn = ND ? &Xn[0] : nullptr;
  if (sctl_cuda_kernel_matrix(kernel_id, real_bytes<Real>::value, (int)(digits >= 16 ? -1 : digits), Ns, Nt, &Xs[0], xn, &Xt[0], &M[0][0], Nt * K1,
                              ctx_ptr, ctx_ptr ? (int64_t)sizeof(Real) : 0))
    die("sctl_cuda_kernel_matrix");
}

}  // namespace detail

// Micro-kernel descriptors: what the reference deduces from each uKerMatrix signature (generic-kernel.hpp:22-41)
// plus Name()/FLOPS()/uKerScaleFactor<Real>().  The arithmetic itself lives on the device (csrc/pair_math.cuh).
namespace kernel_impl {

#define SCTL_CUDA_UKERNEL(STRUCT, ID, NAME, K0_, K1_, ND_, FLOPS_, NUM, DEN)                      \
  struct STRUCT {                                                                                 \
    static constexpr int kernel_id = ID;                                                          \
    static constexpr Integer KDIM0 = K0_, KDIM1 = K1_, N_DIM = ND_;                               \
    static const std::string& Name() {                                                            \
      static const std::string name = NAME;                                                       \
      return name;                                                                                \
    }                                                                                             \
    static constexpr Integer FLOPS() { return FLOPS_; }                                           \
    template <class Real> static constexpr Real uKerScaleFactor() {                               \
      return (Real)(NUM) / ((Real)(DEN) * (Real)3.14159265358979323846264338327950288L);          \
    }                                                                                             \
  }

SCTL_CUDA_UKERNEL(Laplace3D_FxU, SCTL_CUDA_LAPLACE3D_FXU, "Laplace3D-FxU", 1, 1, 0, 6, 1, 4);      // kernel_functions.hpp:32-47
SCTL_CUDA_UKERNEL(Laplace3D_DxU, SCTL_CUDA_LAPLACE3D_DXU, "Laplace3D-DxU", 1, 1, 3, 14, 1, 4);     // :61-79
SCTL_CUDA_UKERNEL(Laplace3D_FxdU, SCTL_CUDA_LAPLACE3D_FXDU, "Laplace3D-FxdU", 1, 3, 0, 11, -1, 4); // :81-110
SCTL_CUDA_UKERNEL(Stokes3D_FxU, SCTL_CUDA_STOKES3D_FXU, "Stokes3D-FxU", 3, 3, 0, 23, 1, 8);        // :125-176
SCTL_CUDA_UKERNEL(Stokes3D_DxU, SCTL_CUDA_STOKES3D_DXU, "Stokes3D-DxU", 3, 3, 3, 26, 3, 4);        // :180-217
SCTL_CUDA_UKERNEL(Stokes3D_FxT, SCTL_CUDA_STOKES3D_FXT, "Stokes3D-FxT", 3, 9, 0, 39, -3, 4);       // :220-254
SCTL_CUDA_UKERNEL(Stokes3D_FSxU, SCTL_CUDA_STOKES3D_FSXU, "Stokes3D-FSxU", 4, 3, 0, 26, 1, 8);     // :256-324
SCTL_CUDA_UKERNEL(Stokes3D_FxUP, SCTL_CUDA_STOKES3D_FXUP, "Stokes3D-FxUP", 3, 4, 0, 26, 1, 8);     // :326-364
SCTL_CUDA_UKERNEL(Laplace3D_Fxd2U, SCTL_CUDA_LAPLACE3D_FXD2U, "Laplace3D-Fxd2U", 1, 9, 0, 25, 1, 4);  // :366-389
SCTL_CUDA_UKERNEL(Laplace3D_DxdU, SCTL_CUDA_LAPLACE3D_DXDU, "Laplace3D-DxdU", 1, 3, 3, 22, 1, 4);     // :391-412
SCTL_CUDA_UKERNEL(BiotSavart3D_FxU, SCTL_CUDA_BIOTSAVART3D_FXU, "BiotSavart3D-FxU", 3, 3, 0, 15, 1, 4);      // :414-448
SCTL_CUDA_UKERNEL(BiotSavart3D_FxdU, SCTL_CUDA_BIOTSAVART3D_FXDU, "BiotSavart3D-FxdU", 3, 9, 0, 51, -1, 4);  // :450-483
// Helmholtz: SetCtxPtr(&mu) with mu of the evaluation's Real type (kernel_functions.hpp:484-488)
SCTL_CUDA_UKERNEL(Helmholtz3D_FxU, SCTL_CUDA_HELMHOLTZ3D_FXU, "Helmholtz3D-FxU", 2, 2, 0, 20, 1, 4);               // :490-533
SCTL_CUDA_UKERNEL(Helmholtz3D_DxU, SCTL_CUDA_HELMHOLTZ3D_DXU, "Helmholtz3D-DxU", 2, 2, 3, 29, -1, 4);              // :535-578
SCTL_CUDA_UKERNEL(Helmholtz3D_FxdU, SCTL_CUDA_HELMHOLTZ3D_FXDU, "Helmholtz3D-FxdU", 2, 6, 0, 35, 1, 4);            // :580-627
SCTL_CUDA_UKERNEL(HelmholtzDiff3D_FxdU, SCTL_CUDA_HELMHOLTZDIFF3D_FXDU, "HelmholtzDiff3D-FxdU", 2, 6, 0, 41, 1, 4); // :629-659
#undef SCTL_CUDA_UKERNEL

}  // namespace kernel_impl

template <class uKernel> class GenericKernel : public uKernel {
  static constexpr Integer DIM = 3;
  static constexpr Integer KDIM0 = uKernel::KDIM0;
  static constexpr Integer KDIM1 = uKernel::KDIM1;
  static constexpr Integer N_DIM = uKernel::N_DIM;

 public:
  GenericKernel() : ctx_ptr(nullptr) {}

  static constexpr Integer CoordDim() { return DIM; }
  static constexpr Integer NormalDim() { return N_DIM; }
  static constexpr Integer SrcDim() { return KDIM0; }
  static constexpr Integer TrgDim() { return KDIM1; }

  void SetCtxPtr(void* ctx) { ctx_ptr = ctx; }
  const void* GetCtxPtr() const { return ctx_ptr; }

  // Runtime-digits trampoline: the fn-pointer + opaque-self pair ParticleFMM stores (generic-kernel.hpp:99,
  // generic-kernel.txx:72-100, fmm-wrapper.txx:450-451).  digits >= 16 means working precision.
  template <class Real, bool enable_openmp, class VecOut, class VecIn>
  static void Eval(VecOut& v_trg, const VecIn& r_trg, const VecIn& r_src, const VecIn& n_src, const VecIn& v_src, Integer digits, const char* self) {
    const GenericKernel* ker = reinterpret_cast<const GenericKernel*>(self);
    detail::eval_host<Real>(uKernel::kernel_id, KDIM0, KDIM1, N_DIM, digits, v_trg, r_trg, r_src, n_src, v_src, ker ? ker->ctx_ptr : nullptr);
  }

  // v_trg += scale * sum_s K(x_t - x_s, n_s) v_s   (generic-kernel.hpp:112).  enable_openmp has no GPU meaning.
  template <class Real, bool enable_openmp = false, Integer digits = -1, class VecOut, class VecIn>
  void Eval(VecOut& v_trg, const VecIn& r_trg, const VecIn& r_src, const VecIn& n_src, const VecIn& v_src) const {
    detail::eval_host<Real>(uKernel::kernel_id, KDIM0, KDIM1, N_DIM, digits, v_trg, r_trg, r_src, n_src, v_src, ctx_ptr);
  }

  // M[(s*K0+k0)][(t*K1+k1)] = scale * K_{k0,k1}(x_t - x_s, n_s), overwriting   (generic-kernel.hpp:124).
  template <class Real, bool enable_openmp = false, Integer digits = -1, class Mat, class VecIn>
  void KernelMatrix(Mat& M, const VecIn& Xt, const VecIn& Xs, const VecIn& Xn) const {
    detail::kernel_matrix_host<Real>(uKernel::kernel_id, KDIM0, KDIM1, N_DIM, digits, M, Xt, Xs, Xn, ctx_ptr);
  }

 private:
  void* ctx_ptr;
};

using Laplace3D_FxU = GenericKernel<kernel_impl::Laplace3D_FxU>;
using Laplace3D_DxU = GenericKernel<kernel_impl::Laplace3D_DxU>;
using Laplace3D_FxdU = GenericKernel<kernel_impl::Laplace3D_FxdU>;
using Stokes3D_FxU = GenericKernel<kernel_impl::Stokes3D_FxU>;
using Stokes3D_DxU = GenericKernel<kernel_impl::Stokes3D_DxU>;
using Stokes3D_FxT = GenericKernel<kernel_impl::Stokes3D_FxT>;
using Stokes3D_FSxU = GenericKernel<kernel_impl::Stokes3D_FSxU>;
using Stokes3D_FxUP = GenericKernel<kernel_impl::Stokes3D_FxUP>;
using Laplace3D_Fxd2U = GenericKernel<kernel_impl::Laplace3D_Fxd2U>;
using Laplace3D_DxdU = GenericKernel<kernel_impl::Laplace3D_DxdU>;
using BiotSavart3D_FxU = GenericKernel<kernel_impl::BiotSavart3D_FxU>;
using BiotSavart3D_FxdU = GenericKernel<kernel_impl::BiotSavart3D_FxdU>;
using Helmholtz3D_FxU = GenericKernel<kernel_impl::Helmholtz3D_FxU>;
using Helmholtz3D_DxU = GenericKernel<kernel_impl::Helmholtz3D_DxU>;
using Helmholtz3D_FxdU = GenericKernel<kernel_impl::Helmholtz3D_FxdU>;
using HelmholtzDiff3D_FxdU = GenericKernel<kernel_impl::HelmholtzDiff3D_FxdU>;

}  // namespace sctl_cuda

#endif  // SCTL_CUDA_GENERIC_KERNEL_HPP_
