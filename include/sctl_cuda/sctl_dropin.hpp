// sctl_cuda/sctl_dropin.hpp -- routes SCTL's own kernel objects to the GPU.
//
// Include this header after (or instead of) <sctl.hpp> -- e.g. `g++ -include sctl_cuda/sctl_dropin.hpp ...` --
// and link with -lsctl_cuda.  It explicitly specialises the member templates
//     sctl::GenericKernel<kernel_impl::X>::Eval<Real,omp,digits>         (generic-kernel.txx:102-211)
//     sctl::GenericKernel<kernel_impl::X>::KernelMatrix<Real,omp,digits> (generic-kernel.txx:213-329)
// for X in {Laplace3D_FxU, Laplace3D_DxU, Laplace3D_FxdU, Stokes3D_FxU, Stokes3D_DxU} (the north-star five) and
// {Stokes3D_FxT, Stokes3D_FSxU, Stokes3D_FxUP, Laplace3D_Fxd2U, Laplace3D_DxdU, BiotSavart3D_FxU, BiotSavart3D_FxdU,
//  Helmholtz3D_FxU, Helmholtz3D_DxU, Helmholtz3D_FxdU, HelmholtzDiff3D_FxdU} -- all sixteen kernels -- so every caller that
// goes through those entry points -- user code, the runtime-digits trampoline Eval<Real,omp>(..., digits, self)
// (generic-kernel.txx:72-100) that ParticleFMM stores (fmm-wrapper.txx:450-451), BoundaryIntegralOp's
// KernelMatrix calls (boundary_integral.txx:1197,1211) -- runs on the B200 with unchanged source.  SCTL's headers
// are not modified; uKerMatrix/uKerApply stay available
// to PVFMM-style callers.  Profile::IncrementCounter(FLOP, Ns*Nt*FLOPS()) is preserved (generic-kernel.txx:210,224).
#ifndef SCTL_CUDA_SCTL_DROPIN_HPP_
#define SCTL_CUDA_SCTL_DROPIN_HPP_

#include <sctl.hpp>

#include "generic_kernel.hpp"

namespace sctl {

#define SCTL_CUDA_DROPIN(UKER, ID, K0, K1, ND)                                                                                              \
  template <> template <class Real, bool enable_openmp, Integer digits>                                                                      \
  void GenericKernel<kernel_impl::UKER>::Eval(Vector<Real>& v_trg, const Vector<Real>& r_trg, const Vector<Real>& r_src,                     \
                                              const Vector<Real>& n_src, const Vector<Real>& v_src) const {                                  \
    ::sctl_cuda::detail::eval_host<Real>(ID, K0, K1, ND, digits, v_trg, r_trg, r_src, n_src, v_src, ctx_ptr);                                \
    Profile::IncrementCounter(ProfileCounter::FLOP, (r_src.Dim() / 3) * (r_trg.Dim() / 3) * kernel_impl::UKER::FLOPS());                     \
  }                                                                                                                                          \
  template <> template <class Real, bool enable_openmp, Integer digits>                                                                      \
  void GenericKernel<kernel_impl::UKER>::KernelMatrix(Matrix<Real>& M, const Vector<Real>& Xt, const Vector<Real>& Xs, const Vector<Real>& Xn) const { \
    ::sctl_cuda::detail::kernel_matrix_host<Real>(ID, K0, K1, ND, digits, M, Xt, Xs, Xn, ctx_ptr);                                                    \
    Profile::IncrementCounter(ProfileCounter::FLOP, (Xs.Dim() / 3) * (Xt.Dim() / 3) * kernel_impl::UKER::FLOPS());                           \
  }

SCTL_CUDA_DROPIN(Laplace3D_FxU, SCTL_CUDA_LAPLACE3D_FXU, 1, 1, 0)
SCTL_CUDA_DROPIN(Laplace3D_DxU, SCTL_CUDA_LAPLACE3D_DXU, 1, 1, 3)
SCTL_CUDA_DROPIN(Laplace3D_FxdU, SCTL_CUDA_LAPLACE3D_FXDU, 1, 3, 0)
SCTL_CUDA_DROPIN(Stokes3D_FxU, SCTL_CUDA_STOKES3D_FXU, 3, 3, 0)
SCTL_CUDA_DROPIN(Stokes3D_DxU, SCTL_CUDA_STOKES3D_DXU, 3, 3, 3)
SCTL_CUDA_DROPIN(Stokes3D_FxT, SCTL_CUDA_STOKES3D_FXT, 3, 9, 0)
SCTL_CUDA_DROPIN(Stokes3D_FSxU, SCTL_CUDA_STOKES3D_FSXU, 4, 3, 0)
SCTL_CUDA_DROPIN(Stokes3D_FxUP, SCTL_CUDA_STOKES3D_FXUP, 3, 4, 0)
SCTL_CUDA_DROPIN(Laplace3D_Fxd2U, SCTL_CUDA_LAPLACE3D_FXD2U, 1, 9, 0)
SCTL_CUDA_DROPIN(Laplace3D_DxdU, SCTL_CUDA_LAPLACE3D_DXDU, 1, 3, 3)
SCTL_CUDA_DROPIN(BiotSavart3D_FxU, SCTL_CUDA_BIOTSAVART3D_FXU, 3, 3, 0)
SCTL_CUDA_DROPIN(BiotSavart3D_FxdU, SCTL_CUDA_BIOTSAVART3D_FXDU, 3, 9, 0)
SCTL_CUDA_DROPIN(Helmholtz3D_FxU, SCTL_CUDA_HELMHOLTZ3D_FXU, 2, 2, 0)
SCTL_CUDA_DROPIN(Helmholtz3D_DxU, SCTL_CUDA_HELMHOLTZ3D_DXU, 2, 2, 3)
SCTL_CUDA_DROPIN(Helmholtz3D_FxdU, SCTL_CUDA_HELMHOLTZ3D_FXDU, 2, 6, 0)
SCTL_CUDA_DROPIN(HelmholtzDiff3D_FxdU, SCTL_CUDA_HELMHOLTZDIFF3D_FXDU, 2, 6, 0)
#undef SCTL_CUDA_DROPIN

}  // namespace sctl

#endif  // SCTL_CUDA_SCTL_DROPIN_HPP_
