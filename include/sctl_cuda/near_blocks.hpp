// sctl_cuda/near_blocks.hpp -- C++ host layer over sctl_cuda_near_* (include/sctl_cuda.h): the two near-field block
// phases of sctl::BoundaryIntegralOp on the device, operator resident in HBM.
//
//   SubtractDirect   "KNearSubtract"  (reference include/sctl/boundary_integral.txx:1165-1240)
//   Apply            "EvalNear"       (boundary_integral.txx:1324-1340)
//   ApplyAccumulate  "EvalNear" + the scatter / per-target accumulation that ends ComputeNearInterac (:1366-1381)
//
// The constructor takes BoundaryIntegralOp's own work-list vectors (near_blk_elem / near_blk_t0 / near_blk_cnt,
// :1009-1032) and per-element tables and flattens them into the per-block table of the C ABI, so a maintainer can
// construct it at the end of SetupNearSingular from the members that already exist.
// Works with any vector type that has Dim() and operator[] (sctl::Vector<T>, sctl_cuda::Vector<T>).
#ifndef SCTL_CUDA_NEAR_BLOCKS_HPP_
#define SCTL_CUDA_NEAR_BLOCKS_HPP_

#include <vector>

#include "generic_kernel.hpp"

namespace sctl_cuda {

template <class Real> class NearBlocks {
 public:
  NearBlocks(const NearBlocks&) = delete;
  NearBlocks& operator=(const NearBlocks&) = delete;

  // Tables as BoundaryIntegralOp holds them (all Vector<Long>): per block near_blk_elem, near_blk_t0, near_blk_cnt;
  // per element elem_nds_cnt, elem_nds_dsp, elem_nds_dsp_far, near_elem_dsp, K_near_dsp.
  template <class LongVec>
  NearBlocks(Integer KDIM0, Integer KDIM1, const LongVec& near_blk_elem, const LongVec& near_blk_t0, const LongVec& near_blk_cnt, const LongVec& elem_nds_cnt,
             const LongVec& elem_nds_dsp, const LongVec& elem_nds_dsp_far, const LongVec& near_elem_dsp, const LongVec& K_near_dsp)
      : h_(nullptr) {
    const Long n = near_blk_elem.Dim();
    SCTL_CUDA_ASSERT(near_blk_t0.Dim() == n && near_blk_cnt.Dim() == n);
    std::vector<int64_t> f_dsp(n), far_dsp(n), ns(n), u_dsp(n), nt(n), k_dsp(n);
    for (Long b = 0; b < n; b++) {
      const Long e = near_blk_elem[b], t0 = near_blk_t0[b];
      const Long src_dof = elem_nds_cnt[e] * KDIM0;           // :1326
      f_dsp[b] = elem_nds_dsp[e];                             // F_  = F + elem_nds_dsp[e]*KDIM0             (:1336)
      far_dsp[b] = elem_nds_dsp_far[e];                       // X   = X_far + elem_nds_dsp_far[e]*COORD_DIM (:1177)
      ns[b] = elem_nds_cnt[e];
      u_dsp[b] = near_elem_dsp[e] + t0;                       // U_  = U_near + (near_elem_dsp[e]+t0)*KDIM1   (:1337)
      nt[b] = near_blk_cnt[b];
      k_dsp[b] = K_near_dsp[e] * KDIM0 * KDIM1 + t0 * KDIM1 * src_dof;  // (:1183, :1335)
    }
    if (sctl_cuda_near_create(detail::real_bytes<Real>::value, (int)KDIM0, (int)KDIM1, n, f_dsp.data(), far_dsp.data(), ns.data(), u_dsp.data(), nt.data(), k_dsp.data(), &h_))
      detail::die("sctl_cuda_near_create");
  }
  ~NearBlocks() { sctl_cuda_near_destroy(h_); }

  // K_near as the CPU singular / near quadrature left it (BoundaryIntegralOp::K_near); stays in HBM afterwards
  template <class Vec> void SetMatrix(const Vec& K_near) {
    if (sctl_cuda_near_set_matrix(h_, K_near.Dim() ? &K_near[0] : nullptr, K_near.Dim())) detail::die("sctl_cuda_near_set_matrix");
  }
  template <class Vec> void GetMatrix(Vec& K_near) const {
    if (sctl_cuda_near_get_matrix(h_, K_near.Dim() ? &K_near[0] : nullptr, K_near.Dim())) detail::die("sctl_cuda_near_get_matrix");
  }

  // K_near_[i][j] -= wts[s] * KernelMatrix[j][i] for every block (element lists without a density operator); ker supplies the kernel id (by Name()) and its context pointer.
  // Xn_trg_near non-empty selects trg_normal_dot_prod (:1186-1200); the object is then built with KDIM1 = Kernel::TrgDim()/3.
  template <class Kernel, class Vec> void SubtractDirect(const Kernel& ker, Integer digits, const Vec& X_far, const Vec& Xn_far, const Vec& wts_far, const Vec& Xtrg_near, const Vec& Xn_trg_near = Vec()) {
    int id = -1;
    for (int k = 0;; k++) {
      const char* n = nullptr;
      if (sctl_cuda_kernel_info(k, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &n)) break;
      if (Kernel::Name() == n) id = k;
    }
    SCTL_CUDA_ASSERT(id >= 0);
    const void* ctx = ker.GetCtxPtr();
    if (sctl_cuda_near_subtract_direct(h_, id, (int)(digits >= 16 ? -1 : digits), X_far.Dim() / 3, X_far.Dim() ? &X_far[0] : nullptr, Xn_far.Dim() ? &Xn_far[0] : nullptr,
                                       wts_far.Dim() ? &wts_far[0] : nullptr, Xtrg_near.Dim() / 3, Xtrg_near.Dim() ? &Xtrg_near[0] : nullptr, Xn_trg_near.Dim() ? &Xn_trg_near[0] : nullptr, ctx,
                                       ctx ? (int64_t)sizeof(Real) : 0))
      detail::die("sctl_cuda_near_subtract_direct");
  }

  // U_near = blockwise K_near * F (overwrites; :1324-1340).  U_near must have near_elem_dsp.back()+near_elem_cnt.back() targets.
  template <class VecOut, class VecIn> void Apply(VecOut& U_near, const VecIn& F) const {
    if (sctl_cuda_near_apply(h_, F.Dim() ? &F[0] : nullptr, F.Dim(), U_near.Dim() ? &U_near[0] : nullptr, U_near.Dim())) detail::die("sctl_cuda_near_apply");
  }

  // Tables of ComputeNearInterac's tail as BoundaryIntegralOp holds them (Vector<Long>): near_trg_cnt, near_trg_dsp (one per
  // target of the operator) and near_scatter_index (one per near pair); one-process Comm (ScatterForward is a local gather).
  template <class LongVec> void SetTargets(const LongVec& near_trg_cnt, const LongVec& near_trg_dsp, const LongVec& near_scatter_index) {
    const Long n = near_trg_cnt.Dim(), m = near_scatter_index.Dim();
    SCTL_CUDA_ASSERT(near_trg_dsp.Dim() == n);
    std::vector<int64_t> cnt(n), dsp(n), sc(m);
    for (Long i = 0; i < n; i++) { cnt[i] = near_trg_cnt[i]; dsp[i] = near_trg_dsp[i]; }
    for (Long i = 0; i < m; i++) sc[i] = near_scatter_index[i];
    if (sctl_cuda_near_set_targets(h_, n, cnt.data(), dsp.data(), m, sc.data())) detail::die("sctl_cuda_near_set_targets");
  }

  // U += near-field correction, exactly the loops at :1324-1340 and :1366-1381 (U must already have Ntrg*KDIM1_ values, e.g.
  // the far field); only F and U cross PCIe.
  template <class VecOut, class VecIn> void ApplyAccumulate(VecOut& U, const VecIn& F) const {
    if (sctl_cuda_near_apply_accumulate(h_, F.Dim() ? &F[0] : nullptr, F.Dim(), U.Dim() ? &U[0] : nullptr, U.Dim())) detail::die("sctl_cuda_near_apply_accumulate");
  }

  sctl_cuda_near_t Handle() const { return h_; }

 private:
  sctl_cuda_near_t h_;
};

}  // namespace sctl_cuda

#endif  // SCTL_CUDA_NEAR_BLOCKS_HPP_
