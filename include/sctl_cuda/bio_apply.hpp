// sctl_cuda/bio_apply.hpp -- one object for the per-iteration part of sctl::BoundaryIntegralOp::ComputePotential
// (reference: include/sctl/boundary_integral.txx:716-722), built from the two device sessions:
//
//   ComputeFarField     (:1248-1309)  F_far = density at the far-field nodes * wts_far; direct sum over all far-field
//                                     nodes at every target (ParticleFMM::EvalDirect -> ParticleDirect, coordinates
//                                     resident in HBM); with trg_normal_dot_prod the 3*KDIM1_ target components are
//                                     contracted with the target normals (:1293-1303)
//   ComputeNearInterac  (:1311-1383)  block GEMVs with the resident K_near, ScatterForward, per-target accumulation
//                                     (NearBlocks::ApplyAccumulate)
//
// Setup (near lists, singular / near quadrature, K_near) stays where it is -- on the CPU in BoundaryIntegralOp, or
// NearBlocks::SubtractDirect for the "KNearSubtract" phase; this class is what a GMRES iteration calls.  A maintainer
// constructs it at the end of BoundaryIntegralOp::Setup from members that already exist;
// tests/cpp/test_near_blocks_ref.cpp does exactly that against the reference's own ComputePotential.
//
// ComputePotential runs the whole apply as one device-side pipeline (sctl_cuda_bio_apply, csrc/bio_session.inc): the
// density goes up once per GPU, F * wts and the normal contraction are device kernels, the near-field GEMV runs beside the
// far-field launches on its own stream, the shards of a multi-GPU far field are collected on the first GPU, U comes down
// once.  ComputePotentialStaged is the same apply from the two sessions' host-array entries (far field down, up again for
// the accumulation, down again) -- kept as the cross-check: without trg_normal_dot_prod the two are bit-identical.
// Works with any vector type that has Dim(), ReInit(n) and operator[] (sctl::Vector<Real>, sctl_cuda::Vector<Real>).
#ifndef SCTL_CUDA_BIO_APPLY_HPP_
#define SCTL_CUDA_BIO_APPLY_HPP_

#include <memory>
#include <vector>

#include "mini_containers.hpp"
#include "near_blocks.hpp"
#include "particle_direct.hpp"

namespace sctl_cuda {

template <class Real> class BioApply {
 public:
  BioApply(const BioApply&) = delete;
  BioApply& operator=(const BioApply&) = delete;

  // ker: the operator's kernel (sctl:: or sctl_cuda::); digits: BoundaryIntegralOp's accuracy in digits (:562);
  // X_far / Xn_far / wts_far: far-field nodes, normals, quadrature weights; Xtrg / Xn_trg: targets and, with
  // trg_normal_dot_prod, their normals.  n_gpus > 1 shards the targets of the far field over the node's GPUs.
  template <class Kernel, class Vec>
  BioApply(const Kernel& ker, bool trg_normal_dot_prod, Integer digits, const Vec& X_far, const Vec& Xn_far, const Vec& wts_far, const Vec& Xtrg, const Vec& Xn_trg, int n_gpus = 1)
      : far_(n_gpus), ndot_(trg_normal_dot_prod), k0_(Kernel::SrcDim()), k1_(Kernel::TrgDim()), Nsrc_(X_far.Dim() / 3), Ntrg_(Xtrg.Dim() / 3) {
    SCTL_CUDA_ASSERT(X_far.Dim() == Nsrc_ * 3 && Xtrg.Dim() == Ntrg_ * 3 && wts_far.Dim() == Nsrc_);
    SCTL_CUDA_ASSERT(!ndot_ || (k1_ % 3 == 0 && Xn_trg.Dim() == Ntrg_ * 3));
    wts_.assign(Nsrc_, Real(0));
    for (Long i = 0; i < Nsrc_; i++) wts_[i] = wts_far[i];
    if (ndot_) {
      nrm_.assign(Ntrg_ * 3, Real(0));
      for (Long i = 0; i < Ntrg_ * 3; i++) nrm_[i] = Xn_trg[i];
    }
    far_.SetAccuracy(digits);
    far_.AddSrc("Src", ker, ker);  // the names BoundaryIntegralOp uses (:573-575)
    far_.AddTrg("Trg", ker, ker);
    far_.SetKernelS2T("Src", "Trg", ker);
    far_.SetSrcCoord("Src", X_far, Xn_far);  // normals always travel, as in SetupFar
    far_.SetTrgCoord("Trg", Xtrg);
  }

  ~BioApply() { sctl_cuda_bio_destroy(bio_); }

  // Hands over the near-field session (operator set, SubtractDirect done or K_near uploaded after the CPU did it, SetTargets
  // called).  Without one, ComputePotential is the far field alone (element lists with dist_far = 0).
  void SetNear(std::unique_ptr<NearBlocks<Real>> near) {
    sctl_cuda_bio_destroy(bio_);  // the pipeline is rebuilt around the new near-field session at the next apply
    bio_ = nullptr;
    near_ = std::move(near);
  }
  NearBlocks<Real>* Near() { return near_.get(); }

  Long SrcCount() const { return Nsrc_; }
  Long TrgCount() const { return Ntrg_; }
  Integer TrgDof() const { return ndot_ ? k1_ / 3 : k1_; }  // KDIM1_ (:1290)

  // U = far field + near correction for the node density F.  F_far_nodes: the density at the far-field nodes where the
  // element list has a far-field density operator (GetFarFieldDensity, :1267); empty (the default of every element list in
  // the reference tree, :527-529) means the far-field nodes are the nodes and F itself is used (:1277-1283).
  template <class VecOut, class VecIn> void ComputePotential(VecOut& U, const VecIn& F, const VecIn& F_far_nodes = VecIn()) {
    SCTL_CUDA_ASSERT((F_far_nodes.Dim() ? F_far_nodes.Dim() : F.Dim()) == Nsrc_ * k0_);
    const Integer k1r = TrgDof();
    if (U.Dim() != Ntrg_ * k1r) U.ReInit(Ntrg_ * k1r);
    if (!bio_) {
      if (sctl_cuda_bio_create(far_.Handle(), far_.SrcSlot("Src"), far_.TrgSlot("Trg"), near_ ? near_->Handle() : nullptr, ndot_ ? 1 : 0, Nsrc_ ? wts_.data() : nullptr, Nsrc_,
                               ndot_ && Ntrg_ ? nrm_.data() : nullptr, ndot_ ? Ntrg_ * 3 : 0, &bio_))
        detail::die("sctl_cuda_bio_create");
    }
    far_.RefreshContexts("Trg");
    const Integer digits = far_.Accuracy();
    if (sctl_cuda_bio_apply(bio_, (int)(digits >= 16 ? -1 : digits), F.Dim() ? &F[0] : nullptr, F.Dim(), F_far_nodes.Dim() ? &F_far_nodes[0] : nullptr, F_far_nodes.Dim(),
                            U.Dim() ? &U[0] : nullptr, U.Dim()))
      detail::die("sctl_cuda_bio_apply");
  }

  // Device times of the last ComputePotential (ms): until the far field is complete on the first GPU, the tail behind it,
  // and the whole apply including the download of U.
  void LastTiming(double& far_ms, double& tail_ms, double& total_ms) const {
    far_ms = tail_ms = total_ms = 0;
    if (bio_) sctl_cuda_bio_last_timing(bio_, &far_ms, &tail_ms, &total_ms);
  }

  // Work items of the near-field GEMV that each GPU of the far-field session ran in the last ComputePotential (items[p] for
  // the p-th GPU; all zero: the near session's own GPU ran all of it).
  std::vector<int64_t> NearShards() const {
    std::vector<int64_t> items(16, 0);
    if (bio_) sctl_cuda_bio_near_shards(bio_, items.data(), (int)items.size());
    return items;
  }

  // The same apply through the two sessions' host-array entries, phase after phase (the cross-check of the pipeline above).
  template <class VecOut, class VecIn> void ComputePotentialStaged(VecOut& U, const VecIn& F, const VecIn& F_far_nodes = VecIn()) {
    const VecIn& Fsrc = F_far_nodes.Dim() ? F_far_nodes : F;
    SCTL_CUDA_ASSERT(Fsrc.Dim() == Nsrc_ * k0_);
    if (F_far_.Dim() != Nsrc_ * k0_) F_far_.ReInit(Nsrc_ * k0_);
    for (Long i = 0; i < Nsrc_; i++)
      for (Integer j = 0; j < k0_; j++) F_far_[i * k0_ + j] = Fsrc[i * k0_ + j] * wts_[i];  // F_far <-- F * wts_far (:1270-1283)
    far_.SetSrcDensity("Src", F_far_);
    const Integer k1r = TrgDof();
    if (U.Dim() != Ntrg_ * k1r) U.ReInit(Ntrg_ * k1r);
    if (ndot_) {
      far_.Eval(U3_, "Trg");
      for (Long i = 0; i < Ntrg_; i++)
        for (Integer k = 0; k < k1r; k++) {
          Real v = 0;  // U.SetZero() then += for l = 0, 1, 2 (:1291-1301)
          for (Integer l = 0; l < 3; l++) v += U3_[(i * k1r + k) * 3 + l] * nrm_[i * 3 + l];
          U[i * k1r + k] = v;
        }
    } else {
      far_.Eval(U, "Trg");  // overwrites, like U.SetZero() + fmm.Eval (:1291,1305)
    }
    if (near_) near_->ApplyAccumulate(U, F);
  }

 private:
  ParticleDirect<Real> far_;
  std::unique_ptr<NearBlocks<Real>> near_;
  sctl_cuda_bio_t bio_ = nullptr;  // borrows far_ and near_: destroyed before them
  bool ndot_;
  Integer k0_, k1_;
  Long Nsrc_, Ntrg_;
  std::vector<Real> wts_, nrm_;
  Vector<Real> F_far_, U3_;
};

}  // namespace sctl_cuda

#endif  // SCTL_CUDA_BIO_APPLY_HPP_
