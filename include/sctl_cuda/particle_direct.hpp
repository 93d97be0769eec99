// sctl_cuda/particle_direct.hpp -- device-resident direct evaluator with ParticleFMM's interface.
//
// Mirrors the subset of sctl::ParticleFMM<Real,DIM> (reference: include/sctl/fmm-wrapper.hpp:35-200) that its direct
// path uses -- SetAccuracy, AddSrc, AddTrg, SetKernelS2T, DeleteSrc, DeleteTrg, SetSrcCoord, SetSrcDensity,
// SetTrgCoord, Eval, EvalDirect (fmm-wrapper.txx:551-623) -- with the same names, argument order and meaning, on top
// of sctl_cuda_direct_* (include/sctl_cuda.h).  What changes is where the data lives: SetSrcCoord / SetTrgCoord
// upload once and the coordinates stay in HBM; an Eval ships only the density in and the potential out, which is
// the access pattern of BoundaryIntegralOp::ComputeFarField (boundary_integral.txx:1286-1305: SetSrcDensity + Eval
// per GMRES iteration, coordinates fixed since Setup).  Kernel arguments may be sctl:: kernels (with SCTL on the
// include path) or the sctl_cuda:: mirrors: they are identified by Name().  Works with any vector type that has
// Dim(), ReInit(n) and operator[] (sctl::Vector<Real>, sctl_cuda::Vector<Real>).
//
// Differences from ParticleFMM: no Comm (one process drives n_gpus GPUs of the node), no periodicity, no PVFMM; the
// FMM translation kernels given to SetKernels/AddSrc/AddTrg are accepted and ignored except for their dimensions.
// Devices: ParticleDirect(1) lives on the device that is current on the constructing thread (one rank per GPU: every
// rank's evaluator stays on the rank's own GPU); ParticleDirect(n > 1) takes devices 0..n-1; ParticleDirect(devices)
// takes an explicit list.
// Kernel context (the Helmholtz wavenumber): like the reference, which copies the kernel object with its ctx_ptr
// (fmm-wrapper.txx:442-443) and dereferences the pointer at every Eval, the pointer given to SetKernelS2T is kept and
// read again at every EvalDirect -- changing the wavenumber between two evaluations takes effect without a new
// SetKernelS2T.  The pointee must therefore outlive the interaction, exactly as for ParticleFMM.
#ifndef SCTL_CUDA_PARTICLE_DIRECT_HPP_
#define SCTL_CUDA_PARTICLE_DIRECT_HPP_

#include <map>
#include <string>
#include <vector>

#include "generic_kernel.hpp"

namespace sctl_cuda {

namespace detail {
inline int kernel_id_from_name(const std::string& name) {
  for (int id = 0;; id++) {
    const char* n = nullptr;
    if (sctl_cuda_kernel_info(id, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &n)) return -1;
    if (n && name == n) return id;
  }
}
}  // namespace detail

template <class Real, Integer DIM = 3> class ParticleDirect {
  static_assert(DIM == 3, "the device kernels are three-dimensional");

 public:
  ParticleDirect(const ParticleDirect&) = delete;
  ParticleDirect& operator=(const ParticleDirect&) = delete;

  // n_gpus > 1 shards every target set by contiguous block over GPUs 0..n_gpus-1 (results agree to rounding for any n_gpus)
  explicit ParticleDirect(int n_gpus = 1) : h_(nullptr), digits_(10) {  // default accuracy: fmm-wrapper.txx:205
    if (sctl_cuda_direct_create(detail::real_bytes<Real>::value, n_gpus, &h_)) detail::die("sctl_cuda_direct_create");
  }
  explicit ParticleDirect(const std::vector<int>& devices) : h_(nullptr), digits_(10) {
    if (sctl_cuda_direct_create_on(detail::real_bytes<Real>::value, (int)devices.size(), devices.data(), &h_)) detail::die("sctl_cuda_direct_create_on");
  }
  ~ParticleDirect() { sctl_cuda_direct_destroy(h_); }

  void SetAccuracy(Integer digits) { digits_ = digits; }

  template <class KerM2M, class KerM2L, class KerL2L> void SetKernels(const KerM2M&, const KerM2L&, const KerL2L&) {}

  template <class KerS2M, class KerS2L> void AddSrc(const std::string& name, const KerS2M&, const KerS2L&) {
    SCTL_CUDA_ASSERT(src_.find(name) == src_.end());  // "Source name already exists." (fmm-wrapper.txx)
    src_[name] = take_slot(free_src_, next_src_);
  }
  template <class KerM2T, class KerL2T> void AddTrg(const std::string& name, const KerM2T&, const KerL2T&) {
    SCTL_CUDA_ASSERT(trg_.find(name) == trg_.end());
    Trg t;
    t.slot = take_slot(free_trg_, next_trg_);
    t.dim = KerM2T::TrgDim();
    t.Nt = 0;
    trg_[name] = t;
  }

  template <class KerS2T> void SetKernelS2T(const std::string& src_name, const std::string& trg_name, const KerS2T& ker_s2t) {
    SCTL_CUDA_ASSERT(src_.find(src_name) != src_.end());
    SCTL_CUDA_ASSERT(trg_.find(trg_name) != trg_.end());
    SCTL_CUDA_ASSERT(KerS2T::TrgDim() == trg_[trg_name].dim);  // CheckKernelDims (fmm-wrapper.txx:627-)
    const int id = detail::kernel_id_from_name(KerS2T::Name());
    SCTL_CUDA_ASSERT(id >= 0);
    S2T e;
    e.src = src_[src_name];
    e.trg = trg_[trg_name].slot;
    e.kid = id;
    e.ctx = ker_s2t.GetCtxPtr();
    push_s2t(e);
    drop_s2t(e.src, e.trg);
    if (e.ctx) s2t_ctx_.push_back(e);  // re-read at every EvalDirect
  }

  void DeleteSrc(const std::string& name) {
    SCTL_CUDA_ASSERT(src_.find(name) != src_.end());
    const int slot = src_[name];
    for (auto& t : trg_) {
      sctl_cuda_direct_set_kernel_s2t(h_, slot, t.second.slot, -1, nullptr, 0);
      drop_s2t(slot, t.second.slot);
    }
    sctl_cuda_direct_set_src_coord(h_, slot, 0, nullptr, nullptr, 0);
    sctl_cuda_direct_set_src_density(h_, slot, 0, nullptr);
    src_.erase(name);
    free_src_.push_back(slot);  // slot numbers are bounded by the C ABI: reuse them
  }
  void DeleteTrg(const std::string& name) {
    SCTL_CUDA_ASSERT(trg_.find(name) != trg_.end());
    const int slot = trg_[name].slot;
    for (auto& s : src_) {
      sctl_cuda_direct_set_kernel_s2t(h_, s.second, slot, -1, nullptr, 0);
      drop_s2t(s.second, slot);
    }
    sctl_cuda_direct_set_trg_coord(h_, slot, 0, nullptr);
    trg_.erase(name);
    free_trg_.push_back(slot);
  }

  template <class Vec> void SetSrcCoord(const std::string& name, const Vec& src_coord, const Vec& src_normal) {
    SCTL_CUDA_ASSERT(src_.find(name) != src_.end());
    const Long Ns = src_coord.Dim() / DIM;
    SCTL_CUDA_ASSERT(src_coord.Dim() == Ns * DIM);
    SCTL_CUDA_ASSERT(src_normal.Dim() == 0 || src_normal.Dim() == Ns * DIM);
    const int nd = src_normal.Dim() ? (int)DIM : 0;
    if (sctl_cuda_direct_set_src_coord(h_, src_[name], Ns, Ns ? &src_coord[0] : nullptr, nd ? &src_normal[0] : nullptr, nd)) detail::die("sctl_cuda_direct_set_src_coord");
  }
  template <class Vec> void SetSrcCoord(const std::string& name, const Vec& src_coord) { SetSrcCoord(name, src_coord, Vec()); }

  template <class Vec> void SetSrcDensity(const std::string& name, const Vec& src_density) {
    SCTL_CUDA_ASSERT(src_.find(name) != src_.end());
    if (sctl_cuda_direct_set_src_density(h_, src_[name], src_density.Dim(), src_density.Dim() ? &src_density[0] : nullptr)) detail::die("sctl_cuda_direct_set_src_density");
  }

  template <class Vec> void SetTrgCoord(const std::string& name, const Vec& trg_coord) {
    SCTL_CUDA_ASSERT(trg_.find(name) != trg_.end());
    const Long Nt = trg_coord.Dim() / DIM;
    SCTL_CUDA_ASSERT(trg_coord.Dim() == Nt * DIM);
    if (sctl_cuda_direct_set_trg_coord(h_, trg_[name].slot, Nt, Nt ? &trg_coord[0] : nullptr)) detail::die("sctl_cuda_direct_set_trg_coord");
    trg_[name].Nt = Nt;
  }

  // U is resized if needed and OVERWRITTEN (fmm-wrapper.txx:560-563)
  template <class Vec> void EvalDirect(Vec& U, const std::string& trg_name) const {
    auto it = trg_.find(trg_name);
    SCTL_CUDA_ASSERT(it != trg_.end());  // "Target name does not exist." (:555)
    const Long n = it->second.Nt * it->second.dim;
    if (U.Dim() != n) U.ReInit(n);
    for (const S2T& e : s2t_ctx_)  // the reference dereferences ctx_ptr at every Eval: pick up a changed wavenumber
      if (e.trg == it->second.slot) push_s2t(e);
    if (sctl_cuda_direct_eval(h_, it->second.slot, (int)(digits_ >= 16 ? -1 : digits_), n ? &U[0] : nullptr, n)) detail::die("sctl_cuda_direct_eval");
  }
  template <class Vec> void Eval(Vec& U, const std::string& trg_name) const { EvalDirect(U, trg_name); }

  // For the layers built on top (bio_apply.hpp): the C-ABI session and the slot numbers behind the names
  sctl_cuda_direct_t Handle() const { return h_; }
  int SrcSlot(const std::string& name) const {
    auto it = src_.find(name);
    SCTL_CUDA_ASSERT(it != src_.end());
    return it->second;
  }
  int TrgSlot(const std::string& name) const {
    auto it = trg_.find(name);
    SCTL_CUDA_ASSERT(it != trg_.end());
    return it->second.slot;
  }
  Integer Accuracy() const { return digits_; }
  // pushes the kernels' current contexts (wavenumbers) into the session, as EvalDirect does before it evaluates
  void RefreshContexts(const std::string& trg_name) const {
    const int slot = TrgSlot(trg_name);
    for (const S2T& e : s2t_ctx_)
      if (e.trg == slot) push_s2t(e);
  }

 private:
  struct Trg {
    int slot;
    Integer dim;
    Long Nt;
  };
  struct S2T {
    int src, trg, kid;
    const void* ctx;
  };
  void push_s2t(const S2T& e) const {
    if (sctl_cuda_direct_set_kernel_s2t(h_, e.src, e.trg, e.kid, e.ctx, e.ctx ? (int64_t)sizeof(Real) : 0)) detail::die("sctl_cuda_direct_set_kernel_s2t");
  }
  void drop_s2t(int src, int trg) {
    for (size_t i = 0; i < s2t_ctx_.size();)
      if (s2t_ctx_[i].src == src && s2t_ctx_[i].trg == trg) s2t_ctx_.erase(s2t_ctx_.begin() + i);
      else i++;
  }
  static int take_slot(std::vector<int>& free_list, int& next) {
    if (free_list.empty()) return next++;
    const int s = free_list.back();
    free_list.pop_back();
    return s;
  }
  sctl_cuda_direct_t h_;
  Integer digits_;
  int next_src_ = 0, next_trg_ = 0;
  std::vector<int> free_src_, free_trg_;
  std::vector<S2T> s2t_ctx_;  // interactions whose kernel has a context pointer
  std::map<std::string, int> src_;
  std::map<std::string, Trg> trg_;
};

}  // namespace sctl_cuda

#endif  // SCTL_CUDA_PARTICLE_DIRECT_HPP_
