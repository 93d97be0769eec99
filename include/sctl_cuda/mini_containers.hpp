// sctl_cuda/mini_containers.hpp -- the few members of sctl::Vector / sctl::Matrix the kernel layer touches
// (Dim, begin, operator[], ReInit, SetZero; reference: include/sctl/vector.hpp, matrix.hpp), for building and
// testing the host layer where SCTL itself is not installed.  Callers that have SCTL pass sctl::Vector directly.
#ifndef SCTL_CUDA_MINI_CONTAINERS_HPP_
#define SCTL_CUDA_MINI_CONTAINERS_HPP_

#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <vector>

namespace sctl_cuda {

template <class T> class Vector {
 public:
  Vector() : ptr_(nullptr), dim_(0), own_(true) {}
  explicit Vector(int64_t dim, T* data = nullptr, bool own_data = true) : ptr_(nullptr), dim_(0), own_(true) { Init(dim, data, own_data); }
  Vector(std::initializer_list<T> v) : ptr_(nullptr), dim_(0), own_(true) {
    Init((int64_t)v.size(), nullptr, true);
    int64_t i = 0;
    for (const T& x : v) ptr_[i++] = x;
  }
  Vector(const Vector& o) : ptr_(nullptr), dim_(0), own_(true) {
    Init(o.dim_, nullptr, true);
    if (dim_) std::memcpy(ptr_, o.ptr_, sizeof(T) * dim_);
  }
  Vector& operator=(const Vector& o) {
    if (this != &o) {
      ReInit(o.dim_);
      if (dim_) std::memcpy(ptr_, o.ptr_, sizeof(T) * dim_);
    }
    return *this;
  }
  ~Vector() { Free(); }

  int64_t Dim() const { return dim_; }
  T* begin() { return ptr_; }
  const T* begin() const { return ptr_; }
  T& operator[](int64_t i) { return ptr_[i]; }
  const T& operator[](int64_t i) const { return ptr_[i]; }
  void SetZero() {
    if (dim_) std::memset(ptr_, 0, sizeof(T) * dim_);
  }
  // Like sctl::Vector::ReInit: a non-owning view is dropped, owned storage is reallocated (contents undefined).
  void ReInit(int64_t dim, T* data = nullptr, bool own_data = true) {
    Free();
    Init(dim, data, own_data);
  }

 private:
  void Init(int64_t dim, T* data, bool own_data) {
    dim_ = dim;
    own_ = own_data;
    if (own_) {
      ptr_ = dim ? new T[dim] : nullptr;
      if (data && dim) std::memcpy(ptr_, data, sizeof(T) * dim);
    } else {
      ptr_ = data;
    }
  }
  void Free() {
    if (own_ && ptr_) delete[] ptr_;
    ptr_ = nullptr;
    dim_ = 0;
    own_ = true;
  }
  T* ptr_;
  int64_t dim_;
  bool own_;
};

template <class T> class Matrix {
 public:
  Matrix() : r_(0), c_(0) {}
  Matrix(int64_t r, int64_t c) : r_(r), c_(c), d_((size_t)(r * c)) {}
  int64_t Dim(int i) const { return i == 0 ? r_ : c_; }
  void ReInit(int64_t r, int64_t c) {
    r_ = r;
    c_ = c;
    d_.assign((size_t)(r * c), T());
  }
  void SetZero() { std::fill(d_.begin(), d_.end(), T()); }
  T* begin() { return d_.data(); }
  const T* begin() const { return d_.data(); }
  T* operator[](int64_t i) { return d_.data() + i * c_; }           // row pointer, like sctl::Matrix::operator[]
  const T* operator[](int64_t i) const { return d_.data() + i * c_; }
  T& operator()(int64_t i, int64_t j) { return d_[(size_t)(i * c_ + j)]; }
  const T& operator()(int64_t i, int64_t j) const { return d_[(size_t)(i * c_ + j)]; }

 private:
  int64_t r_, c_;
  std::vector<T> d_;
};

}  // namespace sctl_cuda

#endif  // SCTL_CUDA_MINI_CONTAINERS_HPP_
