// MOCK of rmm::device_uvector / rmm::device_async_resource_ref: stream-ordered allocation, released in the destructor
#pragma once
#include <cuda_runtime_api.h>
#include <cstddef>
#include <stdexcept>
namespace rmm {
struct device_async_resource_ref {
  static int& live() { static int n = 0; return n; }   // the test checks that every scratch buffer came from here
};
template <typename T>
class device_uvector {
 public:
  device_uvector(std::size_t n, cudaStream_t s, device_async_resource_ref) : n_(n), s_(s)
  {
    if (n && cudaMallocAsync(reinterpret_cast<void**>(&p_), n * sizeof(T), s) != cudaSuccess) throw std::runtime_error("mock rmm: bad_alloc");
    ++device_async_resource_ref::live();
  }
  ~device_uvector() { if (p_) cudaFreeAsync(p_, s_); }
  device_uvector(const device_uvector&) = delete;
  T* data() { return p_; }
  std::size_t size() const { return n_; }
 private:
  T* p_ = nullptr;
  std::size_t n_;
  cudaStream_t s_;
};
}  // namespace rmm
