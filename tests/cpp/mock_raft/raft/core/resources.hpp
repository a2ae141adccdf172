// MOCK of cpp/include/raft/core/resources.hpp:38-128 for the boundary compile test (tests/test_cpp_shim.py):
// a type-erased resource container with NO convenience members -- in particular no workspace(), no stream();
// everything is reached through raft::resource::get_*() free functions, as in the reference.
#pragma once
#include <cuda_runtime_api.h>
namespace raft {
class resources {
 public:
  resources() = default;
  resources(const resources&) = delete;
  // test-only hook of the mock (the real class stores resource factories); not used by the shim
  void mock_set_stream(cudaStream_t s) { stream_ = s; }
  cudaStream_t mock_stream() const { return stream_; }
 private:
  cudaStream_t stream_ = nullptr;
};
using device_resources = resources;
}  // namespace raft
