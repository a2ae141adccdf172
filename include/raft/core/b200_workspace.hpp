// Customization point for the temporaries of one shim call.  The C ABI never allocates; the reference's rule
// is that a primitive takes its scratch from the handle's workspace memory resource
// (cpp/include/raft/core/resource/device_memory_resource.hpp:100-129,187-203: a limiting adaptor over the
// current device resource, 1/4 of device memory by default).  So:
//   with real RAFT (RAFT_B200_USE_REAL_RAFT): an rmm::device_uvector<char> on
//       raft::resource::get_workspace_resource_ref(handle), stream-ordered on the handle's stream -- it is released
//       behind the kernels the call enqueued, exactly like the reference's own temporaries;
//   stand-alone: cudaMallocAsync / cudaFreeAsync on the same stream (the default pool keeps the memory, so the
//       steady state costs no driver call).
// Either way nothing synchronises and the scratch lives until the work enqueued so far has run.
#pragma once
#include "b200_compat.hpp"
#ifdef RAFT_B200_USE_REAL_RAFT
#include <raft/core/resource/device_memory_resource.hpp>
#include <rmm/device_uvector.hpp>
#endif

namespace raft {
namespace b200 {

class scoped_workspace {
 public:
#ifdef RAFT_B200_USE_REAL_RAFT
  scoped_workspace(raft::resources const& handle, std::size_t bytes)
    : buf_(bytes + 256, raft::resource::get_cuda_stream(handle), raft::resource::get_workspace_resource_ref(handle))
  {
  }
  void* data() { return align(buf_.data()); }
#else
  scoped_workspace(raft::resources const& handle, std::size_t bytes) : stream_(raft::resource::get_cuda_stream(handle))
  {
    if (bytes && cudaMallocAsync(&p_, bytes + 256, stream_) != cudaSuccess)
      throw raft::cuda_error("raft_b200: workspace allocation failed");
  }
  ~scoped_workspace()
  {
    if (p_) cudaFreeAsync(p_, stream_);
  }
  void* data() { return align(p_); }
#endif
  scoped_workspace(const scoped_workspace&)            = delete;
  scoped_workspace& operator=(const scoped_workspace&) = delete;

 private:
  static void* align(void* p)  // the C ABI wants 256-byte aligned scratch
  {
    return reinterpret_cast<void*>((reinterpret_cast<std::uintptr_t>(p) + 255) & ~static_cast<std::uintptr_t>(255));
  }
#ifdef RAFT_B200_USE_REAL_RAFT
  rmm::device_uvector<char> buf_;
#else
  cudaStream_t stream_;
  void* p_ = nullptr;
#endif
};

}  // namespace b200
}  // namespace raft
