// Minimal stand-in for the parts of raft/core that the distance shim needs, used ONLY when the real
// RAFT headers are not on the include path (RMM / CCCL 3 are not installable offline; SURVEY.md
// hard part F).  With real RAFT present, define RAFT_B200_USE_REAL_RAFT and include
// <raft/core/resources.hpp>, <raft/core/device_mdspan.hpp>, <raft/core/kvp.hpp>, <raft/core/error.hpp>
// before the distance headers: the shim only uses the names below, with the reference's semantics.
//
//   raft::resources + raft::resource::get_cuda_stream   cpp/include/raft/core/resources.hpp:38-128,
//                                                       cpp/include/raft/core/resource/cuda_stream.hpp:58-64
//   workspace memory resource                           cpp/include/raft/core/resource/device_memory_resource.hpp:100-129
//   raft::device_matrix_view / make_device_matrix_view   cpp/include/raft/core/device_mdspan.hpp:94-97,169-176
//   raft::KeyValuePair                                  cpp/include/raft/core/kvp.hpp:20-62
//   raft::logic_error / raft::cuda_error / RAFT_EXPECTS cpp/include/raft/core/error.hpp:218-239
#pragma once
#ifndef RAFT_B200_USE_REAL_RAFT
#include <cuda_runtime_api.h>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace raft {

struct exception : std::runtime_error { using std::runtime_error::runtime_error; };
struct logic_error : exception { using exception::exception; };
struct cuda_error : exception { using exception::exception; };

#define RAFT_EXPECTS(cond, msg)                                   \
  do {                                                            \
    if (!(cond)) throw ::raft::logic_error(std::string("RAFT failure: ") + (msg)); \
  } while (0)

struct row_major {};
struct col_major {};

// device_matrix_view<T, IdxT, Layout>: data_handle() + extent(i), like the mdspan alias it replaces
template <typename T, typename IdxT = int, typename Layout = row_major>
class device_matrix_view {
 public:
  using element_type = T;
  using index_type   = IdxT;
  using layout_type  = Layout;
  device_matrix_view(T* p, IdxT rows, IdxT cols) : p_(p), r_(rows), c_(cols) {}
  T* data_handle() const { return p_; }
  IdxT extent(int i) const { return i == 0 ? r_ : c_; }
 private:
  T* p_;
  IdxT r_, c_;
};
template <typename T, typename IdxT = int, typename Layout = row_major>
device_matrix_view<T, IdxT, Layout> make_device_matrix_view(T* p, IdxT rows, IdxT cols)
{
  return device_matrix_view<T, IdxT, Layout>(p, rows, cols);
}

template <typename K, typename V>
struct KeyValuePair {
  K key;
  V value;
};

// raft::resources: a stream plus a grow-only workspace (the default workspace resource of the
// reference is a pool limited to 1/4 of device memory; here: cudaMallocAsync on the handle's stream)
class resources {
 public:
  explicit resources(cudaStream_t s = nullptr) : stream_(s) {}
  resources(const resources&) = delete;
  ~resources() { if (ws_) cudaFreeAsync(ws_, stream_); }
  cudaStream_t stream() const { return stream_; }
  void* workspace(std::size_t bytes) const
  {
    if (bytes > ws_bytes_) {
      if (ws_) cudaFreeAsync(ws_, stream_);
      if (cudaMallocAsync(&ws_, bytes, stream_) != cudaSuccess) throw cuda_error("workspace allocation failed");
      ws_bytes_ = bytes;
    }
    return ws_;
  }
  std::size_t workspace_bytes() const { return ws_bytes_; }
 private:
  cudaStream_t stream_;
  mutable void* ws_              = nullptr;
  mutable std::size_t ws_bytes_ = 0;
};
using device_resources = resources;
using handle_t         = resources;

namespace resource {
inline cudaStream_t get_cuda_stream(resources const& h) { return h.stream(); }
inline void sync_stream(resources const& h)
{
  if (cudaStreamSynchronize(h.stream()) != cudaSuccess) throw cuda_error("cudaStreamSynchronize failed");
}
}  // namespace resource
}  // namespace raft
#endif
