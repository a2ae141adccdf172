// MOCK of cpp/include/raft/core/resource/cuda_stream.hpp:58-64,97-100
#pragma once
#include <raft/core/resources.hpp>
#include <raft/core/error.hpp>
namespace raft::resource {
inline cudaStream_t get_cuda_stream(resources const& res) { return res.mock_stream(); }
inline void sync_stream(resources const& res)
{
  if (cudaStreamSynchronize(get_cuda_stream(res)) != cudaSuccess) throw raft::cuda_error("sync failed");
}
}  // namespace raft::resource
