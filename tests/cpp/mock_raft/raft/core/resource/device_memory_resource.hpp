// MOCK of cpp/include/raft/core/resource/device_memory_resource.hpp:187-203: the workspace resource of a handle
#pragma once
#include <raft/core/resources.hpp>
#include <rmm/device_uvector.hpp>
namespace raft::resource {
inline rmm::device_async_resource_ref get_workspace_resource_ref(resources const&) { return rmm::device_async_resource_ref{}; }
}  // namespace raft::resource
