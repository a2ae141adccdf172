// MOCK of cpp/include/raft/core/error.hpp:37-58,218-239 and util/cuda_rt_essentials.hpp:23-52
#pragma once
#include <stdexcept>
#include <string>
namespace raft {
struct exception : std::runtime_error { using std::runtime_error::runtime_error; };
struct logic_error : exception { using exception::exception; };
struct cuda_error : exception { using exception::exception; };
}  // namespace raft
#define RAFT_EXPECTS(cond, msg) do { if (!(cond)) throw ::raft::logic_error(std::string("RAFT failure: ") + (msg)); } while (0)
