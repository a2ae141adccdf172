// MOCK of cpp/include/raft/core/kvp.hpp:20-62
#pragma once
namespace raft {
template <typename _Key, typename _Value>
struct KeyValuePair {
  typedef _Key Key;
  typedef _Value Value;
  Key key;
  Value value;
};
}  // namespace raft
