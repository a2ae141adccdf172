// raft::matrix::argmin -- shim over b2d_row_argmin (include/raft_b200.h); signature of
// cpp/include/raft/matrix/argmin.cuh:25-37 (row-major matrix view in, vector of column indices out).
#pragma once
#include "../distance/distance.cuh"

namespace raft {
namespace matrix {

template <typename math_t, typename idx_t, typename matrix_idx_t>
void argmin(raft::resources const& handle, raft::device_matrix_view<const math_t, matrix_idx_t, raft::row_major> in,
            idx_t* out, matrix_idx_t out_extent)
{
  static_assert(std::is_same<math_t, float>::value && sizeof(idx_t) == 4, "raft_b200: argmin is provided for <float, int>");
  RAFT_EXPECTS(out_extent == in.extent(0), "Size of output vector must equal number of rows in input matrix.");
  raft::distance::detail::b2d_check(b2d_row_argmin(raft::resource::get_cuda_stream(handle), reinterpret_cast<int32_t*>(out),
                                                   in.data_handle(), in.extent(1), in.extent(0), in.extent(1)));
}

}  // namespace matrix
}  // namespace raft
