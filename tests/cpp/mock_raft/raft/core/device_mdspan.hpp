// MOCK of cpp/include/raft/core/mdspan_types.hpp:19-32 + device_mdspan.hpp:94-97,169-199 (without the
// host_device_accessor: the shim only touches data_handle / extent / stride)
#pragma once
#include <cuda/std/mdspan>
#include <cstdint>
namespace raft {
using cuda::std::dynamic_extent;
using cuda::std::extents;
using cuda::std::layout_left;
using cuda::std::layout_right;
using cuda::std::layout_stride;
using layout_c_contiguous = layout_right;
using layout_f_contiguous = layout_left;
using row_major           = layout_right;
using col_major           = layout_left;
template <typename IndexType>
using matrix_extent = cuda::std::extents<IndexType, dynamic_extent, dynamic_extent>;
template <typename ElementType, typename IndexType = std::uint32_t, typename LayoutPolicy = layout_c_contiguous>
using device_matrix_view = cuda::std::mdspan<ElementType, matrix_extent<IndexType>, LayoutPolicy>;
template <typename ElementType, typename IndexType = std::uint32_t, typename LayoutPolicy = layout_c_contiguous>
auto constexpr make_device_matrix_view(ElementType* ptr, IndexType n_rows, IndexType n_cols)
{
  return device_matrix_view<ElementType, IndexType, LayoutPolicy>{ptr, matrix_extent<IndexType>{n_rows, n_cols}};
}
template <typename ElementType, typename IndexType = std::uint32_t>
auto make_device_strided_matrix_view(ElementType* ptr, IndexType n_rows, IndexType n_cols, IndexType stride)
{
  cuda::std::array<IndexType, 2> strides{stride, IndexType(1)};
  return device_matrix_view<ElementType, IndexType, layout_stride>{
    ptr, typename layout_stride::template mapping<matrix_extent<IndexType>>{matrix_extent<IndexType>{n_rows, n_cols}, strides}};
}
}  // namespace raft
