"""Mirror of the removed ``pylibraft.distance`` package."""
from .distance_type import DISTANCE_TYPES, SUPPORTED_DISTANCES, DistanceType  # noqa: F401
from .fused_l2_nn import (fused_distance_nn, fused_l2_nn, fused_l2_nn_argmin, fused_l2_nn_sharded,  # noqa: F401
                          shard_bounds)
from .pairwise_distance import distance, pairwise_distance, pairwise_distance_raw  # noqa: F401
from .host_api import HostPairwise, pairwise_distance_host  # noqa: F401
