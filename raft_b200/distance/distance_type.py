"""raft::distance::DistanceType (enum values SURVEY.md 8(a1)) and the metric-string table of
pylibraft.distance.pairwise_distance (Python bullet)."""
import enum


class DistanceType(enum.IntEnum):
    L2Expanded = 0
    L2SqrtExpanded = 1
    CosineExpanded = 2
    L1 = 3
    L2Unexpanded = 4
    L2SqrtUnexpanded = 5
    InnerProduct = 6
    Linf = 7
    Canberra = 8
    LpUnexpanded = 9
    CorrelationExpanded = 10
    JaccardExpanded = 11
    HellingerExpanded = 12
    Haversine = 13
    BrayCurtis = 14
    JensenShannon = 15
    HammingUnexpanded = 16
    KLDivergence = 17
    RusselRaoExpanded = 18
    DiceExpanded = 19
    Precomputed = 100


DISTANCE_TYPES = {
    "l2": DistanceType.L2SqrtExpanded,
    "euclidean": DistanceType.L2SqrtExpanded,
    "sqeuclidean": DistanceType.L2Expanded,
    "l2_unexpanded": DistanceType.L2SqrtUnexpanded,
    "sqeuclidean_unexpanded": DistanceType.L2Unexpanded,
    "l1": DistanceType.L1,
    "cityblock": DistanceType.L1,
    "manhattan": DistanceType.L1,
    "inner_product": DistanceType.InnerProduct,
    "chebyshev": DistanceType.Linf,
    "linf": DistanceType.Linf,
    "canberra": DistanceType.Canberra,
    "cosine": DistanceType.CosineExpanded,
    "lp": DistanceType.LpUnexpanded,
    "minkowski": DistanceType.LpUnexpanded,
    "correlation": DistanceType.CorrelationExpanded,
    "hellinger": DistanceType.HellingerExpanded,
    "jensenshannon": DistanceType.JensenShannon,
    "hamming": DistanceType.HammingUnexpanded,
    "kl_divergence": DistanceType.KLDivergence,
    "russellrao": DistanceType.RusselRaoExpanded,
    "jaccard": DistanceType.JaccardExpanded,
    "dice": DistanceType.DiceExpanded,
    "braycurtis": DistanceType.BrayCurtis,
}

SUPPORTED_DISTANCES = sorted(DISTANCE_TYPES)

# metric of the reference enum that is outside this engine's scope (2-d lat/lon only in the reference)
UNSUPPORTED = {"haversine"}


def resolve_metric(metric) -> DistanceType:
    if isinstance(metric, DistanceType):
        return metric
    if isinstance(metric, int):
        return DistanceType(metric)
    if isinstance(metric, str):
        if metric in DISTANCE_TYPES:
            return DISTANCE_TYPES[metric]
        if metric in DistanceType.__members__:
            return DistanceType[metric]
    raise ValueError("metric %s is not supported" % (metric,))
