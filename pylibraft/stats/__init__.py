from raft_b200.stats import silhouette_score, trustworthiness_score  # noqa: F401
