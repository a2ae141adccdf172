from raft_b200.stats import silhouette_score  # noqa: F401
