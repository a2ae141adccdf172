from raft_b200.neighbors import brute_force  # noqa: F401
