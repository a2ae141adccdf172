"""pylibraft.config.set_output_as (python/pylibraft/pylibraft/config.py:9-35)."""
from raft_b200.common.outputs import set_output_as  # noqa: F401
