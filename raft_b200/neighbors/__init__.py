"""Mirror of the (removed) ``pylibraft.neighbors`` entry points that sit directly behind the distance path."""
from . import brute_force  # noqa: F401
