"""Output conversion (mirrors python/pylibraft/pylibraft/common/outputs.py:42-84 and
pylibraft.config.set_output_as, python/pylibraft/pylibraft/config.py:9-35)."""
from __future__ import annotations

import functools

from .device_ndarray import device_ndarray

_output_as = "raft"


def set_output_as(output):
    """'raft' (device_ndarray), 'torch', 'cupy' or a callable taking a device_ndarray."""
    global _output_as
    if not callable(output) and output not in ("raft", "torch", "cupy"):
        raise ValueError("Unsupported output option %s" % (output,))
    _output_as = output


def _convert(v):
    if not isinstance(v, device_ndarray):
        return v
    if callable(_output_as):
        return _output_as(v)
    if _output_as == "torch":
        return v.tensor
    if _output_as == "cupy":
        import cupy  # noqa: F401  (optional)
        return cupy.asarray(v)
    return v


def auto_convert_output(f):
    @functools.wraps(f)
    def wrapper(*args, **kwargs):
        ret = f(*args, **kwargs)
        if isinstance(ret, tuple):
            return tuple(_convert(r) for r in ret)
        return _convert(ret)

    return wrapper
