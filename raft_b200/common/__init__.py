"""Mirror of pylibraft.common (python/pylibraft/pylibraft/common/__init__.py)."""
from .cai_wrapper import cai_wrapper  # noqa: F401
from .device_ndarray import device_ndarray  # noqa: F401
from .handle import DeviceResources, Handle, Stream, auto_sync_handle  # noqa: F401
from .outputs import auto_convert_output, set_output_as  # noqa: F401
