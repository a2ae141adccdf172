from raft_b200.common import (DeviceResources, Handle, Stream, auto_convert_output, auto_sync_handle,  # noqa: F401
                              cai_wrapper, device_ndarray)
