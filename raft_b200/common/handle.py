"""DeviceResources / Handle / Stream (mirrors python/pylibraft/pylibraft/common/handle.pyx:26-116,
201-227).  The C++ raft::resources this stands in for carries a CUDA stream and a workspace
memory resource (cpp/include/raft/core/resource/cuda_stream.hpp:58-64,
cpp/include/raft/core/resource/device_memory_resource.hpp:100-129): here a torch stream and a
grow-only device scratch buffer."""
from __future__ import annotations

import functools

import torch


class Stream:
    """pylibraft.common.Stream: owns a CUDA stream."""

    def __init__(self, device=None):
        self._s = torch.cuda.Stream(device=device)

    def sync(self):
        self._s.synchronize()

    def get_ptr(self) -> int:
        return int(self._s.cuda_stream)

    @property
    def torch_stream(self):
        return self._s


class DeviceResources:
    def __init__(self, stream=None, n_streams: int = 0, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("raft_b200 needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        if stream is None:
            self._stream = torch.cuda.current_stream(self.device)
        elif isinstance(stream, Stream):
            self._stream = stream.torch_stream
        elif isinstance(stream, torch.cuda.Stream):
            self._stream = stream
        else:  # raw cudaStream_t
            self._stream = torch.cuda.ExternalStream(int(stream), device=self.device)
        self._ws = None

    # -- raft::resource::get_cuda_stream
    @property
    def stream_ptr(self) -> int:
        return int(self._stream.cuda_stream)

    @property
    def torch_stream(self):
        return self._stream

    def sync(self):
        """raft::resource::sync_stream (cuda_stream.hpp:83-86)."""
        self._stream.synchronize()

    def getHandle(self):
        return self

    # -- workspace memory resource: grow-only scratch, stream-ordered with this handle's stream
    def workspace(self, nbytes: int) -> torch.Tensor:
        nbytes = max(int(nbytes), 256)
        if self._ws is None or self._ws.numel() < nbytes:
            with torch.cuda.stream(self._stream):
                self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws


Handle = DeviceResources  # legacy alias (handle.pyx:121-198)


def auto_sync_handle(f):
    """handle=None -> create one and sync on exit (handle.pyx:201-227)."""

    @functools.wraps(f)
    def wrapper(*args, handle=None, **kwargs):
        sync = handle is None
        handle = handle if handle is not None else DeviceResources()
        ret = f(*args, handle=handle, **kwargs)
        if sync:
            handle.sync()
        return ret

    return wrapper
