"""Device array container (mirrors python/pylibraft/pylibraft/common/device_ndarray.py:10-160).

The reference backs it with rmm.DeviceBuffer; RMM is not installable offline, so the storage
here is a torch CUDA tensor.  The surface is the same: shape / dtype / strides /
__cuda_array_interface__ / copy_to_host() / device_ndarray.empty()."""
from __future__ import annotations

import numpy as np
import torch

_NP2TORCH = {np.dtype(np.float32): torch.float32, np.dtype(np.float16): torch.float16,
             np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32,
             np.dtype(np.int64): torch.int64, np.dtype(np.uint8): torch.uint8}


class device_ndarray:
    def __init__(self, np_ndarray_or_tensor):
        if isinstance(np_ndarray_or_tensor, torch.Tensor):
            t = np_ndarray_or_tensor
            if not t.is_cuda:
                t = t.cuda()
        else:
            a = np.ascontiguousarray(np_ndarray_or_tensor)
            t = torch.from_numpy(a).cuda()
        self._t = t

    @classmethod
    def empty(cls, shape, dtype=np.float32, order="C"):
        if order != "C":
            raise ValueError("only C order outputs are allocated by this container")
        t = torch.empty(tuple(shape) if not isinstance(shape, int) else (shape,),
                        dtype=_NP2TORCH[np.dtype(dtype)], device="cuda")
        return cls(t)

    @property
    def tensor(self) -> torch.Tensor:
        return self._t

    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def dtype(self):
        return np.dtype(str(self._t.dtype).replace("torch.", ""))

    @property
    def strides(self):
        return tuple(s * self._t.element_size() for s in self._t.stride())

    @property
    def c_contiguous(self):
        return self._t.is_contiguous()

    @property
    def __cuda_array_interface__(self):
        return self._t.__cuda_array_interface__

    def copy_to_host(self) -> np.ndarray:
        return self._t.cpu().numpy()

    def __repr__(self):
        return f"device_ndarray(shape={self.shape}, dtype={self.dtype})"
