"""__cuda_array_interface__ view of an input (mirrors
python/pylibraft/pylibraft/common/cai_wrapper.py:10-36 and ai_wrapper.py:23-83)."""
from __future__ import annotations

import numpy as np


class cai_wrapper:
    def __init__(self, obj):
        if not hasattr(obj, "__cuda_array_interface__"):
            raise TypeError("expected an object exposing __cuda_array_interface__")
        self.obj = obj  # keep the owner alive
        self.cai = obj.__cuda_array_interface__
        self.dtype = np.dtype(self.cai["typestr"])
        self.shape = tuple(int(s) for s in self.cai["shape"])
        self.data = int(self.cai["data"][0] or 0)
        strides = self.cai.get("strides")
        if strides is None:
            self.c_contiguous, self.f_contiguous = True, len(self.shape) <= 1
            st, acc = [], self.dtype.itemsize
            for s in reversed(self.shape):
                st.append(acc)
                acc *= s
            self.strides = tuple(reversed(st))
        else:
            self.strides = tuple(int(s) for s in strides)
            self.c_contiguous = self._is_contig(self.shape, self.strides, self.dtype.itemsize, "C")
            self.f_contiguous = self._is_contig(self.shape, self.strides, self.dtype.itemsize, "F")

    @staticmethod
    def _is_contig(shape, strides, itemsize, order):
        dims = range(len(shape) - 1, -1, -1) if order == "C" else range(len(shape))
        acc = itemsize
        for d in dims:
            if shape[d] != 1 and strides[d] != acc:
                return False
            acc *= shape[d]
        return True

    def validate_shape_dtype(self, expected_dims=None, expected_dtype=None):
        """common/ai_wrapper.py:65-79."""
        if expected_dims is not None and len(self.shape) != expected_dims:
            raise ValueError(f"unexpected shape {self.shape} - expected {expected_dims} dimensions")
        if expected_dtype is not None and self.dtype != np.dtype(expected_dtype):
            raise TypeError(f"invalid dtype {self.dtype} - expected {expected_dtype}")
