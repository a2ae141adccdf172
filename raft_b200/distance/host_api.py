"""Host-buffer entry point: pairwise distances for inputs that live in HOST memory.

This is the end-to-end path a caller without device arrays takes (and what bench.py's `e2e`
measures): X and Y are copied host->device, the [m,n] result is produced in row slabs on the device
and streamed back device->host through a ring of pinned slab buffers, the copy of slab s
overlapping the kernel of slab s+1 (two CUDA streams).  The result is too large to keep on the
device next to anything else at the BASELINE shapes (40 GB at 100k x 100k, 160 GB at 200k x 200k;
SURVEY.md hard part B), so slab streaming is also how those shapes are meant to be consumed."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from ..common import DeviceResources
from .distance_type import resolve_metric


class HostPairwise:
    """Reusable state (pinned staging, device slabs, workspace) for repeated host-buffer calls."""

    def __init__(self, m, n, k, slab_rows=None, device=None):
        self.m, self.n, self.k = int(m), int(n), int(k)
        if slab_rows is None:
            slab_rows = max(128, min(self.m, (1 << 30) // max(4 * self.n, 1) // 128 * 128))  # ~1 GiB slabs
        self.slab_rows = int(min(slab_rows, self.m))
        self.h = DeviceResources(device=device)
        dev = self.h.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.x_dev = torch.empty((self.m, self.k), dtype=torch.float32, device=dev)
        self.y_dev = torch.empty((self.n, self.k), dtype=torch.float32, device=dev)
        self.x_pin = torch.empty((self.m, self.k), dtype=torch.float32).pin_memory()
        self.y_pin = torch.empty((self.n, self.k), dtype=torch.float32).pin_memory()
        self.d_slab = [torch.empty((self.slab_rows, self.n), dtype=torch.float32, device=dev) for _ in range(2)]
        self.h_slab = [torch.empty((self.slab_rows, self.n), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.slab_done = [torch.cuda.Event() for _ in range(2)]
        self.copy_done = [torch.cuda.Event() for _ in range(2)]
        self.h2d_bytes = 4 * self.k * (self.m + self.n)
        self.d2h_bytes = 4 * self.m * self.n

    def run(self, X: np.ndarray, Y: np.ndarray, metric="sqeuclidean", p=2.0, out: np.ndarray | None = None,
            consume=None):
        """Computes all slabs.  out: optional host [m,n] float32 array to fill; consume: optional
        callback(row0, pinned_slab_view) invoked when a slab has landed in pinned host memory."""
        L = _lib.lib()
        mt = int(resolve_metric(metric))
        h, st = self.h, self.h.torch_stream
        self.x_pin.copy_(torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)))
        self.y_pin.copy_(torch.from_numpy(np.ascontiguousarray(Y, dtype=np.float32)))
        with torch.cuda.stream(st):
            self.x_dev.copy_(self.x_pin, non_blocking=True)
            self.y_dev.copy_(self.y_pin, non_blocking=True)
        pending = []
        for s, r0 in enumerate(range(0, self.m, self.slab_rows)):
            b = s & 1
            rows = min(self.slab_rows, self.m - r0)
            if s >= 2:   # the device slab is free once its previous D2H copy has finished
                st.wait_event(self.copy_done[b])
                self._deliver(pending.pop(0), out, consume)
            need = L.b2d_pairwise_workspace_bytes(mt, _lib.B2D_F32, rows, self.n, self.k)
            ws = h.workspace(need)
            with torch.cuda.stream(st):
                _lib.check(L.b2d_pairwise_distance(h.stream_ptr, mt, _lib.B2D_F32, self.x_dev[r0:].data_ptr(), self.k,
                                                   self.y_dev.data_ptr(), self.k, self.d_slab[b].data_ptr(), self.n,
                                                   rows, self.n, self.k, 1, float(p), ws.data_ptr(), ws.numel()))
                self.slab_done[b].record(st)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.slab_done[b])
                self.h_slab[b][:rows].copy_(self.d_slab[b][:rows], non_blocking=True)
                self.copy_done[b].record(self.copy_stream)
            pending.append((b, r0, rows))
        for item in pending:
            self._deliver(item, out, consume)
        return out

    def _deliver(self, item, out, consume):
        b, r0, rows = item
        self.copy_done[b].synchronize()
        view = self.h_slab[b][:rows]
        if out is not None:
            out[r0:r0 + rows] = view.numpy()
        if consume is not None:
            consume(r0, view)


def pairwise_distance_host(X: np.ndarray, Y: np.ndarray, metric="euclidean", p=2.0, out=None, slab_rows=None):
    """numpy in, numpy out (mirrors pylibraft.distance.pairwise_distance for host arrays)."""
    m, k = X.shape
    n = Y.shape[0]
    if out is None:
        out = np.empty((m, n), dtype=np.float32)
    HostPairwise(m, n, k, slab_rows=slab_rows).run(X, Y, metric=metric, p=p, out=out)
    return out
