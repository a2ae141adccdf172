"""Where does a rank's time go in the 8-GPU fusedL2NN step?  Emulates one rank on one GPU: the bounds every rank gets
from the first exchange come from world x head rows, so a (world x head)-row call stands in for it (not timed); then the
rank's own sub-chunks are timed one by one (CUDA events), with and without the exact sample pass."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from raft_b200 import _lib
from raft_b200.common import DeviceResources
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0)
m, k = 1_000_000, 96
c = bench.centers_device(k, torch, dev)
q = bench.blobs_device(m, k, 1234, c, torch, dev)
shard = bench.blobs_device(1_000_000, k, 4321, c, torch, dev)
other = bench.blobs_device(7 * 32768, k, 999, c, torch, dev)     # the other ranks' heads
L = _lib.lib(); h = DeviceResources()
ws = h.workspace(L.b2d_fused_l2_nn_workspace_bytes(m, 1_000_000, k))
keys = torch.empty(m, dtype=torch.int64, device=dev)
def call(y, off, init):
    _lib.check(L.b2d_fused_l2_nn_keys(h.stream_ptr, keys.data_ptr(), q.data_ptr(), k, y.data_ptr(), k, None, None, m, y.shape[0], k, off, init, ws.data_ptr(), ws.numel()))
plans = {"[8192,rest]": [8192], "[8192,65536,rest]": [8192, 65536], "[16384,rest]": [16384], "[16384,131072,rest]": [16384, 131072], "[4096,32768,rest]": [4096, 32768]}
for name, plan in plans.items():
    for rep in range(2):
        ev = []
        done = 0
        with torch.cuda.stream(h.torch_stream):
            sizes = plan + [1_000_000 - sum(plan)]
            for ci, rows in enumerate(sizes):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(h.torch_stream)
                call(shard[done:done + rows], done, 1 if ci == 0 else 0)
                b.record(h.torch_stream)
                ev.append((a, b))
                done += rows
                if ci == 0 and len(sizes) > 1:      # "exchange": the other 7 ranks' heads lower the bounds (scaled with the head size)
                    call(other[: 7 * rows], 10_000_000, 0)
        torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    print(f"{name:28s} per call ms {[round(v, 1) for v in ms]}  sum {sum(ms):.1f}")
