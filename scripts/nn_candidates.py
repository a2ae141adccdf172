"""Diagnostics: screened fusedL2NN on Gaussian / make_blobs-like / far-from-origin data: time and the candidate statistics of
the last chunk (b2d_debug_nn_stats).  argv: gauss|blobs|offset m n k"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from raft_b200.distance import fused_l2_nn
from raft_b200.common import DeviceResources
kind = sys.argv[1]; m = int(sys.argv[2]); n = int(sys.argv[3]); k = int(sys.argv[4])
dev = torch.device("cuda", 0)
h = DeviceResources()
if kind == "blobs":
    c = bench.centers_device(k, torch, dev)
    x = bench.blobs_device(m, k, 1234, c, torch, dev)
    y = bench.blobs_device(n, k, 4321, c, torch, dev)
else:
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(m, k, device="cuda", generator=g) * 3
    y = torch.randn(n, k, device="cuda", generator=g) * 3
    if kind == "offset":     # Gaussian cloud far from the origin
        x += 20.0; y += 20.0
fn = lambda: fused_l2_nn(x, y, sqrt=False, handle=h)
fn(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); fn(); b.record(); torch.cuda.synchronize()
print(f"{kind} {m}x{n}x{k}: {a.elapsed_time(b):.2f} ms")
import ctypes
from raft_b200 import _lib
L = _lib.lib()
ws = h.workspace(L.b2d_fused_l2_nn_workspace_bytes(m, n, k))      # the handle's cached workspace: the one the call used
st = (ctypes.c_uint * 7)()
L.b2d_debug_nn_stats(h.stream_ptr, ws.data_ptr(), m, n, k, st)
print("last chunk: list slots %d, overflow %d, go_screen %d, go_exact %d, candidates after the trial %.2f / row, found by the screen %.2f / row"
      % (st[0], st[1], st[2], st[3], st[5] / m, st[6] / m))
