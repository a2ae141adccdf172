"""Per-tile clock64 timeline of the coarse fusedL2NN kernel on SM 0 (library built with -DSC_TRACE; RAFT_B200_LIB points at it)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from raft_b200 import _lib
from raft_b200.common import DeviceResources
import bench
dev = torch.device("cuda", 0)
m, n, k = 1_000_000, 1 << 20, 96
c = bench.centers_device(k, torch, dev)
q = bench.blobs_device(m, k, 1234, c, torch, dev); db = bench.blobs_device(n, k, 4321, c, torch, dev)
L = _lib.lib(); h = DeviceResources()
ws = h.workspace(L.b2d_fused_l2_nn_workspace_bytes(m, n, k)); keys = torch.empty(m, dtype=torch.int64, device=dev)
for rep in range(2):
    _lib.check(L.b2d_fused_l2_nn_keys(h.stream_ptr, keys.data_ptr(), q.data_ptr(), k, db.data_ptr(), k, None, None, m, n, k, 0, 1, ws.data_ptr(), ws.numel()))
raw = ctypes.CDLL(os.environ["RAFT_B200_LIB"])
allbuf = np.zeros(8 * 4096 + 3 * 16 * 2048, dtype=np.int64)
raw.b2d_debug_screen_trace(allbuf.ctypes.data_as(ctypes.c_void_p))
buf = allbuf[:8 * 4096].reshape(8, 4096); wt = allbuf[8 * 4096:].reshape(3, 16, 2048)
T = buf - buf[2, 200]
names = ["tempty seen by issuer", "afull seen", "MMAs+commits issued", "tfull seen (w0)", "w0 main pass done", "w0 arrives tempty", "w15 arrives tempty", "tfull seen (w9)"]
print("tile   " + "  ".join(f"{i}" for i in range(8)))
for t in range(200, 232):
    print(t, " ".join(f"{int(T[e, t]):7d}" for e in range(8)))
d = lambda a, b, lo=300, hi=3000: float(np.median(buf[a, lo:hi] - buf[b, lo:hi]))
print("period per tile (issue to issue, same stage /2):", float(np.median(buf[2, 302:3000] - buf[2, 300:2998])) / 2)
print("issue -> tfull seen by w0 (MMA exec + c1):", d(3, 2))
print("tfull seen -> w0 main pass done:", d(4, 3), " -> w0 arrive:", d(5, 3), " w15 arrive:", d(6, 3))
print("last(w0,w15) arrive -> issuer sees tempty of the NEXT tile on this stage:", float(np.median(buf[0, 302:3000] - np.maximum(buf[5, 300:2998], buf[6, 300:2998]))))
print("tempty seen -> afull seen:", d(1, 0), " -> issued:", d(2, 1))

seen, arr, hit = wt[0][:, 300:1800], wt[1][:, 300:1800], wt[2][:, 300:1800]
dur = arr - seen
print("per warp-tile duration (tfull seen -> arrive): no-hit median %.0f p90 %.0f ; hit median %.0f p90 %.0f ; hit share %.3f" % (
    np.median(dur[hit == 0]), np.percentile(dur[hit == 0], 90), np.median(dur[hit == 1]), np.percentile(dur[hit == 1], 90), hit.mean()))
first = seen.min(axis=0); last_arr = arr.max(axis=0)
print("tile: first warp sees tfull -> last warp arrives: median %.0f p90 %.0f" % (np.median(last_arr - first), np.percentile(last_arr - first, 90)))
late = seen - first
print("lateness of a warp in seeing tfull (busy with the previous tile): median %.0f p90 %.0f max %.0f" % (np.median(late), np.percentile(late, 90), late.max()))
prev_hit = wt[2][:, 299:1799]
print("lateness when the warp had a hit in the previous tile: median %.0f ; without: median %.0f" % (np.median(late[prev_hit == 1]), np.median(late[prev_hit == 0])))
for t in range(300, 306):
    print(t, "seen-first", (wt[0][:, t] - wt[0][:, t].min()).tolist(), "dur", (wt[1][:, t] - wt[0][:, t]).tolist(), "hit", wt[2][:, t].tolist())
