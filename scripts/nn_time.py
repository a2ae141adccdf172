"""fusedL2NN timing on the benchmark's data: 1M queries x N db rows x 96 (CUDA events, 3 repetitions).
argv: [db rows = 1048576] [sort = 0|1: db rows pre-sorted by squared norm] ; RAFT_B200_LIB selects the library build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from raft_b200 import _lib
from raft_b200.common import DeviceResources
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
srt = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda", 0)
m, k = 1_000_000, 96
c = bench.centers_device(k, torch, dev)
q = bench.blobs_device(m, k, 1234, c, torch, dev)
db = bench.blobs_device(n, k, 4321, c, torch, dev)
if srt:
    db = db[torch.argsort((db * db).sum(1))].contiguous()
L = _lib.lib(); h = DeviceResources()
ws = h.workspace(L.b2d_fused_l2_nn_workspace_bytes(m, n, k))
keys = torch.empty(m, dtype=torch.int64, device=dev)
import threading, pynvml
pynvml.nvmlInit(); _h = pynvml.nvmlDeviceGetHandleByIndex(0); _clk = []; _stop = False
def _sample():
    import time
    while not _stop:
        _clk.append(pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)); time.sleep(0.01)
_t = threading.Thread(target=_sample); _t.start()
ms = []
for rep in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(h.torch_stream):
        a.record(h.torch_stream)
        _lib.check(L.b2d_fused_l2_nn_keys(h.stream_ptr, keys.data_ptr(), q.data_ptr(), k, db.data_ptr(), k, None, None, m, n, k, 0, 1, ws.data_ptr(), ws.numel()))
        b.record(h.torch_stream)
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
_stop = True; _t.join()
print("sm clock MHz during the calls: median", sorted(_clk)[len(_clk) // 2], "min", min(_clk), "max", max(_clk))
print(f"fusedL2NN 1M x {n} x 96 sorted={srt} lib={os.environ.get('RAFT_B200_LIB', 'default')}: ms {[round(v, 1) for v in ms]}")
# exactness spot check: 2048 sampled queries against the fp64 arg-min
from oracle import device_check as dc
rows = torch.randperm(m, device=dev)[:2048]
rv, ri = dc.nn_exact_fp64(q[rows], db)
idx = (keys[rows] & 0xffffffff).to(torch.int64)
print("strict idx match", float((idx == ri).float().mean()))
