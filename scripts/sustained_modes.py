"""Sustained (power-capped) step time of the 100k x 100k x 128 pairwise step: TMA-store epilogue vs the
direct register->global epilogue (forced by an 8-byte-misaligned output base)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from raft_b200.distance import pairwise_distance
from raft_b200.common import DeviceResources
m = n = 100000; k = 128; steps = 60
h = DeviceResources()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(m, k, device="cuda", generator=g) * 3; y = torch.randn(n, k, device="cuda", generator=g) * 3
buf = torch.empty(m * n + 4, device="cuda")
outs = {"tma-store": buf[:m * n].view(m, n), "direct st.v2": buf[2:m * n + 2].view(m, n)}
for rep in range(2):
    for name, out in outs.items():
        for _ in range(3): pairwise_distance(x, y, out=out, metric="sqeuclidean", handle=h)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            a.record(); pairwise_distance(x, y, out=out, metric="sqeuclidean", handle=h); b.record()
        torch.cuda.synchronize()
        per = [a.elapsed_time(b) for a, b in ev]
        print(f"{name:14s} first10 {sum(per[:10])/10:.2f} ms  last30 {sum(per[30:])/30:.2f} ms  min {min(per):.2f}")
        time.sleep(2.0)
