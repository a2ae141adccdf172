import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from raft_b200.distance import pairwise_distance, fused_l2_nn
from raft_b200.common import DeviceResources
what = sys.argv[1]; m = int(sys.argv[2]); n = int(sys.argv[3]); k = int(sys.argv[4]); iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
h = DeviceResources()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(m, k, device="cuda", generator=g) * 3
y = torch.randn(n, k, device="cuda", generator=g) * 3
if what == "nn":
    fn = lambda: fused_l2_nn(x, y, sqrt=False, handle=h)
elif what.startswith("knn"):   # knn<neighbours>, e.g. knn16: x = queries, y = dataset
    from raft_b200.neighbors import brute_force
    kk = int(what[3:] or 10)
    fn = lambda: brute_force.knn(y, x, k=kk, handle=h)
elif what.endswith("_f16"):   # fp16 inputs, fp32 accumulate / output
    out = torch.empty(m, n, device="cuda")
    xh, yh = x.half(), y.half()
    fn = lambda: pairwise_distance(xh, yh, out=out, metric=what[:-4], handle=h)
else:
    out = torch.empty(m, n, device="cuda")
    fn = lambda: pairwise_distance(x, y, out=out, metric=what, handle=h)
for _ in range(2): fn()
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
for a, b in ev:
    a.record(); fn(); b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)[len(ev)//2]
print(f"{what} {m}x{n}x{k} mode={os.environ.get('B2D_STORE_MODE','0')}: {ms:.3f} ms {m*n/ms/1e6:.1f} Gpairs/s out {m*n*4/ms/1e6:.0f} GB/s")
