import sys, torch, os
sys.path.insert(0, ".")
from raft_b200.distance import fused_distance_nn
from raft_b200.common import DeviceResources
h = DeviceResources()
if len(sys.argv) > 1 and sys.argv[1] == "exact":   # the exact kernel only (no screening)
    from raft_b200 import _lib
    _lib.lib().b2d_set_option(b"nn_screen", 0.0)
m, n, k = 1000000, 1000000, 96
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(m, k, device="cuda", generator=g) + 0.5; y = torch.randn(n, k, device="cuda", generator=g) + 0.5
for metric in ("cosine", "correlation"):
    f = lambda: fused_distance_nn(x, y, metric=metric, handle=h)
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); f(); b.record(); torch.cuda.synchronize()
    print(metric, f"{m}x{n}x{k}: {a.elapsed_time(b):.1f} ms")
