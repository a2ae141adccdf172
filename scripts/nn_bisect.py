import sys, torch
sys.path.insert(0, ".")
from raft_b200.distance import fused_l2_nn
m, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = torch.randn(m, k, device="cuda") * 2; y = torch.randn(n, k, device="cuda") * 2
i, v = fused_l2_nn(x, y, sqrt=False); torch.cuda.synchronize()
d = torch.cdist(x[:8], y).pow(2).min(dim=1)
print("ok", m, n, k, (i[:8].long() == d.indices).all().item())
