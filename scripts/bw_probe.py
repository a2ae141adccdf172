import torch, time
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/it
n = 10_000_000_000
x = torch.empty(n, dtype=torch.float32, device="cuda")
ms = t(lambda: x.fill_(1.0)); print(f"fill 40GB: {ms:.3f} ms {4*n/ms/1e6:.0f} GB/s")
ms = t(lambda: x.zero_()); print(f"zero 40GB: {ms:.3f} ms {4*n/ms/1e6:.0f} GB/s")
y = torch.empty(n//2, dtype=torch.float32, device="cuda")
ms = t(lambda: y.copy_(x[:n//2])); print(f"copy 20GB->20GB: {ms:.3f} ms {2*4*(n//2)/ms/1e6:.0f} GB/s (r+w)")
ms = t(lambda: x[:n//2].sum()); print(f"read 20GB: {ms:.3f} ms {4*(n//2)/ms/1e6:.0f} GB/s")
