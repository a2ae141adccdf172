import sys, time, torch
sys.path.insert(0, ".")
from raft_b200.stats import silhouette_score
from raft_b200.common import DeviceResources
h = DeviceResources()
n, k, nl = 100000, 128, 10
lab = torch.randint(0, nl, (n,), device="cuda", dtype=torch.int32)
x = torch.randn(n, k, device="cuda") + lab[:, None].float()
for metric in ("sqeuclidean", "euclidean", "sqeuclidean_unexpanded", "cosine"):
    silhouette_score(x, lab, nl, metric=metric, handle=h); torch.cuda.synchronize()
    t = time.perf_counter(); v = silhouette_score(x, lab, nl, metric=metric, handle=h); torch.cuda.synchronize()
    print(f"silhouette {n}x{k}, {nl} labels, {metric}: {(time.perf_counter()-t)*1e3:.1f} ms  score {v:.4f}")
from raft_b200.stats import trustworthiness_score
emb = (x @ torch.randn(k, 2, device="cuda")).contiguous()
for nn_ in (5, 15):
    trustworthiness_score(x, emb, n_neighbors=nn_, handle=h); torch.cuda.synchronize()
    t = time.perf_counter(); v = trustworthiness_score(x, emb, n_neighbors=nn_, handle=h); torch.cuda.synchronize()
    print(f"trustworthiness {n}x{k} -> 2-d, n_neighbors {nn_}: {(time.perf_counter()-t)*1e3:.1f} ms  score {v:.4f}")
