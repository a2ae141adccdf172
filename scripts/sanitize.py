"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python scripts/sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from raft_b200.distance import pairwise_distance, fused_l2_nn, fused_distance_nn
from raft_b200.neighbors import brute_force
from raft_b200.stats import silhouette_score
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g)
x, y = r(300, 96), r(20000, 96)
i, v = fused_l2_nn(x, y, sqrt=False)                       # screened search: sample / trial / main / exact
print("nn screened", int(i.sum()), float(v.sum()))
i, v = fused_l2_nn(r(96, 128), r(40000, 128), sqrt=False)  # single-tile work items, many per SM
i, v = fused_distance_nn(r(200, 96) + 0.5, r(20000, 96) + 0.5, metric="correlation")   # screened, cosine family
i, v = fused_l2_nn(r(300, 40), r(700, 40))                 # exact kernel only
i, v = fused_distance_nn(r(130, 200), r(300, 200), metric="cosine")   # streaming (k > 128) arg-min
d = pairwise_distance(r(257, 100), r(515, 100), metric="sqeuclidean")  # TMA-store epilogue? n % 4 != 0 -> direct
d = pairwise_distance(r(256, 128), r(512, 128), metric="euclidean")    # TMA-store epilogue
d = pairwise_distance(r(100, 700), r(90, 700), metric="sqeuclidean")   # K-chunked
d = pairwise_distance(r(130, 33), r(70, 33), metric="cityblock")       # SIMT loader
d = pairwise_distance(r(128, 64), r(256, 64), metric="chebyshev")      # TMA-fed FP32 kernel
dd, ii = brute_force.knn(r(3000, 32), r(200, 32), k=8)
dd, ii = brute_force.knn(r(2000, 70) + 0.3, r(100, 70) + 0.3, k=5, metric="cosine")
print("knn", int(ii.sum()))
lab = torch.randint(0, 4, (500,), device="cuda", dtype=torch.int32, generator=g)
print("silhouette", silhouette_score(r(500, 16), lab, 4))
from raft_b200.stats import trustworthiness_score
xx = r(600, 20)
print("trustworthiness", trustworthiness_score(xx, xx[:, :3].contiguous(), n_neighbors=7))
# round 2: k <= 64 store path (3-D TMA map), 2-CTA kernel, per-row exponents, fp64, ratio metrics, BrayCurtis, argmin,
# kNN overflow repair (database ordered by decreasing distance)
d = pairwise_distance(r(256, 64), r(512, 64), metric="sqeuclidean")
from raft_b200 import _lib
_lib.lib().b2d_set_option(b"pairwise_2cta", 1.0)
d = pairwise_distance(r(384, 128), r(768, 128), metric="sqeuclidean")
_lib.lib().b2d_set_option(b"pairwise_2cta", 0.0)
xo = r(300, 96); xo[7] *= 1e-9; xo[9] *= 1e6
d = pairwise_distance(xo, r(260, 96), metric="sqeuclidean")
d = pairwise_distance(r(70, 33).double(), r(90, 33).double(), metric="correlation")
b = (r(200, 40) > 0).float()
d = pairwise_distance(b, (r(130, 40) > 0).float(), metric="jaccard")
d = pairwise_distance(r(130, 33).abs(), r(70, 33).abs(), metric="braycurtis")
from raft_b200.matrix import argmin
print("argmin", int(argmin(r(300, 777)).sum()))
q = r(64, 32); base = r(5000, 32)
order = torch.argsort(torch.cdist(q[:1], base)[0], descending=True)
dd, ii = brute_force.knn(base[order].contiguous(), q, k=16)
print("knn ordered", int(ii.sum()))
torch.cuda.synchronize()
print("done")
