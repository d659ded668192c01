#!/usr/bin/env python
"""Back-to-back launch stress of din_rt64_kernel (direct launches, then a CUDA graph of them)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from sparrowrecsys_b200.features import synthetic_features
from sparrowrecsys_b200.model import CTRModel
from sparrowrecsys_b200.spec import default_spec
from sparrowrecsys_b200.weights import init_weights

V = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
spec = default_spec("din", emb_dim=64, hist_len=200, n_movies=V, n_users=5000)
m = CTRModel(spec, init_weights(spec, 2), 0)
print("kernel", m.kernel_name, flush=True)
dbs = [m.to_device(synthetic_features(spec, B, seed=s, uniform_history=True)) for s in range(4)]
out = torch.empty(B, dtype=torch.float32, device="cuda:0")
for n in (1, 10, 200):
    t0 = time.time()
    for i in range(n):
        m.predict_device(dbs[i % 4], out)
    torch.cuda.synchronize()
    print("direct x%d ok %.1f ms" % (n, (time.time() - t0) * 1e3), flush=True)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for i in range(24):
            m.predict_device(dbs[i % 4], out, stream=s)
    for rep in range(20):
        g.replay()
    s.synchronize()
print("graph ok", flush=True)
print("status", m.status(), flush=True)
