"""Latency of the ranking tail on the GPU box: `srs_topk_device` alone (CUDA events) and one
whole `srs_rank_host` request (800 candidates, NeuralCF on the reference's shipped weights;
host clock around the synchronous call).  Writes gpurun_out/rank_latency_r01.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_golden_weights                      # noqa: E402
from sparrowrecsys_b200.model import CTRModel                 # noqa: E402
from sparrowrecsys_b200.ranking import topk_device            # noqa: E402
from sparrowrecsys_b200.spec import default_spec              # noqa: E402

out = {"topk_device_us": {}, "gpu": torch.cuda.get_device_name(0)}
for n, k in ((800, 10), (4096, 100), (65536, 100), (1 << 20, 100)):
    s = torch.rand(n, device="cuda")
    for _ in range(5):
        topk_device(s, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        topk_device(s, k)
    e1.record()
    torch.cuda.synchronize()
    out["topk_device_us"]["n=%d,k=%d" % (n, k)] = round(e0.elapsed_time(e1) * 1e3 / reps, 2)

spec = default_spec("neuralcf")
f = {"movieId": np.arange(1, 801, dtype=np.int32), "userId": np.full(800, 10351, np.int32)}
with CTRModel(spec, load_golden_weights("neuralcf_002")) as m:
    for _ in range(20):
        m.rank(f, 10)
    lat = []
    for _ in range(300):
        t0 = time.perf_counter()
        m.rank(f, 10)
        lat.append((time.perf_counter() - t0) * 1e6)
    out["rank_host_neuralcf_800_us"] = {"median": round(float(np.median(lat)), 1),
                                        "p99": round(float(np.percentile(lat, 99)), 1),
                                        "note": "includes Python encode_batch of the feature dict"}
    lat = []
    for _ in range(300):
        t0 = time.perf_counter()
        p = m.predict(f)
        np.argsort(-p[:, 0], kind="stable")[:10]
        lat.append((time.perf_counter() - t0) * 1e6)
    out["predict_then_host_sort_800_us"] = {"median": round(float(np.median(lat)), 1)}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "rank_latency_r01.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out))
