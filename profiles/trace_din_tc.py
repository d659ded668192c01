#!/usr/bin/env python
"""Print the phase timeline (SM cycles) of one worker of the tensor-core DIN kernel.
Run on the GPU box:  python profiles/trace_din_tc.py [batch]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparrowrecsys_b200 import _lib
from sparrowrecsys_b200.features import synthetic_features
from sparrowrecsys_b200.model import CTRModel
from sparrowrecsys_b200.spec import baseline_spec
from sparrowrecsys_b200.weights import init_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
spec = baseline_spec("cfg3_din")
m = CTRModel(spec, init_weights(spec, 2), 0)
db = m.to_device(synthetic_features(spec, B, seed=1))
out = torch.empty(B, dtype=torch.float32, device="cuda:0")
lib = _lib.load()
for _ in range(5):
    m.predict_device(db, out)
torch.cuda.synchronize()
lib.srs_debug_din_trace(m._h, 1, None)
buf = (C.c_uint64 * 40)()
for rep in range(3):
    m.predict_device(db, out)
    _lib.check(lib.srs_debug_din_trace(m._h, 1, buf))
    t = np.array(buf[:], dtype=np.int64)
    t0 = t[0]
    names = {0: "entry", 1: "prologue", 2: "phase0", 3: "weights", 30: "layer1", 31: "group", 32: "exit"}
    line = []
    prev = t0
    fine = {20: "t1.built", 21: "t1.stwait", 22: "t1.sync", 23: "t1.issued", 24: "t1.prefetched",
            25: "t1.mma_done", 26: "t1.epilogue"}
    print("  tile-1 detail: " + " ".join("%s=%d" % (fine[i], t[i] - t[5 - 1]) for i in sorted(fine) if t[i]))
    for i in list(range(0, 4 + 8)) + [30, 31, 32]:
        if t[i] == 0:
            continue
        line.append("%s=%d(+%d)" % (names.get(i, "tile%d" % (i - 4)), t[i] - t0, t[i] - prev))
        prev = t[i]
    print(" ".join(line))
