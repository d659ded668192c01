#!/usr/bin/env python
"""Phase timeline (SM cycles, CTA 0) of the row-tile DIN kernel (csrc/din_rt.cu).
Run on the GPU box:  python profiles/trace_din_rt.py [batch]"""
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
assert m.kernel_name in ("din_rt_kernel", "din_rtp_kernel"), m.kernel_name   # SRS_DIN_IMPL=rtp: the pipelined kernel
if len(sys.argv) > 2:
    m.set_sm_limit(int(sys.argv[2]))
db = m.to_device(synthetic_features(spec, B, seed=1))
out = torch.empty(B, dtype=torch.float32, device="cuda:0")
lib = _lib.load()
for _ in range(5):
    m.predict_device(db, out)
torch.cuda.synchronize()
lib.srs_debug_din_trace(m._h, 1, None)
buf = (C.c_uint64 * 40)()
names = {0: "entry", 1: "prologue", 2: "phase0", 3: "first_d1_ready", 4: "first_tile_pooled", 5: "tiles_done",
         6: "layer1", 7: "group", 8: "exit"}
if m.kernel_name == "din_rtp_kernel":   # slots: 3/10 first d1 of consumer 0/1, 4/11 consumers done, 5 first pooled group at the top MLP,
    names = {0: "entry", 1: "prologue", 21: "g.reg_dec", 20: "g.first_tile_issued", 24: "g.tile0_delivered", 25: "g.tile6_delivered",
             3: "c0_first_d1", 10: "c1_first_d1", 26: "i.entry", 27: "i.mma1(0)", 28: "b.tile0_built", 29: "b.tile3_built", 17: "b.tile4_built", 15: "i.mma1(4)", 12: "c0.tile4_d1", 13: "c0.tile4_gate",
             14: "c0.tile4_pooled_prev", 16: "i.pool(4)", 18: "i.mma1(6)", 5: "top.g0_pooled_ready", 6: "top.g0_layer1", 22: "top.g0_epi1",
             23: "top.g0_layer2", 7: "top.g0_done", 4: "c0_done", 11: "c1_done", 8: "exit"}
for rep in range(3):
    m.predict_device(db, out)
    _lib.check(lib.srs_debug_din_trace(m._h, 1, buf))
    t = np.array(buf[:], dtype=np.int64)
    prev = t[0]
    line = []
    for i in sorted(names, key=lambda i: t[i]):
        if t[i] == 0:
            continue
        line.append("%s=%d(+%d)" % (names[i], t[i] - t[0], t[i] - prev))
        prev = t[i]
    print(" ".join(line))
    fine = {10: "c.d1_ready", 11: "c.gate", 12: "c.synced", 13: "c.mma2_issued", 14: "c.mma1_next_issued",
            15: "c.d2_ready", 16: "c.pooled", 20: "p.tile0_full", 21: "p.tile2_full", 22: "p.tile4_full",
            23: "p.tile2_built", 17: "c.d2_ready(warp1)", 18: "c.synced(warp1)", 24: "x_built", 25: "l1_issued", 30: "p0.rows_requested", 31: "p0.ids_stored", 32: "p0.pads_zeroed", 33: "p0.sync1", 34: "p0.gathers_issued", 35: "p0.cand_stored"}
    print("   second tile of consumer 0 / producer 0 (cycles since entry): " +
          " ".join("%s=%d" % (fine[i], t[i] - t[0]) for i in sorted(fine) if t[i]))

if m.kernel_name == "din_rtp_kernel":
    tl = (C.c_uint64 * 768)()
    _lib.check(lib.srs_debug_din_timeline(m._h, tl))
    a = np.array(tl[:], dtype=np.int64).reshape(12, 64)
    t0 = int(np.array(buf[:], dtype=np.int64)[0])
    kinds = ["issued", "delivered", "B_built", "mma1", "c_d1", "gate_done", "pool_mma", "pooled", "i.a_full", "i.waits", "i.mmas", "i.iter"]
    print("per-tile timeline of CTA 0 (cycles since kernel entry)")
    print("tile " + " ".join("%9s" % k for k in kinds))
    for K in range(64):
        if a[0, K] == 0:
            break
        print("%4d " % K + " ".join("%9d" % (a[i, K] - t0 if a[i, K] else -1) for i in range(12)))
