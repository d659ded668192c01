#!/usr/bin/env python
"""Cycles per small tcgen05.mma (run on the GPU box): python profiles/umma_bench.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: F401  (initialises the CUDA context the same way the library's users do)
from sparrowrecsys_b200 import _lib
lib = _lib.load()
out = (C.c_uint64 * 2)()
print("%-9s %-4s %-5s %-6s %-5s %10s %10s %9s" % ("issue", "N", "n", "A", "acc", "issue_cyc", "total_cyc", "cyc/mma"))
for uniform in (0, 1):
    for N in (32, 64, 128):
        for a_tmem in (0, 1):
            for two in (0, 1):
                for n in (12, 192):
                    best = None
                    for rep in range(3):
                        _lib.check(lib.srs_debug_umma_bench(N, n, a_tmem, two | (uniform << 1), 0, out))
                        best = (out[0], out[1]) if best is None or out[1] < best[1] else best
                    print("%-9s %-4d %-5d %-6s %-5s %10d %10d %9.1f" % (
                        "uniform" if uniform else "divergent", N, n, "tmem" if a_tmem else "smem",
                        "two" if two else "one", best[0], best[1], best[1] / n))
