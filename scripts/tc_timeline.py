"""Dump the clock64 milestone timeline of the tcgen05 GEMM (CTA 0,0) for a few shapes."""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
import llama3_np_b200  # noqa
from llama3_np_b200 import _cabi

lib = _cabi.lib()
for rows, n, k, bf in [(256, 288, 288, 0), (256, 288, 768, 0), (256, 1536, 288, 0), (256, 288, 288, 1), (256, 32000, 288, 0)]:
    rng = np.random.default_rng(0)
    x = rng.standard_normal((rows, k)).astype(np.float32)
    w = rng.standard_normal((n, k)).astype(np.float32)
    out = np.empty((rows, n), np.float32)
    for it in range(2):
        lib.l3_debug_tc_timeline(0, 1, None)
        assert lib.l3_op_linear(0, _cabi.f32p(x), _cabi.f32p(w), rows, n, k, 3, bf, _cabi.f32p(out)) == 0
    buf = (C.c_uint64 * 64)()
    lib.l3_debug_tc_timeline(0, 0, buf)
    t = np.array(list(buf), dtype=np.int64)
    t0 = t[0]
    rel = lambda i: int(t[i] - t0) if t[i] else None
    print(f"rows={rows} n={n} k={k} bf16={bf}")
    print("  prologue done", rel(1))
    print("  producer issued kb:", [rel(2 + i) for i in range(12)])
    print("  mma saw full kb  :", [rel(16 + i) for i in range(12)])
    print("  mma committed acc", rel(30), " epi: accbar", rel(32), " phase1", rel(33), " phase2", rel(34), " end", rel(35))
