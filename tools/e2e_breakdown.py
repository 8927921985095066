"""Developer tool: where the end-to-end step spends its time (host clock inside the two entry points vs device kernel times)."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import Gpu, bench_state  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "d3dHsm"
c, ys, su = bench_state(name)
g = Gpu(c, ys, su, 1, 0, None, torch, split=False)
lib = g.lib
lib.ue_gpu_timing.argtypes = [C.POINTER(C.c_double)] * 3 + [C.c_int64]
for h, label in ((g.pin, "pinned"), (g.pag, "pageable")):
    for _ in range(20):
        g.step_e2e(h)
    t = [C.c_double(0) for _ in range(3)]
    lib.ue_gpu_timing(*[C.byref(x) for x in t], 1)
    n = 500
    t0 = time.perf_counter()
    for _ in range(n):
        g.step_e2e(h)
    dt = (time.perf_counter() - t0) / n
    lib.ue_gpu_timing(*[C.byref(x) for x in t], 0)
    jm, rm = C.c_double(0), C.c_double(0)
    lib.ue_gpu_last_kernel_ms.argtypes = [C.POINTER(C.c_double)] * 2
    lib.ue_gpu_last_kernel_ms(C.byref(jm), C.byref(rm))
    print("%s %s: step %.1f us = pandf1 %.1f + jac_calc %.1f + other calls / Python %.1f   (no L2 flush between steps)" %
          (name, label, dt * 1e6, t[0].value / n * 1e6, t[1].value / n * 1e6, (dt - (t[0].value + t[1].value) / n) * 1e6))
g.close()
