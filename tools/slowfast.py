"""Developer tool: time of the all-ix (X-point cut) windows and of the ordinary ones in the column kernel, separately."""
import ctypes as C
import os
import sys

sys.path.insert(0, ".")
from uedge_b200.cases import load_grid_npz, refine_grid  # noqa: E402
from uedge_b200.cases2 import jupyter_case, load_gen  # noqa: E402

f = int(sys.argv[1]) if len(sys.argv) > 1 else 4
c, yl = jupyter_case(grid=refine_grid(load_grid_npz(), f, f) if f > 1 else None)
b = c.bbb
for mode in ("all", "slow", "fast"):
    if mode == "all":
        os.environ.pop("UE_GEN_DEBUG_LIST", None)
    else:
        os.environ["UE_GEN_DEBUG_LIST"] = mode
    g = load_gen().bind(c)
    f0 = g.pandf1(yl)
    ts = []
    for _ in range(3):
        try:
            g.jac_calc(yl, f0, b.lbw, b.ubw, b.nnzmx)
        except Exception as e:  # noqa: BLE001
            pass
        km = [C.c_double(0) for _ in range(3)]
        g._f("last_kernel_ms")(*[C.byref(x) for x in km]); ts.append(km[1].value)
    print("mesh %dx neq %d: %s columns %.3f ms" % (f, b.neq, mode, min(ts)), flush=True)
    g._f("finalize")()
