for n in 0 4 9 18 37 74 148; do echo "UE_GEN_FULL_GRID=$n"; UE_GEN_FULL_GRID=$n python - <<'PY'
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np
from uedge_b200.cases import load_grid_npz, refine_grid
from uedge_b200.cases2 import jupyter_case, load_gen
for f in (1, 2, 4):
    c, yl = jupyter_case(grid=refine_grid(load_grid_npz(), f, f) if f > 1 else None)
    g = load_gen().bind(c)
    g.pandf1(yl)
    ts = []
    for _ in range(5):
        g.pandf1(yl)
        km = [C.c_double(0) for _ in range(3)]
        g._f("last_kernel_ms")(*[C.byref(x) for x in km]); ts.append(km[0].value)
    print("  mesh %dx: NC %d residual kernel %.3f ms" % (f, (c.com.nx + 2) * (c.com.ny + 2), min(ts)), flush=True)
    g._f("finalize")()
PY
done
