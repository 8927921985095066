# Developer tool: column kernel of the drift case with fewer resident warps (per-warp latency = time x slots / unknowns)
for f in 4 8; do for s in 592 1184 2368; do echo "mesh ${f}x slots $s"; UE_GEN_SLOTS=$s python - <<PY
import sys, ctypes as C
sys.path.insert(0, ".")
from uedge_b200.cases import load_grid_npz, refine_grid
from uedge_b200.cases2 import jupyter_case, load_gen
c, yl = jupyter_case(grid=refine_grid(load_grid_npz(), $f, $f))
b = c.bbb
g = load_gen().bind(c)
f0 = g.pandf1(yl)
ts = []
for _ in range(2):
    g.jac_calc(yl, f0, b.lbw, b.ubw, b.nnzmx)
    km = [C.c_double(0) for _ in range(3)]
    g._f("last_kernel_ms")(*[C.byref(x) for x in km]); ts.append(km[1].value)
t = min(ts)
print("   columns %.2f ms, per-warp latency %.2f ms per unknown" % (t, t * $s / b.neq), flush=True)
PY
done; done
