"""Developer tool: wall-clock of the general path (ue_gen_*) against the general oracle on the same host, per case."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from tests.test_oracle2_golden import twin  # noqa: E402
from tests.util import psetnk_inputs  # noqa: E402
from uedge_b200.cases import box2_case  # noqa: E402
from uedge_b200.cases import load_grid_npz, refine_grid  # noqa: E402
from uedge_b200.cases2 import Oracle2, box2_initial_state, d3d_full_physics_case, inputex_case, jupyter_case, load_gen  # noqa: E402


def cases():
    c, yl, _ = inputex_case("default"); yield "input_example", c, yl
    c = box2_case(isupgon=1); yield "box2 (inertial atoms)", c, box2_initial_state(c)
    c1, c2, yl = twin("d3dHsm"); y, su = psetnk_inputs(c1, yl); yield "d3dHsm via general path", c2, y
    c, yl = d3d_full_physics_case(); yield "d3d mesh, full physics", c, yl
    c, yl = jupyter_case(); yield "jupyter drift case 16x8", c, yl
    c, yl = jupyter_case(grid=refine_grid(load_grid_npz(), 4, 4)); yield "jupyter drift case 4x", c, yl
    c, yl = d3d_full_physics_case(refine_grid(load_grid_npz(), 4, 4)); yield "d3d 4x mesh, full physics", c, yl
    if "--8x" in sys.argv:
        c, yl = jupyter_case(grid=refine_grid(load_grid_npz(), 8, 8)); yield "jupyter drift case 8x", c, yl


for name, c, yl in (cases() if __name__ == "__main__" else ()):
    b = c.bbb
    g, o = load_gen().bind(c), Oracle2().bind(c)
    f = g.pandf1(yl); o.pandf1(yl)
    g.jac_calc(yl, f, b.lbw, b.ubw, b.nnzmx)
    n = 10
    t = time.perf_counter()
    for _ in range(n):
        g.pandf1(yl)
    tr = (time.perf_counter() - t) / n
    t = time.perf_counter()
    for _ in range(n):
        j = g.jac_calc(yl, f, b.lbw, b.ubw, b.nnzmx)
    tj = (time.perf_counter() - t) / n
    import ctypes as C
    km = [C.c_double(0) for _ in range(3)]
    g._f("last_kernel_ms")(*[C.byref(x) for x in km])
    t = time.perf_counter(); o.pandf1(yl); tro = time.perf_counter() - t
    t = time.perf_counter(); o.jac_calc(yl, f, b.lbw, b.ubw, b.nnzmx); tjo = time.perf_counter() - t
    print("%-26s neq %5d nnz %6d | GPU residual %.3f ms, Jacobian %.3f ms (kernels: residual %.3f, columns %.3f, CSR %.3f) | oracle (1 thread) residual %.3f ms, "
          "Jacobian %.2f ms | x%.1f" % (name, b.neq, len(j[0]), tr * 1e3, tj * 1e3, km[0].value, km[1].value, km[2].value, tro * 1e3, tjo * 1e3, tjo / tj), flush=True)
