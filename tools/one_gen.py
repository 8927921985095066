"""Developer tool: a few residual + Jacobian calls of the general path (target for ncu captures)."""
import sys

sys.path.insert(0, ".")
from tools.time_general import cases  # noqa: E402
from uedge_b200.cases2 import load_gen  # noqa: E402

want = sys.argv[1] if len(sys.argv) > 1 else "input_example"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
for name, c, yl in cases():
    if not name.startswith(want):
        continue
    b = c.bbb
    g = load_gen().bind(c)
    for i in range(n):
        f = g.pandf1(yl)
        j = g.jac_calc(yl, f, b.lbw, b.ubw, b.nnzmx)
    print(name, "nnz", len(j[0]))
