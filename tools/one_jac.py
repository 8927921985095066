"""Developer tool: a few residual + Jacobian calls (target for ncu captures)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from tests.util import make_case, psetnk_inputs, bind
from uedge_b200.capi import load_gpu
name = sys.argv[1] if len(sys.argv) > 1 else "d3dHsm"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c, yl = make_case(name, perturb=1e-3)
gpu = bind(load_gpu(), c)
b = c.bbb
y, su = psetnk_inputs(c, yl)
gpu.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
for i in range(n):
    y2 = y.copy(); y2[: b.neq] *= 1 + 1e-7 * i
    f = gpu.pandf1(y2)
    j = gpu.jac_calc(y2, f, b.lbw, b.ubw, b.nnzmx)
print("nnz", len(j[0]))
