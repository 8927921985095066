"""Developer tool: entries in which the general path's Jacobian differs from the oracle's (drift case on a refined mesh)."""
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, ".")
from uedge_b200.cases import load_grid_npz, refine_grid  # noqa: E402
from uedge_b200.cases2 import Oracle2, jupyter_case, load_gen  # noqa: E402

f = int(sys.argv[1]) if len(sys.argv) > 1 else 4
c, yl = jupyter_case(grid=refine_grid(load_grid_npz(), f, f))
b = c.bbb
o, g = Oracle2().bind(c), load_gen().bind(c)
fo, fg = o.pandf1(yl), g.pandf1(yl)
jo = o.jac_calc(yl, fo, b.lbw, b.ubw, b.nnzmx)
for rep in range(2):
    jg = g.jac_calc(yl, fg, b.lbw, b.ubw, b.nnzmx)
    print("nnz", len(jo[0]), len(jg[0]), "identical", all(np.array_equal(p, q) for p, q in zip(jo, jg)))
    A = sp.csr_matrix((jo[0], jo[1] - 1, jo[2] - 1), shape=(b.neq, b.neq))
    B = sp.csr_matrix((jg[0], jg[1] - 1, jg[2] - 1), shape=(b.neq, b.neq))
    Ab = sp.csr_matrix((np.ones(len(jo[0])), jo[1] - 1, jo[2] - 1), shape=(b.neq, b.neq))
    Bb = sp.csr_matrix((np.ones(len(jg[0])), jg[1] - 1, jg[2] - 1), shape=(b.neq, b.neq))
    D = (Bb - Ab).tocoo()
    nv = b.numvar
    k = 0
    for r, cc, v in zip(D.row, D.col, D.data):
        if v != 0 and k < 30:
            k += 1
            print(" row cell", (int(c.igyl[r, 0]), int(c.igyl[r, 1])), "var", r % nv, "| col cell", (int(c.igyl[cc, 0]), int(c.igyl[cc, 1])), "var", cc % nv, "| gpu", B[r, cc], "oracle", A[r, cc])
    V = (A - B).tocoo()
    print(" value differences on the common pattern:", int((np.abs(V.data) > 0).sum()))
