"""Shared helpers for the parity tests."""
import os
import re


import numpy as np

from uedge_b200.capi import UeLib
from uedge_b200.cases import (apply_overrides, box2_case, d3dhsm_case, forthon_case1, initial_profiles, load_grid_npz,  # noqa: F401
                              load_rate_tables_npz, load_state_npz, make_case, make_slab_case, psetnk_inputs, refine_grid,
                              refine_state)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "libue_oracle.so")


def oracle():
    """The CPU oracle bound through the same ctypes class as the product (checker only)."""
    return UeLib(ORACLE_LIB, "ue_ora_")


def bind(lib, c):
    lib.load_static(c.static_inputs())
    lib.init()
    return lib


def csr_to_dense_rows(jac, ja, ia):
    return [(ja[ia[i] - 1 : ia[i + 1] - 1], jac[ia[i] - 1 : ia[i + 1] - 1]) for i in range(len(ia) - 1)]


def newton_solve(lib, c, yl, iters=14):
    """Plain Newton with backtracking on ||f|| using the library's own residual and FD Jacobian
    (direct sparse solve on the host).  Returns (yl*, history of max|f|)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    b = c.bbb
    neq = b.neq
    yl = yl.copy()
    su = c.suscal(yl)
    lib.step_params(np.full(neq, 1e20), yl[:neq], su, np.ones(neq))
    hist = []
    for _ in range(iters):
        y = yl.copy()
        y[neq] = 1.0
        f = lib.pandf1(y)
        jac, ja, ia = lib.jac_calc(y, f, b.lbw, b.ubw, b.nnzmx)
        J = sp.csr_matrix((jac, ja - 1, ia - 1), shape=(neq, neq)).tocsc()
        d = spl.spsolve(J, -f)
        lam, n0 = 1.0, np.linalg.norm(f)
        fn = f
        while lam > 1e-4:
            yn = yl.copy()
            yn[:neq] += lam * d
            if (yn[:neq].reshape(-1, 5)[:, [0, 2, 3, 4]] > 0).all():
                fn = lib.pandf1(yn)
                if np.linalg.norm(fn) < n0:
                    break
            lam *= 0.5
        yl = yn
        hist.append(float(np.abs(fn).max()))
        if hist[-1] < 1e-7:
            break
    return yl, hist
