"""Shared helpers for the parity tests."""
import os
import re


import numpy as np

from uedge_b200.capi import UeLib
from uedge_b200.cases import (box2_case, d3dhsm_case, forthon_case1, initial_profiles, load_grid_npz,
                              load_rate_tables_npz, load_state_npz, refine_grid, refine_state)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "libue_oracle.so")


def oracle():
    """The CPU oracle bound through the same ctypes class as the product (checker only)."""
    return UeLib(ORACLE_LIB, "ue_ora_")


def apply_overrides(c, overrides):
    for k, v in (overrides or {}).items():
        pkg, nm = k.split(".")
        ns = getattr(c, pkg)
        cur = getattr(ns, nm) if nm in ns else None
        if isinstance(cur, np.ndarray) and not isinstance(v, np.ndarray):
            cur = cur.copy()
            cur.flat[0] = v   # species-indexed inputs: species 1
            v = cur
        setattr(ns, nm, v)


def make_slab_case(name, perturb=0.0, seed=1234, overrides=None):
    """`case1`: Forthon_case1 at the steady state its reference output prints (ng: the frozen initial profile);
    `box2d`: pyexamples/box2 with diffusive atoms at ueinit-like smooth profiles."""
    c = forthon_case1() if name == "case1" else box2_case()
    apply_overrides(c, overrides)
    c.setup()
    if name == "case1":
        z = np.load(os.path.join(ROOT, "tests", "golden", "case1_state.npz"))
        yl = c.set_state(z["ni"], z["up"], z["te"], z["ti"], c.initial_ng())
    else:
        yl = c.set_state(*initial_profiles(c))
    if perturb:
        rng = np.random.default_rng(seed)
        yl[: c.bbb.neq] *= 1.0 + perturb * rng.uniform(-1.0, 1.0, c.bbb.neq)
    return c, yl


def make_case(name="d3dHsm", istabon=0, perturb=0.0, seed=1234, overrides=None):
    if name in ("case1", "box2d"):
        return make_slab_case(name, perturb, seed, overrides)
    g = load_grid_npz()
    state = load_state_npz("case2_state.npz" if name == "case2" else "d3dHsm_state.npz")
    m = re.fullmatch(r"d3dHsm(\d+)x", name)
    if m:  # synthetic refinement (BASELINE configs[4] is the 4x one)
        f = int(m.group(1))
        g = refine_grid(g, f, f)
        state = refine_state(state, f, f)
    c = d3dhsm_case(g, istabon=10 if name == "case2" else istabon)
    if overrides:
        for k, v in overrides.items():
            pkg, nm = k.split(".")
            ns = getattr(c, pkg)
            cur = getattr(ns, nm) if nm in ns else None
            if isinstance(cur, np.ndarray) and not isinstance(v, np.ndarray):
                cur = cur.copy()
                cur.flat[0] = v   # species-indexed inputs: species 1
                v = cur
            setattr(ns, nm, v)
    if c.com.istabon == 10:
        c.set_rate_tables(load_rate_tables_npz())
    c.setup()
    yl = c.set_state(*state)
    if perturb:
        rng = np.random.default_rng(seed)
        yl[: c.bbb.neq] *= 1.0 + perturb * rng.uniform(-1.0, 1.0, c.bbb.neq)
    return c, yl


def bind(lib, c):
    lib.load_static(c.static_inputs())
    lib.init()
    return lib


def psetnk_inputs(c, yl):
    """(yl with Jacobian flag, suscal) as psetnk/sfsetnk prepare them (bbb/oderhs.m:9453-9468, 9848-9857)."""
    y = yl.copy()
    y[c.bbb.neq] = 1.0
    return y, c.suscal(yl)


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
