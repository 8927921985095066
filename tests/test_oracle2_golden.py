"""The general oracle (oracle/ue_oracle2.cpp) against the reference's OWN stored vectors, and the d3dHsm-family oracle
(oracle/ue_oracle.cpp, the bit-for-bit twin of the CUDA kernels) against the general one.

Pin 1 — pyexamples/input_example/solution.h5, groups pytests/<subset> (exported to tests/golden/inputex_solution.npz by
tools/make_golden.py): for ten subsets of equations the reference stored its pandf1 output `yldot` and ~45 intermediate
planes of the residual evaluation on an 8x4 NON-ORTHOGONAL single-null mesh with INERTIAL atoms (nhsp=2, isupgon=1),
methg=66 and the POTENTIAL equation.  Every plane the subset computes must agree to 5e-9 of the plane's largest value
(the stencil weights of the non-orthogonal mesh come out of geometric line intersections, which the reference evaluated
with -Ofast: that sets a noise floor of ~1e-11..1e-10 on everything interpolated with them; planes that do not touch
the stencils, e.g. hcxij, agree to 4e-16).  Residuals are differences of fluxes: they are compared on the flux scale.

Pin 2 — on the d3dHsm family (one ion species, diffusive atoms, orthogonal mesh) the two oracles must agree BIT FOR BIT:
residual, Jacobian values, ia/ja.  Together with the GPU tests (CUDA == ue_oracle.cpp bit for bit) this ties the
kernels to arithmetic that reproduces the reference's stored output.
"""
import numpy as np
import pytest

from tests.util import bind, make_case, oracle, psetnk_inputs
from uedge_b200.case2 import Case2
from uedge_b200.cases import box2_case, d3dhsm_case, forthon_case1, load_grid_npz, load_rate_tables_npz, load_state_npz
from tests.refplanes import check_against_reference
from uedge_b200.cases2 import SUBSETS, Oracle2, inputex_case



@pytest.mark.parametrize("subset", SUBSETS)
def test_reference_stored_planes(built, subset):
    c, yl, gold = inputex_case(subset)
    assert int(gold["numvar"]) == c.bbb.numvar
    assert np.array_equal(gold["igyl"][: c.bbb.neq], c.igyl)  # the reference's own unknown ordering (0-based cell indices)
    check_against_reference(Oracle2().bind(c), c, gold, subset, yl)


def twin(name):
    """(v1 case, v2 case, yl) for a d3dHsm-family configuration."""
    c1, yl = make_case(name, perturb=1e-3)
    if name in ("d3dHsm", "case2"):
        c2 = d3dhsm_case(load_grid_npz(), istabon=10 if name == "case2" else 0, cls=Case2)
        if name == "case2":
            c2.set_rate_tables(load_rate_tables_npz())
        st = load_state_npz("case2_state.npz" if name == "case2" else "d3dHsm_state.npz")
    elif name == "case1":
        c2 = forthon_case1(cls=Case2)
        st = None
    c2.setup()
    if st is None:
        st = (c1.ni, c1.up, c1.te, c1.ti, c1.ng)
    c2.set_state2([st[0]], [st[1]], st[2], st[3], ng=st[4])
    return c1, c2, yl


@pytest.mark.parametrize("name", ["d3dHsm", "case2", "case1"])
def test_d3dhsm_family_oracle_is_bit_identical_to_the_general_one(built, name):
    c1, c2, yl = twin(name)
    b = c1.bbb
    assert c2.bbb.neq == b.neq and np.array_equal(c2.igyl, c1.igyl) and np.array_equal(c2.iseqalg, c1.iseqalg)
    o1 = bind(oracle(), c1)
    o2 = Oracle2().bind(c2)
    y, su = psetnk_inputs(c1, yl)
    for o in (o1, o2):
        o.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    f1, f2 = o1.pandf1(y), o2.pandf1(y)
    assert np.array_equal(f1, f2)
    j1 = o1.jac_calc(y, f1, b.lbw, b.ubw, b.nnzmx)
    j2 = o2.jac_calc(y, f2, b.lbw, b.ubw, b.nnzmx)
    assert np.array_equal(j1[2], j2[2]) and np.array_equal(j1[1], j2[1]) and np.array_equal(j1[0], j2[0])


# ---- cross-field drifts + the new potential model: jupyter/PyUedge.ipynb (BASELINE configs[2]) -------------------------------
def _jupyter_fnrm0(mods=None):
    """fnrm0 as nksol prints it for jupyter/case_setup.py at the state of jupyter/d3d.hdf5: sfsetnk (row max-norms of J diag(1/su),
    one residual + one full Jacobian with the ExtendedJacPhi band), then |f sf|."""
    from uedge_b200.cases2 import Oracle2, jupyter_case
    c, yl = jupyter_case(mods)
    b = c.bbb
    o = Oracle2().bind(c)
    y = yl.copy(); y[b.neq] = 1.0
    su = c.suscal(yl)
    o.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    f0 = o.pandf1(y)
    jac, ja, ia = o.jac_calc(y, f0, b.lbw, b.ubw, b.nnzmx)
    rows = np.repeat(np.arange(b.neq), np.diff(ia))
    sf = np.zeros(b.neq)
    np.maximum.at(sf, rows, np.abs(jac * (1.0 / su)[ja - 1]))
    f = o.pandf1(yl)
    assert np.isfinite(f).all() and (sf > 0).all()
    t = ((f / sf) ** 2).reshape(c.com.ny + 2, c.com.nx + 2, b.numvar)
    return float(np.sqrt(t.sum())), t, (c, o, jac, ja, ia)


def test_jupyter_drift_case_fnrm0(built):
    """The one number the reference tree holds for a case with cross-field drifts: cell 17 of jupyter/PyUedge.ipynb prints
    `iter= 0 fnrm= 2.134077960622300` for jupyter/case_setup.py (ExB + grad-B drifts, grad-B currents, isnewpot=1, iphibcc=3,
    Joule heating, sheath conditions from the current) restarted from jupyter/d3d.hdf5.  It passes through one residual and one
    full Jacobian.  The oracle gives 2.2970 (+7.6 %) on the 16x8 mesh of builder/test/facets/gridue (the notebook regenerates its
    own mesh).  What the difference is made of:
      * 1.73 of the 2.30 sits in the last cell column before the outer plate and 0.99 in its guard column - the atoms' density
        and parallel-velocity rows and the ion-energy row there, i.e. the recycling model at the plate, not the drifts;
      * every drift switch moves fnrm0 by 0.002-0.04 only (cfyef +0.014, cf2ef +0.035, cfybf +0.031, cf2bf -0.002 when switched
        off; isnewpot=0 +0.28), so a wrong drift term could not hide in the 0.16 gap, nor produce it;
      * the plate-energy correction that is newer than the notebook (cfloxiplt = 0: recycled atoms carry no power back,
        bbb.v:397) accounts for 0.066 of it: with cfloxiplt = 1 (no correction) fnrm0 = 2.2314 (+4.6 %);
      * recycm (momentum recycling of the atoms at the plates, -0.9 today) moves it by -0.36 when set to 0.
    Same situation as Forthon_case2 (test_case2_fnrm0_breakdown): the neutral model moved on, the remaining gap is in neutral
    rows at the plates."""
    v, t, _ = _jupyter_fnrm0()
    assert abs(v - 2.2969676) < 2e-4
    assert abs(v - 2.134077960622300) < 0.08 * 2.134
    col = np.sqrt(t.sum(axis=(0, 2)))
    assert col[16] > 1.7 and col[17] > 0.95 and np.sqrt((col[1:16] ** 2).sum()) < 1.0
    def era(b, com):
        b.cfloxiplt = 1.0
    v1, _, _ = _jupyter_fnrm0(era)
    assert abs(v1 - 2.134077960622300) < 0.05 * 2.134


def test_jupyter_drift_terms_are_in_the_residual(built):
    """Every drift coefficient of the deck changes the residual and the Jacobian pattern holds the potential band and the dense
    current row of the midplane core cell (boundary.m:1040-1068: the sum of the radial current over the core boundary)."""
    from uedge_b200.cases2 import Oracle2, jupyter_case
    c, yl = jupyter_case()
    b = c.bbb
    o = Oracle2().bind(c)
    f0 = o.pandf1(yl).copy()
    v2ce, vyce, vycb, fqyb, fqxb, wj = [o.plane(n).copy() for n in ("v2ce1", "vyce1", "vycb1", "fqyb", "fqxb", "wjdote")]
    assert all(np.abs(p[1:-1, 1:-1]).max() > 0 for p in (v2ce, vyce, vycb, fqyb, fqxb, wj))
    # ExB drift = E x B / B^2: v2ce = (dphi/dy) / B on the x-faces, from the vertex potentials
    g = c.geo
    phiv = o.plane("phiv")
    iy, ix = 4, 10
    ix2 = int(c.ixp1[iy, ix])
    want = 2.0 * (phiv[iy, ix] - phiv[iy - 1, ix]) * g["gyc"][iy, ix] / (g["btot"][iy, ix] + g["btot"][iy, ix2])
    assert abs(v2ce[iy, ix] - want) <= 1e-14 * abs(want)
    for k in ("cfyef", "cf2ef", "cfybf", "cf2bf", "cfqybf", "cfq2bf", "jhswitch", "isnewpot"):
        def off(bb, com, k=k):
            setattr(bb, k, 0 if k in ("jhswitch", "isnewpot") else 0.0)
        c2, yl2 = jupyter_case(off)
        f2 = Oracle2().bind(c2).pandf1(yl2)
        assert not np.array_equal(f0, f2), k
    jac, ja, ia = o.jac_calc(yl, f0, b.lbw, b.ubw, b.nnzmx)
    nv = b.numvar
    iv_mid = int(c.idx["idxphi"][0, c.com.ixmp])  # phi(ixmp, 0), 1-based
    cols = ja[ia[iv_mid - 1] - 1 : ia[iv_mid] - 1]
    cells = {(int(c.igyl[k - 1, 0]), int(c.igyl[k - 1, 1])) for k in cols}
    core = range(c.com.ixpt1 + 1, c.com.ixpt2 + 1)
    assert sum(1 for i in core if any((i, j) in cells for j in (0, 1, 2))) >= 0.5 * len(core), sorted(cells)
    # ExtendedJacPhi: a potential unknown reaches rows up to four mesh rows away; without it the band is the usual one
    def noext(bb, com):
        bb.ExtendedJacPhi = 0
    c3, yl3 = jupyter_case(noext)
    o3 = Oracle2().bind(c3)
    j3 = o3.jac_calc(yl3, o3.pandf1(yl3), b.lbw, b.ubw, b.nnzmx)
    assert len(j3[0]) <= len(jac)
