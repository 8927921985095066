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
from uedge_b200.cases2 import SUBSETS, Oracle2, inputex_case

TOL = 5.0e-9


def _planes(o, gold, c):
    b = c.bbb
    isn, isu = c.isn, c.isu
    out = []  # (name, ours, gold[iy, ix])
    T = lambda a: a.T
    for nm in ("fnix", "fniy", "visx", "visy", "hcxij", "hcyij"):
        for s in range(2):
            out.append((nm + str(s + 1), o.plane(nm + str(s + 1)), T(gold[nm][:, :, s])))
    for nm in ("hcxe", "hcye", "hcxi", "hcyi", "conxe", "conye", "conxi", "conyi", "floxe", "floye", "floxi", "floyi", "conxg", "conyg", "floxg", "floyg"):
        out.append((nm, o.plane(nm), T(gold[nm])))
    for nm in ("fngx", "fngy"):
        out.append((nm, o.plane(nm), T(gold[nm][:, :, 0])))
    for s in range(2):
        if isu[s]:
            for nm in ("fmix", "fmiy"):
                out.append((nm + str(s + 1), o.plane(nm + str(s + 1)), T(gold[nm][:, :, s])))
    if any(isu):  # scratch planes of the momentum equation hold the last species evaluated
        for nm in ("conx", "cony", "flox", "floy"):
            out.append((nm, o.plane(nm), T(gold[nm])))
    if int(b.isteon):
        out += [("feex", o.plane("feex"), T(gold["feex"])), ("feey", o.plane("feey"), T(gold["feey"]))]
    if int(b.istion):
        out += [("feix", o.plane("feix"), T(gold["feix"])), ("feiy", o.plane("feiy"), T(gold["feiy"]))]
    return out


@pytest.mark.parametrize("subset", SUBSETS)
def test_reference_stored_planes(built, subset):
    c, yl, gold = inputex_case(subset)
    assert int(gold["numvar"]) == c.bbb.numvar
    assert np.array_equal(gold["igyl"][: c.bbb.neq], c.igyl)  # the reference's own unknown ordering (0-based cell indices)
    o = Oracle2().bind(c)
    o.pandf1(yl)
    f = o.pandf1(yl)  # second call: the module state (e.g. upi at ix = nx+1, set by bouncon) is that of a running code
    worst = 0.0
    for name, ours, g in _planes(o, gold, c):
        scale = np.abs(g).max()
        if scale == 0:  # e.g. the conductivities of the atoms: identically zero in the reference too
            assert np.abs(ours).max() == 0, name
            continue
        err = np.abs(ours - g).max() / scale
        worst = max(worst, err)
        assert err <= TOL, "%s/%s: %.3g" % (subset, name, err)
    # residual planes = divergences of the fluxes above: compare on the flux scale
    flux = dict(resco=max(np.abs(gold["fnix"]).max(), np.abs(gold["fniy"]).max()), resmo=max(np.abs(gold["fmix"]).max(), np.abs(gold["fmiy"]).max()),
                resee=max(np.abs(gold["feex"]).max(), np.abs(gold["feey"]).max()), resei=max(np.abs(gold["feix"]).max(), np.abs(gold["feiy"]).max()))
    for s in range(2):
        if c.isn[s]:
            assert np.abs(o.plane("resco%d" % (s + 1)) - gold["resco"][:, :, s].T).max() <= TOL * flux["resco"], subset
        if c.isu[s]:
            assert np.abs(o.plane("resmo%d" % (s + 1)) - gold["resmo"][:, :, s].T).max() <= 2 * TOL * flux["resmo"], subset
    if int(c.bbb.isteon):
        assert np.abs(o.plane("resee") - gold["resee"].T).max() <= TOL * flux["resee"], subset
    if int(c.bbb.istion):
        assert np.abs(o.plane("resei") - gold["resei"].T).max() <= TOL * flux["resei"], subset
    if int(c.bbb.isphion):  # resphi = factor * (sum of currents): compare on the scale of the currents
        fac = c.bbb.nurlxp * c.bbb.dx0 ** 2 / c.bbb.sigbar0
        cur = max(np.abs(o.plane("fqx")).max(), np.abs(o.plane("fqy")).max())
        d = np.abs(o.plane("resphi") - gold["resphi"].T) / fac
        assert d.max() <= TOL * cur, "%s resphi %.3g of %.3g A" % (subset, d.max(), cur)
        if subset == "phi":  # away from the large parallel currents (rows 2, 3) the stored residual itself is reproduced
            a, g = o.plane("resphi")[2:4, 1:9], gold["resphi"].T[2:4, 1:9]
            assert np.corrcoef(a.ravel(), g.ravel())[0, 1] > 0.95
    gy = gold["yldot"][: c.bbb.neq]
    if subset == "te":  # the one subset whose stored output vector is not a converged ~0: 43 non-zero rows up to 2.2e4
        assert np.abs(gy).max() > 2.0e4 and np.count_nonzero(gy) >= 40
        assert np.abs(f - gy).max() <= 2.0e-8 * np.abs(gy).max()
        assert abs(np.sqrt(np.sum(f * f)) - float(gold["fnrm"])) <= 1e-8 * float(gold["fnrm"])  # fnrm as stored (sfscal = 1)


def _twin(name):
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
    c1, c2, yl = _twin(name)
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
