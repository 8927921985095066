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
