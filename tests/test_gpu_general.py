"""GPU parity of the GENERAL path (libuegpu.so, entry points ue_gen_* of include/ue_gen.h) through the C ABI:
  * against the reference's OWN stored vectors (pyexamples/input_example/solution.h5: pandf1 output and ~45 field planes for ten
    subsets of equations on an 8x4 non-orthogonal mesh with inertial atoms and the potential equation);
  * BIT FOR BIT against the general oracle: residual, Jacobian values, ia/ja - input_example subsets, pyexamples/box2 exactly as
    its deck runs it (inertial atoms), and the d3dHsm family with psetnk's scalings and a finite time step."""
import numpy as np
import pytest

from tests.refplanes import check_against_reference
from tests.test_oracle2_golden import twin
from tests.util import psetnk_inputs
from uedge_b200.cases import box2_case
from uedge_b200.cases2 import SUBSETS, Oracle2, all_drifts, braginskii_current, box2_initial_state, d3d_full_physics_case, gas_energy_case, inputex_case, jupyter_case, load_gen, switch_variant

pytestmark = pytest.mark.gpu


def same(o, g, c, yl, step=None):
    b = c.bbb
    if step is not None:
        for lib in (o, g):
            lib.step_params(*step)
    for lib in (o, g):
        lib.pandf1(yl)
    fo, fg = o.pandf1(yl), g.pandf1(yl)
    assert np.array_equal(fo, fg)
    for nm in ("fnix1", "fniy1", "feex", "feiy", "fmix1", "visx1", "hcxe", "resee", "resei", "fngx", "fngy", "fqx", "fqy"):
        assert np.array_equal(o.plane(nm), g.plane(nm)), nm
    jo, jg = o.jac_calc(yl, fo, b.lbw, b.ubw, b.nnzmx), g.jac_calc(yl, fg, b.lbw, b.ubw, b.nnzmx)
    assert len(jo[0]) > b.neq
    assert np.array_equal(jo[2], jg[2]) and np.array_equal(jo[1], jg[1]) and np.array_equal(jo[0], jg[0])
    return jo, jg


@pytest.mark.parametrize("subset", SUBSETS)
def test_cuda_reproduces_the_reference_stored_vectors(built, subset):
    c, yl, gold = inputex_case(subset)
    check_against_reference(load_gen().bind(c), c, gold, subset, yl)


@pytest.mark.parametrize("subset", SUBSETS)
def test_input_example_bit_identical_to_oracle(built, subset):
    c, yl, _ = inputex_case(subset)
    same(Oracle2().bind(c), load_gen().bind(c), c, yl)


def test_box2_as_its_deck_runs_it(built):
    c = box2_case(isupgon=1)  # inertial atoms (box2_in.py:114-131)
    yl = box2_initial_state(c)
    same(Oracle2().bind(c), load_gen().bind(c), c, yl)


@pytest.mark.parametrize("seed", range(24))
def test_switch_combinations(built, seed):
    """random combinations of differencing schemes (0-8 in x and y), flux-limit / viscosity / conductivity options, rate models,
    boundary options and 4th-order terms on the input_example case: bit-identical to the oracle (non-finite values included)."""
    mods, desc = switch_variant(seed)
    c, yl, _ = inputex_case("default", mods=mods)
    b = c.bbb
    o, g = Oracle2().bind(c), load_gen().bind(c)
    for lib in (o, g):
        lib.pandf1(yl)
    fo, fg = o.pandf1(yl), g.pandf1(yl)
    assert np.array_equal(fo, fg, equal_nan=True), desc
    jo, jg = o.jac_calc(yl, fo, b.lbw, b.ubw, b.nnzmx), g.jac_calc(yl, fg, b.lbw, b.ubw, b.nnzmx)
    assert all(np.array_equal(p, q, equal_nan=True) for p, q in zip(jo, jg)), desc


def test_full_physics_on_the_d3d_mesh(built):
    """non-orthogonal stencils + inertial atoms + methg=66 + potential on the 16x8 DIII-D mesh of the headline case (1 260 unknowns,
    X-point cuts with the 5-point stencils): bit-identical to the oracle, also with psetnk's scalings and a finite time step."""
    c, yl = d3d_full_physics_case()
    b = c.bbb
    same(Oracle2().bind(c), load_gen().bind(c), c, yl)
    su = c.suscal(yl)
    step = (np.full(b.neq, 1e-5), 0.9995 * yl[: b.neq], su, 1.0 / np.maximum(np.abs(yl[: b.neq]), 1e-3))
    y = yl.copy(); y[b.neq] = 1.0
    same(Oracle2().bind(c), load_gen().bind(c), c, y, step)


@pytest.mark.parametrize("refine", [1, 2])
def test_jupyter_drift_case(built, refine):
    """BASELINE configs[2] (jupyter/case_setup.py): cross-field drifts, grad-B currents, isnewpot=1, iphibcc=3, Joule heating,
    sheath conditions from the current; 16x8 mesh (1 260 unknowns) and its 2x refinement (4 284): residual, drift planes and the
    Jacobian with the wider band of the potential unknowns, bit for bit"""
    from uedge_b200.cases import load_grid_npz, refine_grid
    c, yl = jupyter_case(grid=refine_grid(load_grid_npz(), refine, refine) if refine > 1 else None)
    o, g = Oracle2().bind(c), load_gen().bind(c)
    same(o, g, c, yl)
    o.pandf1(yl); g.pandf1(yl)
    for nm in ("vyce1", "vycb1", "v2ce1", "v2cb1", "ve2cb", "veycb", "fqyb", "fqxb", "fqyd", "wjdote", "vex", "vey", "resphi"):
        assert np.array_equal(o.plane(nm), g.plane(nm)), nm


@pytest.mark.parametrize("mods", [all_drifts, braginskii_current])
def test_every_drift_part(built, mods):
    """the diamagnetic, resistive, B x grad(T) and classical parts, the diamagnetic and charge-exchange currents switched on as well; the
    radial current of the classical viscosity model"""
    c, yl = jupyter_case(mods)
    same(Oracle2().bind(c), load_gen().bind(c), c, yl)


@pytest.mark.parametrize("deck", ["jupyter", "inputex"])
def test_gas_energy_equation(built, deck):
    """istgon = 1 (engbalg): tg as the eighth unknown per cell on the drift case and on the non-orthogonal input_example mesh"""
    c, yl = gas_energy_case(deck=deck)
    o, g = Oracle2().bind(c), load_gen().bind(c)
    same(o, g, c, yl)
    o.pandf1(yl); g.pandf1(yl)
    for nm in ("fegx", "fegy", "segc", "reseg", "conxge", "floyge"):  # (fegx is 0/0 in the last guard column, where dxnog = 0, as in the reference)
        assert np.array_equal(o.plane(nm), g.plane(nm), equal_nan=True), nm


@pytest.mark.parametrize("name", ["d3dHsm", "case2", "case1"])
def test_d3dhsm_family_through_the_general_path(built, name):
    c1, c2, yl = twin(name)
    b = c2.bbb
    y, su = psetnk_inputs(c1, yl)
    step = (np.full(b.neq, 1e-4), 0.999 * y[: b.neq], su, np.ones(b.neq))
    same(Oracle2().bind(c2), load_gen().bind(c2), c2, y, step)


def test_perturbed_states_and_column_range(built):
    c, yl, _ = inputex_case("default")
    b = c.bbb
    o, g = Oracle2().bind(c), load_gen().bind(c)
    rng = np.random.default_rng(7)
    for rep in range(3):
        y = yl.copy(); y[: b.neq] *= 1.0 + 1e-3 * rng.standard_normal(b.neq)
        jo, jg = same(o, g, c, y)
    # columns 101..250 only (ppp column split): the CSR of that range
    C = __import__("ctypes")
    for lib in (o, g):
        lib._f("set_column_range")(C.c_int64(101), C.c_int64(250))
    fo = o.pandf1(y)
    po, pg = o.jac_calc(y, fo, b.lbw, b.ubw, b.nnzmx), g.jac_calc(y, fo, b.lbw, b.ubw, b.nnzmx)
    assert all(np.array_equal(p, q) for p, q in zip(po, pg)) and 0 < len(po[0]) < len(jo[0])
    assert po[1].min() >= 101 and po[1].max() <= 250


@pytest.mark.parametrize("nblk", [2, 5, 148])
def test_grid_mode_residual_on_a_small_mesh(built, monkeypatch, nblk):
    """the full-domain residual as ONE context spread over a co-resident grid (cooperative launch, grid barriers between the loop
    nests; the default beyond 1 024 cells), forced here onto input_example: residual and planes bit-identical to the oracle, and a
    negative density is reported by every block instead of hanging the grid"""
    monkeypatch.setenv("UE_GEN_FULL_GRID", str(nblk))
    c, yl, _ = inputex_case("default")
    o, g = Oracle2().bind(c), load_gen().bind(c)
    same(o, g, c, yl)
    bad = yl.copy(); bad[7] = -1.0
    with pytest.raises(RuntimeError, match="ni is negative"):
        g.pandf1(bad)
    assert np.array_equal(g.pandf1(yl), o.pandf1(yl))


@pytest.mark.parametrize("case", ["jupyter", "gas_energy", "full_physics"])
def test_grid_mode_residual_on_the_4x_mesh(built, case):
    """2 244 cells: the residual runs in grid mode by default (18 blocks of 128 threads, one cell per thread)"""
    from uedge_b200.cases import load_grid_npz, refine_grid
    grid = refine_grid(load_grid_npz(), 4, 4)
    c, yl = {"jupyter": jupyter_case, "gas_energy": gas_energy_case, "full_physics": d3d_full_physics_case}[case](grid=grid)
    o, g = Oracle2().bind(c), load_gen().bind(c)
    fo, fg = o.pandf1(yl), g.pandf1(yl)
    assert np.array_equal(fo, fg)
    for nm in ("fnix1", "feex", "fmix1", "resee", "resei", "fqx", "fqy", "resphi", "vex", "vey"):
        assert np.array_equal(o.plane(nm), g.plane(nm)), nm
    y2 = yl.copy(); y2[: c.bbb.neq] *= 1.0 + 1e-6
    assert np.array_equal(o.pandf1(y2), g.pandf1(y2))


def test_repeated_jacobians_on_the_4x_mesh(built):
    """the persistent column kernel reuses its private planes from unknown to unknown and from call to call: what an earlier
    evaluation left there (here also: another state) must not reach a kept row.  Drift case on the 4x mesh (15 708 unknowns): three
    Jacobians at two states, each bit-identical to the oracle"""
    from uedge_b200.cases import load_grid_npz, refine_grid
    c, yl = jupyter_case(grid=refine_grid(load_grid_npz(), 4, 4))
    b = c.bbb
    o, g = Oracle2().bind(c), load_gen().bind(c)
    y2 = yl.copy(); y2[: b.neq] *= 1.0 + 3e-4
    ref = {}
    for k, y in enumerate((yl, y2, yl)):
        fg = g.pandf1(y)
        jg = g.jac_calc(y, fg, b.lbw, b.ubw, b.nnzmx)
        if id(y) not in ref:
            fo = o.pandf1(y)
            ref[id(y)] = (fo, o.jac_calc(y, fo, b.lbw, b.ubw, b.nnzmx))
        fo, jo = ref[id(y)]
        assert np.array_equal(fo, fg), k
        assert all(np.array_equal(p, q) for p, q in zip(jo, jg)), (k, len(jo[0]), len(jg[0]))


def test_errors_are_reported(built):
    c, yl, _ = inputex_case("default")
    g = load_gen().bind(c)
    bad = yl.copy(); bad[0] = -1.0
    with pytest.raises(RuntimeError, match="ni is negative"):
        g.pandf1(bad)
    g.pandf1(yl)  # and the library keeps working afterwards
    c.bbb.isimpon = 2
    with pytest.raises(RuntimeError, match="isimpon"):
        load_gen().bind(c)
