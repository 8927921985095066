"""GPU parity tests for the slab family (BASELINE configs[0] lineage): analytic slab mesh (mhdgeo=-1, idealgrd), ix=0 as a
symmetry plane (isfixlb=2), four unknowns per cell (isngon=0: frozen atom density) and the package-default rate fits
(istabon=7).  `case1` is builder/test/Forthon_cases/Forthon_case1 at the state its reference output prints.
Everything is compared bit for bit with the CPU oracle, as in test_gpu_parity.py."""
import numpy as np
import pytest

from tests.test_gpu_parity import _check_jac, _jac_pair
from tests.util import bind, make_case, oracle
from uedge_b200.capi import UeError, load_gpu

pytestmark = pytest.mark.gpu


def _pair(name, perturb, overrides=None, seed=1234):
    c, yl = make_case(name, perturb=perturb, overrides=overrides, seed=seed)
    return c, yl, bind(load_gpu(), c), bind(oracle(), c)


@pytest.mark.parametrize("perturb", [0.0, 1e-3, 0.05])
def test_case1_residual_parity(built, perturb):
    c, yl, gpu, ora = _pair("case1", perturb)
    assert c.bbb.numvar == 4 and c.bbb.neq == 384
    fg, fo = gpu.pandf1(yl), ora.pandf1(yl)
    assert np.isfinite(fg).all() and fg.size == 384
    assert np.array_equal(fg, fo), "%d of %d entries differ" % ((fg != fo).sum(), fg.size)


@pytest.mark.parametrize("perturb", [0.0, 1e-2])
def test_case1_jacobian_parity(built, perturb):
    c, yl, gpu, ora = _pair("case1", perturb)
    jg, jo, noise = _jac_pair(c, yl, gpu, ora)
    assert len(jo[0]) > 3000
    _check_jac(jg, jo, noise)


def test_case1_jacobian_with_timestep_term(built):
    c, yl, gpu, ora = _pair("case1", 1e-3, overrides={"bbb.isbcwdt": 1})
    n = c.bbb.neq
    rng = np.random.default_rng(11)
    dt = 10.0 ** rng.uniform(-6, -3, n)
    yo = yl[:n] * (1 + 1e-2 * rng.uniform(-1, 1, n))
    for lib in (gpu, ora):
        lib.set_real("dtreal", 1e-4)
        lib.step_params(dt, yo, np.ones(n), np.ones(n))
    y = yl.copy(); y[n] = -1.0
    assert np.array_equal(gpu.pandf1(y), ora.pandf1(y))
    jg, jo, noise = _jac_pair(c, yl, gpu, ora, dt=dt)
    assert all(np.array_equal(p, q) for p, q in zip(jg, jo))
    for lib in (gpu, ora):
        lib.set_real("dtreal", 1e20)


@pytest.mark.parametrize("ov", [
    {"com.istabon": 7},                        # Campbell fits on the tokamak mesh
    {"bbb.isngon": 0},                         # frozen atoms (numvar = 4) on the tokamak mesh, X-point cuts included
    {"bbb.isngon": 0, "com.istabon": 7, "bbb.isbcwdt": 1},
])
def test_d3dhsm_variants_of_the_new_switches(built, ov):
    c, yl, gpu, ora = _pair("d3dHsm", 1e-3, overrides=ov)
    assert c.bbb.numvar == (4 if ov.get("bbb.isngon", 1) == 0 else 5)
    fg, fo = gpu.pandf1(yl), ora.pandf1(yl)
    assert np.array_equal(fg, fo)
    _check_jac(*_jac_pair(c, yl, gpu, ora))


def test_case1_column_split_and_one_call_form(built):
    """ppp column split (ue_gpu_set_column_range) and ue_gpu_rhs_jac with four unknowns per cell."""
    c, yl, gpu, ora = _pair("case1", 1e-3)
    b = c.bbb
    jg, jo, noise = _jac_pair(c, yl, gpu, ora)
    y = yl.copy(); y[b.neq] = 1.0
    f, (jac, ja, ia) = gpu.rhs_jac(y, b.lbw, b.ubw, b.nnzmx)
    assert np.array_equal(jac, jo[0]) and np.array_equal(ja, jo[1]) and np.array_equal(ia, jo[2])
    half = b.neq // 2
    parts = []
    for lo, hi in ((1, half), (half + 1, b.neq)):
        gpu.set_column_range(lo, hi)
        fg = gpu.pandf1(y)
        parts.append(gpu.jac_calc(y, fg, b.lbw, b.ubw, b.nnzmx))
    gpu.set_column_range(1, b.neq)
    from uedge_b200.split import merge_csr
    jac2, ja2, ia2 = merge_csr(parts, b.neq)
    assert np.array_equal(jac2, jo[0]) and np.array_equal(ja2, jo[1]) and np.array_equal(ia2, jo[2])


@pytest.mark.parametrize("perturb,seed", [(0.0, 1), (1e-2, 2), (0.2, 3)])
def test_box2_diffusive_variant_parity(built, perturb, seed):
    """pyexamples/box2 (slab 6x6 with a core region, iysptrx=2, core power condition iflcore=1, symmetry plane) with the
    deck's physics except the neutral model: diffusive atoms (isupgon=0, isngon=1) instead of inertial ones.  Exercises
    the half-space cut at ixpt2: fluxes/gradients forced to zero (oderhs.m:2447-2466) and up -> 0 rows
    (boundary.m:1772-1785)."""
    c, yl, gpu, ora = _pair("box2d", perturb, seed=seed)
    assert (c.com.nx, c.com.ny, c.com.ixpt1, c.com.ixpt2, c.com.iysptrx) == (6, 6, -1, 3, 2) and c.bbb.neq == 320
    fg, fo = gpu.pandf1(yl), ora.pandf1(yl)
    assert np.isfinite(fo).all() and np.array_equal(fg, fo), "%d of %d entries differ" % ((fg != fo).sum(), fg.size)
    jg, jo, noise = _jac_pair(c, yl, gpu, ora)
    assert len(jo[0]) > 3000
    _check_jac(jg, jo, noise)


@pytest.mark.parametrize("ov", [
    {"bbb.isupcore": 1, "bbb.iflcore": 0},
    {"bbb.isngon": 0, "com.istabon": 7},
    {"bbb.methn": 22, "bbb.methu": 22, "bbb.methe": 22, "bbb.methi": 22, "bbb.methg": 22, "bbb.isbcwdt": 1},
    {"bbb.xlinc": 3, "bbb.xrinc": 2, "bbb.yinc": 3},
])
def test_box2_switch_variants(built, ov):
    c, yl, gpu, ora = _pair("box2d", 0.05, overrides=ov, seed=7)
    fg, fo = gpu.pandf1(yl), ora.pandf1(yl)
    assert np.array_equal(fg, fo)
    _check_jac(*_jac_pair(c, yl, gpu, ora))


def test_box2_rough_state(built):
    """Flow reversals and steep cell-to-cell jumps around the cut."""
    c, yl = make_case("box2d")
    n, nv = c.bbb.neq, c.bbb.numvar
    rng = np.random.default_rng(5)
    y = yl[:n].reshape(-1, nv).copy()
    y[:, [0, 2, 3, 4]] *= 3.0 ** rng.uniform(-1, 1, (y.shape[0], 4))
    y[:, 1] = (np.abs(y[:, 1]) + 0.05) * rng.choice([-1.0, 1.0], y.shape[0]) * rng.uniform(0.2, 2.0, y.shape[0])
    yl[:n] = y.reshape(-1)
    gpu, ora = bind(load_gpu(), c), bind(oracle(), c)
    fg, fo = gpu.pandf1(yl), ora.pandf1(yl)
    assert np.array_equal(fg, fo)
    _check_jac(*_jac_pair(c, yl, gpu, ora))


WALL_BC_SETS = [
    # the wall conditions of pyexamples/input_example/input.py:30-48 and builder/test/level_2/rdd3d_cfdupg.8x4
    {"bbb.istepfc": 3, "bbb.lyte": np.array([0.03, 0.03]), "bbb.matwso": 1, "bbb.recycw": 0.9, "bbb.isnwcono": 1, "bbb.isnwconi": 1,
     "bbb.nwallo": 1.0e18, "bbb.nwalli": 1.0e18},
    # extrapolated wall temperatures (second interior row enters the guard rows), gradient-length wall densities
    {"bbb.istepfc": 2, "bbb.istipfc": 2, "bbb.istewc": 2, "bbb.istiwc": 2, "bbb.isnwcono": 3, "bbb.isnwconi": 3, "bbb.lyni": np.array([0.05, 0.02])},
    {"bbb.istipfc": 3, "bbb.istiwc": 3, "bbb.istewc": 3, "bbb.lyte": np.array([0.02, 0.04]), "bbb.lyti": np.array([0.05, 0.03]),
     "bbb.matwso": 1, "bbb.matwsi": 1, "bbb.recycw": -0.5, "bbb.isrefluxclip": 0},
    {"bbb.matwso": 1, "bbb.matwsi": 1, "bbb.recycw": -2.0, "bbb.albdso": 0.9, "bbb.albdsi": 0.8},
    # core-boundary variants: second-derivative velocity, extrapolated atoms; extrapolated wall densities
    {"bbb.isupcore": 2, "bbb.isngcore": 3, "bbb.isnwcono": 2, "bbb.isnwconi": 2},
    {"bbb.isupcore": 3, "bbb.isngcore": 2, "bbb.iflcore": -1},
    {"bbb.isngcore": 1, "bbb.ngcore": 2.0e15, "bbb.isupcore": 1},
    {"bbb.isngcore": 4, "bbb.iflcore": 0},
]


@pytest.mark.parametrize("name", ["d3dHsm", "box2d"])
@pytest.mark.parametrize("k", range(len(WALL_BC_SETS)))
def test_wall_boundary_condition_variants(built, name, k):
    c, yl, gpu, ora = _pair(name, 0.02, overrides=WALL_BC_SETS[k], seed=20 + k)
    fg, fo = gpu.pandf1(yl), ora.pandf1(yl)
    assert np.isfinite(fo).all() and np.array_equal(fg, fo), "%d of %d entries differ" % ((fg != fo).sum(), fg.size)
    # the variant must actually change guard rows
    c0, yl0 = make_case(name, perturb=0.02, seed=20 + k)
    assert not np.array_equal(bind(oracle(), c0).pandf1(yl0), fo)
    bind(ora, c)
    _check_jac(*_jac_pair(c, yl, gpu, ora))


@pytest.mark.parametrize("name,model_dt,isbcwdt", [("d3dHsm", 0, 0), ("d3dHsm", 1, 1), ("d3dHsm", 2, 0), ("d3dHsm", 3, 1), ("case1", 1, 0), ("box2d", 3, 1)])
def test_set_dt_on_device(built, name, model_dt, isbcwdt):
    """set_dt of the nksol driver (bbb/oderhs.m:9886-10147): f0 = rhsnk(yl) and the per-unknown pseudo time step, left on the
    device for the residual / Jacobian calls that follow."""
    c, yl, gpu, ora = _pair(name, 1e-2, overrides={"bbb.model_dt": model_dt, "bbb.isbcwdt": isbcwdt, "bbb.dtreal": 1e-5})
    n = c.bbb.neq
    rng = np.random.default_rng(31)
    yo = yl[:n] * (1 + 1e-2 * rng.uniform(-1, 1, n)) + 1e-3  # no exact zeros: a zero ylodt entry gives dtuse = 0
    for lib in (gpu, ora):
        lib.step_params(np.full(n, 1e20), yo, np.ones(n), np.ones(n))
    y = yl.copy(); y[n] = -1.0
    (fg, dg), (fo, do) = gpu.set_dt(y), ora.set_dt(y)
    assert np.array_equal(fg, fo) and np.array_equal(dg, do)
    assert np.isfinite(do).all() and (do > 0).all()
    if model_dt == 0:
        assert set(np.unique(do)) <= {1e-5, 1e20}
    else:
        assert len(np.unique(do)) > n // 4
    # the residual and the Jacobian that follow use the dtuse the call left behind (no step_params in between)
    assert np.array_equal(gpu.pandf1(y), ora.pandf1(y))
    y[n] = 1.0
    fg, fo = gpu.pandf1(y), ora.pandf1(y)
    b = c.bbb
    jg, jo = gpu.jac_calc(y, fg, b.lbw, b.ubw, b.nnzmx), ora.jac_calc(y, fo, b.lbw, b.ubw, b.nnzmx)
    assert all(np.array_equal(p, q) for p, q in zip(jg, jo))
    # and a step_params call with the returned vector is recognised as unchanged / gives the same rows
    for lib in (gpu, ora):
        lib.step_params(do, yo, np.ones(n), np.ones(n))
    y[n] = -1.0
    assert np.array_equal(gpu.pandf1(y), ora.pandf1(y))


@pytest.mark.parametrize("name", ["case1", "box2d"])
@pytest.mark.parametrize("seed", range(6))
def test_slab_fuzz(built, name, seed):
    """Coefficient / array-input / integer-switch fuzz on the slab family (same inputs as the CPU logic check)."""
    from tests.test_hostcheck import run_fuzzed
    run_fuzzed(load_gpu(), name, seed)


def test_switch_change_after_init_needs_reinit(built):
    """Lists, derived flags and refusals are built from the switches at ue_gpu_init: a changed integer switch afterwards
    disables the entry points until the caller re-initialises; model_dt (read by ue_gpu_set_dt only) is patched in place."""
    c, yl = make_case("d3dHsm", perturb=1e-3)
    gpu = bind(load_gpu(), c)
    f = gpu.pandf1(yl)
    gpu.set_int("isupcore", int(c.bbb.isupcore[0]))   # unchanged: nothing happens
    assert np.array_equal(gpu.pandf1(yl), f)
    gpu.set_int("model_dt", 2)                         # allowed at run time
    assert np.array_equal(gpu.pandf1(yl), f)
    gpu.set_int("model_dt", int(c.bbb.model_dt))
    gpu.set_int("isupcore", 1 - int(c.bbb.isupcore[0]))
    with pytest.raises(UeError, match="ue_gpu_init not called"):
        gpu.pandf1(yl)
    bind(gpu, c)                                       # re-initialise with the original switch set
    assert np.array_equal(gpu.pandf1(yl), f)
