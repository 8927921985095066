"""CPU logic check of the product's GENERAL path (uedge_b200/csrc/ue_gen_phys.h + ue_gen.cu) without a GPU: the same sources built
for the host (tests/hostcheck, -DUE_GEN_HOST) must reproduce the general oracle BIT FOR BIT - residual, Jacobian values, ia/ja -
both with the loop nests in the reference's order and REVERSED.  On the GPU the iterations of one nest run concurrently on the
threads of a warp or block; identical results in both orders show that no iteration reads what another one of the same nest
writes.  Every Jacobian column is evaluated on a private copy of the base planes (as each warp does), not in place as the
reference and the oracle do - so the test also shows that the reference's perturb / restore sequence leaves no trace."""
import os
import subprocess

import numpy as np
import pytest

from tests.test_oracle2_golden import twin
from tests.util import psetnk_inputs
from uedge_b200.cases import box2_case
from uedge_b200.cases2 import SUBSETS, Lib2, Oracle2, all_drifts, braginskii_current, box2_initial_state, d3d_full_physics_case, gas_energy_case, inputex_case, jupyter_case, switch_variant

HK = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck")


def host(rev):
    subprocess.run(["make", "-C", HK, "libuegen_host_rev.so" if rev else "libuegen_host.so"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return Lib2(os.path.join(HK, "libuegen_host_rev.so" if rev else "libuegen_host.so"), "ue_genh_")


def same(o, h, c, yl, step=None):
    b = c.bbb
    if step is not None:
        for lib in (o, h):
            lib.step_params(*step)
    for lib in (o, h):
        lib.pandf1(yl)
    fo, fh = o.pandf1(yl), h.pandf1(yl)
    assert np.array_equal(fo, fh)
    for nm in ("fnix1", "fniy1", "feex", "feiy", "fmix1", "visx1", "hcxe", "resee", "resei"):  # (before the oracle's in-place Jacobian loop)
        assert np.array_equal(o.plane(nm), h.plane(nm)), nm
    jo, jh = o.jac_calc(yl, fo, b.lbw, b.ubw, b.nnzmx), h.jac_calc(yl, fh, b.lbw, b.ubw, b.nnzmx)
    assert len(jo[0]) > b.neq
    assert np.array_equal(jo[2], jh[2]) and np.array_equal(jo[1], jh[1]) and np.array_equal(jo[0], jh[0])


@pytest.mark.parametrize("rev", [0, 1])
@pytest.mark.parametrize("subset", SUBSETS)
def test_input_example_subsets(built, subset, rev):
    c, yl, _ = inputex_case(subset)
    same(Oracle2().bind(c), host(rev).bind(c), c, yl)


@pytest.mark.parametrize("rev", [0, 1])
def test_box2_as_its_deck_runs_it(built, rev):
    c = box2_case(isupgon=1)  # inertial atoms (box2_in.py:114-131)
    yl = box2_initial_state(c)
    assert c.bbb.numvar == 6 and c.bbb.neq == 384
    same(Oracle2().bind(c), host(rev).bind(c), c, yl)


def test_full_physics_on_the_d3d_mesh(built):
    """everything switched on (non-orthogonal, inertial atoms, methg=66, potential) on the 16x8 DIII-D mesh, loop nests reversed"""
    c, yl = d3d_full_physics_case()
    same(Oracle2().bind(c), host(1).bind(c), c, yl)


@pytest.mark.parametrize("seed", range(16))
def test_switch_combinations(built, seed):
    """random combinations of differencing schemes (0-8 in x and y for every equation), flux-limit / viscosity / conductivity options,
    rate models, boundary options and 4th-order terms on the input_example case: residual and Jacobian bit-identical to the oracle
    (loop nests reversed); combinations that drive a scheme out of its domain (e.g. inverse interpolation of a velocity that changes
    sign) must produce the same non-finite values on both sides."""
    mods, desc = switch_variant(seed)
    c, yl, _ = inputex_case("default", mods=mods)
    b = c.bbb
    o, h = Oracle2().bind(c), host(1).bind(c)
    for lib in (o, h):
        lib.pandf1(yl)
    fo, fh = o.pandf1(yl), h.pandf1(yl)
    assert np.array_equal(fo, fh, equal_nan=True), desc
    jo, jh = o.jac_calc(yl, fo, b.lbw, b.ubw, b.nnzmx), h.jac_calc(yl, fh, b.lbw, b.ubw, b.nnzmx)
    assert all(np.array_equal(p, q, equal_nan=True) for p, q in zip(jo, jh)), desc


@pytest.mark.parametrize("poison", ["1", "2"])
def test_band_copy_of_the_private_planes_is_sufficient(built, monkeypatch, poison):
    """(poison 1: everything outside the copy is NaN - a cell the evaluation needs but the copy lacks loses an entry; poison 2: finite
    garbage that differs from unknown to unknown - a stale cell that reaches a kept row adds a spurious entry, which NaN cannot show
    because a NaN element fails the clip test and is dropped.)  Every Jacobian column works on a private copy of the field planes; only a band of rows around the perturbed cell (plus the
    X-point rows and the line arrays) is copied from the base set.  With UE_GEN_POISON the host build fills everything else with NaN
    first: on a refined mesh (32x16, 18 rows; band 11 rows) the Jacobian stays bit-identical to the oracle, so nothing outside the
    band is read by what the band produces."""
    from uedge_b200.cases import load_grid_npz, refine_grid
    monkeypatch.setenv("UE_GEN_POISON", poison)
    c, yl = d3d_full_physics_case(refine_grid(load_grid_npz(), 2, 2))
    assert c.com.ny + 2 == 18
    same(Oracle2().bind(c), host(0).bind(c), c, yl)


@pytest.mark.parametrize("poison", ["1", "2"])
@pytest.mark.parametrize("subset", ["ni-0", "up-0", "te", "phi", "default"])
def test_window_copy_with_few_unknowns_per_cell(built, monkeypatch, subset, poison):
    """the private copy of a column holds the band rows x the window columns (+2), whole rows at the core and wall boundaries and at
    the X-point, and the line arrays; with one unknown per cell the Jacobian band spans the whole 8x4 mesh, so a stale far cell
    (e.g. the corner cells every wall window sets) would show up as an extra entry - everything else is poisoned with NaN here"""
    monkeypatch.setenv("UE_GEN_POISON", poison)
    c, yl, _ = inputex_case(subset)
    same(Oracle2().bind(c), host(0).bind(c), c, yl)


@pytest.mark.parametrize("name, kw", [("iflcore=1", dict(iflcore=1, pcoree=2e5, pcorei=2e5)), ("isnicore=0", dict(isnicore=(0, 0), curcore=(0, 10.0))),
                                      ("iphibcc=1", dict(iphibcc=1)), ("isnewpot=0", dict(isnewpot=0, rnewpot=0.0))])
@pytest.mark.parametrize("poison", ["1", "2"])
def test_window_copy_with_core_boundary_sums(built, monkeypatch, name, kw, poison):
    """core conditions that sum over the whole core boundary (power, current) or set the potential rows at every core column: rows 0-2
    of the private copy hold all core columns; 2x-refined drift case, everything else poisoned"""
    from uedge_b200.cases import load_grid_npz, refine_grid
    monkeypatch.setenv("UE_GEN_POISON", poison)

    def m(b, com):
        for k, v in kw.items():
            if isinstance(v, tuple):
                a = np.asarray(getattr(b, k)).copy(); a[v[0]] = v[1]; setattr(b, k, a)
            else:
                setattr(b, k, v)
    c, yl = jupyter_case(m, grid=refine_grid(load_grid_npz(), 2, 2))
    same(Oracle2().bind(c), host(0).bind(c), c, yl)


@pytest.mark.parametrize("rev", [0, 1])
def test_jupyter_drift_case(built, rev):
    """jupyter/case_setup.py: ExB and grad-B drifts, grad-B currents, isnewpot=1 with its two core conditions (one of them the sum of
    the radial current over the whole core boundary), Joule heating, the wider Jacobian band of the potential unknowns"""
    c, yl = jupyter_case()
    assert c.bbb.numvar == 7 and c.bbb.neq == 1260
    same(Oracle2().bind(c), host(rev).bind(c), c, yl)


@pytest.mark.parametrize("rev", [0, 1])
def test_every_drift_part(built, rev, monkeypatch):
    """diamagnetic (cfydd, cf2dd), resistive (cfrd) and B x grad(T) (cfbgt) parts, the diamagnetic currents (cfjpy, cfjp2), the classical
    momentum-transfer velocity and conductivities (cfvycr, cfrtaue, cfeta1, cfcl_e, cfcl_i) and the charge-exchange current (cfqyn) on top
    of the deck's ExB / grad-B set; 2x-refined mesh with the private copies poisoned in the forward order"""
    from uedge_b200.cases import load_grid_npz, refine_grid
    if rev == 0:
        monkeypatch.setenv("UE_GEN_POISON", "1")
    c, yl = jupyter_case(all_drifts, grid=refine_grid(load_grid_npz(), 2, 2) if rev == 0 else None)
    same(Oracle2().bind(c), host(rev).bind(c), c, yl)


@pytest.mark.parametrize("rev", [0, 1])
def test_braginskii_radial_current(built, rev, monkeypatch):
    """cfvycf: the radial current from the classical viscosity velocity and its core potential conditions"""
    monkeypatch.setenv("UE_GEN_POISON", "1")
    c, yl = jupyter_case(braginskii_current)
    same(Oracle2().bind(c), host(rev).bind(c), c, yl)


@pytest.mark.parametrize("poison", ["1", "2"])
def test_jupyter_drift_case_band_copy_is_sufficient(built, monkeypatch, poison):
    """the same on the 2x-refined mesh with everything outside the copied band poisoned (NaN): the core conditions of the
    potential read rows 0-2 at every core column whenever the window starts at iy <= 3 - they are inside the band then"""
    from uedge_b200.cases import load_grid_npz, refine_grid
    monkeypatch.setenv("UE_GEN_POISON", poison)
    c, yl = jupyter_case(grid=refine_grid(load_grid_npz(), 2, 2))
    assert c.com.ny + 2 == 18 and c.bbb.neq == 4284
    same(Oracle2().bind(c), host(0).bind(c), c, yl)


@pytest.mark.parametrize("rev", [0, 1])
@pytest.mark.parametrize("deck", ["jupyter", "inputex"])
def test_gas_energy_equation(built, deck, rev):
    """istgon = 1 (engbalg, oderhs.m:7508-7878): tg as the eighth unknown per cell on the drift case (orthogonal DIII-D mesh) and on
    input_example (non-orthogonal: the fegxy term)"""
    c, yl = gas_energy_case(deck=deck)
    assert c.bbb.numvar == 8
    same(Oracle2().bind(c), host(rev).bind(c), c, yl)


@pytest.mark.parametrize("k, opts", list(enumerate([(1, 1, 1, 1, 0), (2, 2, 3, 3, 2), (3, 3, 4, 4, 3), (4, 4, 5, 5, 1), (5, 5, 0, 4, 0)])))
def test_gas_energy_boundary_options(built, k, opts):
    """every wall / plate / core option of the gas temperature (boundary.m:769-852, 1463-1513, 2198-2257, 2880-2937)"""
    pfc, wc, lb, rb, core = opts

    def m(b, com):
        for nm, v in (("istgpfc", pfc), ("istgwc", wc), ("istgcore", core)):
            a = np.asarray(getattr(b, nm)).copy(); a[0] = v; setattr(b, nm, a)
        b.istglb = lb; b.istgrb = rb; b.recyce = 0.3; b.recycwe = 0.2; b.lytg = np.full(12, 0.05)
        b.matwsi = np.asarray(b.matwsi).copy(); b.matwsi[0] = 1
    for deck in ("jupyter", "inputex"):
        c, yl = gas_energy_case(m, deck=deck)
        same(Oracle2().bind(c), host(k & 1).bind(c), c, yl)


@pytest.mark.parametrize("poison", ["1", "2"])
def test_gas_energy_band_copy_is_sufficient(built, monkeypatch, poison):
    from uedge_b200.cases import load_grid_npz, refine_grid
    monkeypatch.setenv("UE_GEN_POISON", poison)
    c, yl = gas_energy_case(grid=refine_grid(load_grid_npz(), 2, 2))
    assert c.bbb.neq == 4896
    same(Oracle2().bind(c), host(0).bind(c), c, yl)


@pytest.mark.parametrize("name", ["d3dHsm", "case2", "case1"])
def test_d3dhsm_family_through_the_general_path(built, name):
    """the d3dHsm family (one ion species, diffusive atoms, orthogonal mesh; DEGAS2 tables for case2) with psetnk's scalings and a
    finite time step: here oracle2 is itself bit-identical to the kernels' oracle (tests/test_oracle2_golden.py)."""
    c1, c2, yl = twin(name)
    b = c2.bbb
    y, su = psetnk_inputs(c1, yl)
    step = (np.full(b.neq, 1e-4), 0.999 * y[: b.neq], su, np.ones(b.neq))
    same(Oracle2().bind(c2), host(1).bind(c2), c2, y, step)


def test_errors_are_reported(built):
    c, yl, _ = inputex_case("default")
    h = host(0).bind(c)
    bad = yl.copy(); bad[0] = -1.0  # negative density
    with pytest.raises(RuntimeError, match="ni is negative"):
        h.pandf1(bad)
    c.bbb.isimpon = 2
    with pytest.raises(RuntimeError, match="isimpon"):
        host(0).bind(c)
