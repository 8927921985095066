"""CPU check of the kernels' LOGIC (no GPU): the device functions of uedge_b200/csrc/ue_device.cuh compiled for the host
(tests/hostcheck) and driven by loops that mirror the kernels of ue_gpu.cu, compared bit for bit with the oracle.  This is a
developer safety net for the build container, which has no GPU; the parity tests proper are the `-m gpu` ones."""
import os
import subprocess

import numpy as np
import pytest

from tests.util import ROOT, bind, make_case, oracle, psetnk_inputs
from uedge_b200.capi import UeLib

HK_DIR = os.path.join(ROOT, "tests", "hostcheck")


@pytest.fixture(scope="module")
def hk(built):
    subprocess.check_call(["make", "-s", "-C", HK_DIR])
    return UeLib(os.path.join(HK_DIR, "libue_hostcheck.so"), "ue_hk_")


def _check(c, yl, hk, jac=True, dt=None):
    ora = bind(oracle(), c)
    bind(hk, c)
    b = c.bbb
    n = b.neq
    fo, fh = ora.pandf1(yl), hk.pandf1(yl)
    assert np.isfinite(fo).all() and np.array_equal(fo, fh), "%d of %d residual entries differ" % ((fo != fh).sum(), n)
    if not jac:
        return
    y, su = psetnk_inputs(c, yl)
    for lib in (ora, hk):
        lib.step_params(np.full(n, 1e20) if dt is None else dt, y[:n], su, np.ones(n))
    fo, fh = ora.pandf1(y), hk.pandf1(y)
    assert np.array_equal(fo, fh)
    jo, jh = ora.jac_calc(y, fo, b.lbw, b.ubw, b.nnzmx), hk.jac_calc(y, fh, b.lbw, b.ubw, b.nnzmx)
    assert np.array_equal(jo[2], jh[2]), "ia differs: nnz %d vs %d" % (len(jo[0]), len(jh[0]))
    assert np.array_equal(jo[1], jh[1]) and np.array_equal(jo[0], jh[0])


CASES = [
    ("d3dHsm", {}),
    ("d3dHsm", {"com.istabon": 7, "bbb.isngon": 0}),
    ("case1", {}),
    ("case1", {"bbb.isbcwdt": 1}),
    ("box2d", {}),
    ("box2d", {"bbb.isupcore": 1, "bbb.iflcore": 0, "bbb.xlinc": 3, "bbb.xrinc": 2, "bbb.yinc": 3}),
    ("box2d", {"bbb.isngon": 0, "com.istabon": 7, "bbb.methn": 22, "bbb.methu": 22, "bbb.methe": 22, "bbb.methi": 22, "bbb.methg": 22}),
]


@pytest.mark.parametrize("name,ov", CASES)
def test_kernel_logic_matches_oracle(hk, name, ov):
    c, yl = make_case(name, perturb=0.03, overrides=ov, seed=17)
    rng = np.random.default_rng(2)
    _check(c, yl, hk, dt=10.0 ** rng.uniform(-6, -2, c.bbb.neq) if ov.get("bbb.isbcwdt") else None)


def test_kernel_logic_wall_boundary_conditions(hk):
    from tests.test_gpu_slab import WALL_BC_SETS
    for name in ("d3dHsm", "box2d"):
        for k, ov in enumerate(WALL_BC_SETS):
            c, yl = make_case(name, perturb=0.02, overrides=ov, seed=20 + k)
            _check(c, yl, hk)


def test_kernel_logic_set_dt(hk):
    for name, model_dt, isbcwdt in (("d3dHsm", 1, 1), ("case1", 2, 0), ("box2d", 3, 1)):
        c, yl = make_case(name, perturb=1e-2, overrides={"bbb.model_dt": model_dt, "bbb.isbcwdt": isbcwdt, "bbb.dtreal": 1e-5})
        ora = bind(oracle(), c)
        bind(hk, c)
        n = c.bbb.neq
        yo = yl[:n] * 1.01 + 1e-3
        for lib in (ora, hk):
            lib.step_params(np.full(n, 1e20), yo, np.ones(n), np.ones(n))
        y = yl.copy(); y[n] = -1.0
        (fo, do), (fh, dh) = ora.set_dt(y), hk.set_dt(y)
        assert np.array_equal(fo, fh) and np.array_equal(do, dh)
        assert np.array_equal(ora.pandf1(y), hk.pandf1(y))


def fuzzed_slab_inputs(name, seed):
    """The coefficient / array-input / integer-switch fuzz of tests/test_gpu_parity.py applied to a slab-family case:
    scaled real coefficients, +-10 % noise on every geometry plane and 1-D array, a dozen random integer switches."""
    from tests.test_gpu_parity import _FUZZ_SKIP, _FUZZ_ZERO_OK, _INT_CHOICES
    rng = np.random.default_rng(13000 + seed)
    c, yl = make_case(name, perturb=5e-3, seed=200 + seed)
    s = c.static_inputs()
    names = [k for k, v in s["reals"].items() if k not in _FUZZ_SKIP | {"deldt", "nwimin", "nwomin"} and v != 0.0 and abs(v) < 1e15]
    for k in rng.choice(names, size=min(60, len(names)), replace=False):
        s["reals"][k] = float(s["reals"][k]) * float(rng.uniform(0.8, 1.25))
    for k in rng.choice(_FUZZ_ZERO_OK, size=6, replace=False):
        if s["reals"][k] == 0.0:
            s["reals"][k] = float(rng.uniform(0.05, 0.3))
    for k in s["planes"]:
        a = np.array(s["planes"][k], dtype=np.float64)
        s["planes"][k] = a * rng.uniform(0.9, 1.1, a.shape)
    for k in ("fgtdx", "fgtdy", "flalfea", "flalfia", "flalfva", "flalfgxa", "flalfgxya", "flalfgya", "tewalli", "tiwalli", "tewallo", "tiwallo",
              "alblb", "albrb", "albedoi", "albedoo"):
        a = np.array(s["lines"][k], dtype=np.float64)
        s["lines"][k] = np.minimum(a * rng.uniform(0.9, 1.1, a.shape), np.where(a <= 1.0, 1.0, np.inf))
    picked = {}
    for k in rng.choice(sorted(_INT_CHOICES), size=12, replace=False):
        picked[str(k)] = int(rng.choice(_INT_CHOICES[str(k)]))
        s["ints"][str(k)] = picked[str(k)]
    if s["ints"]["iflcore"] == 1:
        s["reals"]["pcoree"] = s["reals"]["pcorei"] = 2.5e4
    return c, yl, s, picked, 10.0 ** rng.uniform(-6, -2, c.bbb.neq)


def run_fuzzed(lib, name, seed):
    """Residual and Jacobian of `lib` against the oracle on fuzzed_slab_inputs (bit for bit)."""
    c, yl, s, picked, dt = fuzzed_slab_inputs(name, seed)
    ora = oracle()
    ora.load_static(s)
    try:
        ora.init()
    except Exception:
        pytest.skip("combination refused: %s" % picked)
    lib.load_static(s); lib.init()
    n = c.bbb.neq
    fo, fh = ora.pandf1(yl), lib.pandf1(yl)
    if not np.isfinite(fo).all():
        pytest.skip("non-physical combination: %s" % picked)
    assert np.array_equal(fo, fh), "%s: %d residual entries differ" % (picked, (fo != fh).sum())
    y, su = psetnk_inputs(c, yl)
    for l in (ora, lib):
        l.step_params(dt, y[:n], su, np.ones(n))
    f1, f2 = ora.pandf1(y), lib.pandf1(y)
    jo = ora.jac_calc(y, f1, c.bbb.lbw, c.bbb.ubw, c.bbb.nnzmx)
    jh = lib.jac_calc(y, f2, c.bbb.lbw, c.bbb.ubw, c.bbb.nnzmx)
    assert all(np.array_equal(p, q) for p, q in zip(jo, jh)), picked


@pytest.mark.parametrize("name", ["case1", "box2d"])
@pytest.mark.parametrize("seed", range(6))
def test_kernel_logic_fuzz_slab(hk, name, seed):
    run_fuzzed(hk, name, seed)


@pytest.mark.parametrize("seed", range(10))
def test_kernel_logic_switch_combinations(hk, seed):
    """Random combinations of the single-switch variants of test_gpu_parity.py with the wall / core boundary-condition sets of
    test_gpu_slab.py, on the tokamak mesh and on the box, with the time-step term and nufak on."""
    from tests.test_gpu_parity import VARIANTS
    from tests.test_gpu_slab import WALL_BC_SETS
    rng = np.random.default_rng(17000 + seed)
    ov = {}
    for nme in rng.choice(sorted(VARIANTS), size=int(rng.integers(2, 5)), replace=False):
        ov.update(VARIANTS[str(nme)])
    ov.update(WALL_BC_SETS[int(rng.integers(len(WALL_BC_SETS)))])
    name = ["d3dHsm", "box2d"][seed % 2]
    if name == "box2d":
        ov.pop("com.istabon", None)  # no rate tables loaded for the box
        if ov.get("bbb.iflcore") == 1:
            ov["bbb.pcoree"] = ov["bbb.pcorei"] = 2.5e4
    c, yl = make_case(name, perturb=float(rng.choice([1e-3, 2e-2])), seed=int(rng.integers(1 << 30)), overrides=ov)
    ora = bind(oracle(), c)
    bind(hk, c)
    n = c.bbb.neq
    for lib in (ora, hk):
        lib.set_real("nufak", 1.0e3)
    fo, fh = ora.pandf1(yl), hk.pandf1(yl)
    if not np.isfinite(fo).all():
        pytest.skip("non-physical combination")
    assert np.array_equal(fo, fh), "%s: %d residual entries differ" % (ov, (fo != fh).sum())
    y, su = psetnk_inputs(c, yl)
    dt = 10.0 ** rng.uniform(-6, -2, n)
    for lib in (ora, hk):
        lib.step_params(dt, y[:n], su, np.ones(n))
    f1, f2 = ora.pandf1(y), hk.pandf1(y)
    jo = ora.jac_calc(y, f1, c.bbb.lbw, c.bbb.ubw, c.bbb.nnzmx)
    jh = hk.jac_calc(y, f2, c.bbb.lbw, c.bbb.ubw, c.bbb.nnzmx)
    assert all(np.array_equal(p, q) for p, q in zip(jo, jh)), ov
    for lib in (ora, hk):
        lib.set_real("nufak", 0.0)


def test_kernel_logic_all_single_switch_variants(hk):
    from tests.test_gpu_parity import VARIANTS
    for variant in sorted(VARIANTS):
        c, yl = make_case("d3dHsm", perturb=2e-3, overrides=VARIANTS[variant])
        _check(c, yl, hk)


@pytest.mark.parametrize("name", ["case2", "d3dHsm2x"])
def test_kernel_logic_other_grids(hk, name):
    c, yl = make_case(name, perturb=1e-3)
    _check(c, yl, hk)
