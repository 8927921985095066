"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances are BASELINE.json's: residual 1e-12 relative (to the largest |yldot| of the same
equation type: converged states have yldot -> 0), Jacobian entries 1e-8 relative, CSR pattern
(ia/ja) bit-exact."""
import os

import numpy as np
import pytest

from tests.util import ROOT, bind, make_case, newton_solve, oracle, psetnk_inputs
from uedge_b200.capi import load_gpu

pytestmark = pytest.mark.gpu

RES_RTOL = 1e-12
JAC_RTOL = 1e-8


def _pair(name, perturb, istabon=0):
    c, yl = make_case(name, istabon=istabon, perturb=perturb)
    gpu = bind(load_gpu(), c)
    ora = bind(oracle(), c)
    return c, yl, gpu, ora


def _res_err(fg, fo):
    scale = np.abs(fo).reshape(-1, 5).max(axis=0)
    return (np.abs(fg - fo).reshape(-1, 5) / scale).max()


@pytest.mark.parametrize("name,perturb", [("d3dHsm", 1e-3), ("d3dHsm", 0.05), ("case2", 0.0), ("d3dHsm4x", 1e-3)])
def test_residual_parity(built, name, perturb):
    c, yl, gpu, ora = _pair(name, perturb)
    fg, fo = gpu.pandf1(yl), ora.pandf1(yl)
    assert np.isfinite(fg).all()
    assert _res_err(fg, fo) < RES_RTOL
    # both sides use include/ue_math.h and no FMA contraction: the residual is bit-identical
    assert np.array_equal(fg, fo), "%d of %d entries differ in the last bits" % ((fg != fo).sum(), fg.size)


def test_residual_converged_state_on_gpu(built):
    """The reference's converged state must also be a root of the CUDA residual."""
    c, yl, gpu, ora = _pair("d3dHsm", 0.0)
    f = gpu.pandf1(yl).reshape(-1, 5)
    assert np.abs(f).max() < 5e-6


@pytest.mark.parametrize("isbcwdt", [0, 1])
def test_residual_timestep_term(built, isbcwdt):
    """yl(neq+1)<0 and dtreal<1e15 adds -(yl-ylodt)/dtuse on interior rows, and on the guard rows too when isbcwdt=1
    (oderhs.m:7963-8037; pyexamples/d3dHsmNew starts with isbcwdt=1)."""
    c, yl = make_case("d3dHsm", perturb=1e-3, overrides={"bbb.isbcwdt": isbcwdt})
    gpu, ora = bind(load_gpu(), c), bind(oracle(), c)
    n = c.bbb.neq
    rng = np.random.default_rng(3)
    dt = 10.0 ** rng.uniform(-6, -3, n)
    yo = yl[:n] * (1 + 1e-2 * rng.uniform(-1, 1, n))
    for lib in (gpu, ora):
        lib.set_real("dtreal", 1e-4)
        lib.step_params(dt, yo, np.ones(n), np.ones(n))
    y = yl.copy(); y[n] = -1.0
    fg, fo = gpu.pandf1(y), ora.pandf1(y)
    assert np.array_equal(fg, fo)
    y[n] = 1.0
    assert not np.array_equal(gpu.pandf1(y), fg)  # without the flag the term is absent
    # Jacobian with the term's diagonal contribution (-1/dtuse on every row when isbcwdt=1)
    jg, jo, noise = _jac_pair(c, yl, gpu, ora, dt=dt)
    assert all(np.array_equal(p, q) for p, q in zip(jg, jo))
    for lib in (gpu, ora):
        lib.set_real("dtreal", 1e20)


@pytest.mark.gpu
def test_csc_copy_and_timing_accumulators(built):
    """rcsc/icsc/jcsc (oderhs.m:8620-8752) = the transpose of the CSR the call returned; ttotfe/ttotjf accumulate the time spent in
    the two entry points (com/com.v:500-519)."""
    import ctypes as C
    c, yl = make_case("d3dHsm", perturb=1e-3)
    gpu = bind(load_gpu(), c)
    b = c.bbb
    n = b.neq
    y, su = psetnk_inputs(c, yl)
    gpu.step_params(np.full(n, 1e20), y[:n], su, np.ones(n))
    lib = gpu.lib
    lib.ue_gpu_timing.argtypes = [C.POINTER(C.c_double)] * 3 + [C.c_int64]
    t = [C.c_double(0) for _ in range(3)]
    lib.ue_gpu_timing(*[C.byref(x) for x in t], 1)
    f = gpu.pandf1(y)
    jac, ja, ia = gpu.jac_calc(y, f, b.lbw, b.ubw, b.nnzmx)
    lib.ue_gpu_timing(*[C.byref(x) for x in t], 0)
    assert 0 < t[0].value < 1.0 and 0 < t[1].value < 1.0 and t[2].value == 0.0
    nnz = len(jac)
    rcsc = np.zeros(nnz); icsc = np.zeros(nnz, dtype=np.int64); jcsc = np.zeros(n + 1, dtype=np.int64); got = C.c_int64(0)
    lib.ue_gpu_get_csc.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
    assert lib.ue_gpu_get_csc(nnz, rcsc.ctypes.data, icsc.ctypes.data, jcsc.ctypes.data, C.byref(got)) == 0, lib.ue_gpu_last_error()
    assert got.value == nnz
    rows = np.repeat(np.arange(1, n + 1), np.diff(ia))          # CSR -> (row, col, val), sorted by (col, row)
    order = np.lexsort((rows, ja))
    assert np.array_equal(icsc, rows[order]) and np.array_equal(rcsc, jac[order])
    assert np.array_equal(np.diff(jcsc), np.bincount(ja, minlength=n + 1)[1:]) and jcsc[0] == 1 and jcsc[n] == nnz + 1


@pytest.mark.gpu
def test_vnormnk(built):
    """vnormnk(n, v, s) = sqrt(sum((v*s)**2)) (svr/nksol.m:1404-1419) against the serial sum, to 1e-14 relative (the device sum
    has a fixed tree shape); ue_gpu_fnrm = the same norm of the resident residual with the resident sfscal."""
    import ctypes as C
    c, yl = make_case("d3dHsm", perturb=1e-3)
    gpu = bind(load_gpu(), c)
    n = c.bbb.neq
    rng = np.random.default_rng(5)
    v = rng.standard_normal(n) * 10.0 ** rng.uniform(-8, 8, n); s = 10.0 ** rng.uniform(-3, 3, n)
    ref = 0.0
    for i in range(n):
        ref = ref + (v[i] * s[i]) ** 2
    ref = np.sqrt(ref)
    out = C.c_double(0)
    lib = gpu.lib
    lib.ue_gpu_vnormnk.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    assert lib.ue_gpu_vnormnk(n, v.ctypes.data, s.ctypes.data, C.byref(out)) == 0
    assert abs(out.value - ref) <= 1e-14 * ref
    a = out.value
    assert lib.ue_gpu_vnormnk(n, v.ctypes.data, s.ctypes.data, C.byref(out)) == 0 and out.value == a  # deterministic
    y, su = psetnk_inputs(c, yl)
    sf = 10.0 ** rng.uniform(-2, 2, n)
    gpu.step_params(np.full(n, 1e20), y[:n], su, sf)
    f = gpu.pandf1(y)
    lib.ue_gpu_fnrm.argtypes = [C.POINTER(C.c_double)]
    assert lib.ue_gpu_fnrm(C.byref(out)) == 0
    assert abs(out.value - np.sqrt(np.sum((f * sf) ** 2))) <= 1e-13 * out.value


@pytest.mark.gpu
def test_time_step_inputs_changed_between_jacobian_and_residual(built):
    """The host shim sends dtuse / ylodt / dtreal before EVERY residual: set_dt rewrites dtuse and dtreal may
    change (icntnunk=1) between a Jacobian and the next rhsnk.  A Jacobian at one set of values, then residuals at two others -
    every result must be the oracle's for the values in force at that call (cached graphs and unchanged-vector tests included)."""
    c, yl = make_case("d3dHsm", perturb=1e-3, overrides={"bbb.isbcwdt": 1})
    gpu, ora = bind(load_gpu(), c), bind(oracle(), c)
    b = c.bbb
    n = b.neq
    yj, su = psetnk_inputs(c, yl)  # yl(neq+1) = 1: psetnk's calls (rhsnk + jac_calc)
    y = yj.copy(); y[n] = -1.0      # yl(neq+1) < 0: nksol's own residual calls (oderhs.m:7961-7964)
    rng = np.random.default_rng(11)
    dt1 = 10.0 ** rng.uniform(-6, -3, n)
    for lib in (gpu, ora):
        lib.set_real("dtreal", 1e-4)
        lib.step_params(dt1, 0.9995 * y[:n], su, np.ones(n))
    fg, fo = gpu.pandf1(yj), ora.pandf1(yj)
    jg, jo = gpu.jac_calc(yj, fg, b.lbw, b.ubw, b.nnzmx), ora.jac_calc(yj, fo, b.lbw, b.ubw, b.nnzmx)
    assert np.array_equal(fg, fo) and all(np.array_equal(p, q) for p, q in zip(jg, jo))
    fg, fo = gpu.pandf1(y), ora.pandf1(y)
    assert np.array_equal(fg, fo)
    seen = [fg]
    for k, dtreal in enumerate((1e-5, 1e-4, 1e20)):  # new dtuse and ylodt (set_dt, next exmain) and a new dtreal each time
        dt2 = dt1 * (0.5 + k)
        yo = y[:n] * (1.0 - 1e-3 * (k + 1))
        for lib in (gpu, ora):
            lib.step_params(dt2, yo, su, np.ones(n))
            lib.set_real("dtreal", dtreal)
        fg, fo = gpu.pandf1(y), ora.pandf1(y)
        assert np.array_equal(fg, fo)
        assert all(not np.array_equal(fg, f) for f in seen)
        seen.append(fg)
    for lib in (gpu, ora):
        lib.set_real("dtreal", 1e20)


def _jac_pair(c, yl, gpu, ora, dt=None):
    b = c.bbb
    y, su = psetnk_inputs(c, yl)
    dtuse = np.full(b.neq, 1e20) if dt is None else dt
    for lib in (gpu, ora):
        lib.step_params(dtuse, y[: b.neq], su, np.ones(b.neq))
    fg, fo = gpu.pandf1(y), ora.pandf1(y)
    jg = gpu.jac_calc(y, fg, b.lbw, b.ubw, b.nnzmx)
    jo = ora.jac_calc(y, fo, b.lbw, b.ubw, b.nnzmx)
    # finite-difference noise floor: the two libm's (glibc / CUDA) differ by ~1 ulp in exp/log/pow, i.e.
    # delta f_i ~ ulps * eps * (size of the terms of equation i); an entry is delta f_i / dyl_j.
    rng = np.random.default_rng(99)
    y2 = y.copy()
    y2[: b.neq] *= 1 + 0.05 * rng.uniform(-1, 1, b.neq)
    fscale = np.abs(ora.pandf1(y2)).reshape(-1, b.numvar).max(axis=0)
    ora.pandf1(y)
    dyl = 1e-8 * (np.abs(y[: b.neq]) + 1.0 / su)
    return jg, jo, (fscale, dyl)


def _check_jac(jg, jo, noise):
    (vg, jag, iag), (vo, jao, iao) = jg, jo
    fscale, dyl = noise
    assert np.array_equal(iag, iao), "ia differs: nnz %d vs %d" % (len(vg), len(vo))
    assert np.array_equal(jag, jao), "ja differs"
    rows = np.repeat(np.arange(len(iao) - 1), np.diff(iao))
    floor = 64 * 2.2e-16 * fscale[rows % len(fscale)] / dyl[jao - 1]
    err = np.abs(vg - vo)
    bad = err > JAC_RTOL * np.abs(vo) + floor
    assert not bad.any(), "%d entries off; worst rel %g" % (bad.sum(), (err / np.abs(vo))[bad].max())
    # stronger: identical arithmetic on both sides (ue_math.h, no FMA) => the values are bit-identical,
    # which is what makes the value-dependent sparsity pattern reproducible at all
    assert np.array_equal(vg, vo), "%d of %d Jacobian values differ" % ((vg != vo).sum(), vg.size)


@pytest.mark.parametrize("name,perturb", [("d3dHsm", 1e-3), ("d3dHsm", 0.05), ("case2", 0.0)])
def test_jacobian_parity(built, name, perturb):
    c, yl, gpu, ora = _pair(name, perturb)
    _check_jac(*_jac_pair(c, yl, gpu, ora))


def test_jacobian_parity_with_dt(built):
    c, yl, gpu, ora = _pair("d3dHsm", 1e-2)
    rng = np.random.default_rng(5)
    _check_jac(*_jac_pair(c, yl, gpu, ora, dt=10.0 ** rng.uniform(-6, -2, c.bbb.neq)))


def test_jacobian_parity_4x(built):
    c, yl, gpu, ora = _pair("d3dHsm4x", 1e-3)
    _check_jac(*_jac_pair(c, yl, gpu, ora))


# switch-set variants: every branch the kernels claim to support, residual AND Jacobian bit-identical to the oracle
VARIANTS = {
    "central_differencing": {"bbb.methn": 22, "bbb.methu": 22, "bbb.methe": 22, "bbb.methi": 22, "bbb.methg": 22},
    "mixed_schemes": {"bbb.methn": 23, "bbb.methu": 32, "bbb.methe": 23, "bbb.methi": 32, "bbb.methg": 23},
    "core_power_flux_bc": {"bbb.iflcore": 1, "bbb.pcoree": 4.0e5, "bbb.pcorei": 4.0e5},
    "core_particle_flux_bc": {"bbb.isnicore": 0, "bbb.curcore": 10.0},
    "wall_flux_bcs": {"bbb.istewc": 0, "bbb.istiwc": 0, "bbb.istepfc": 1, "bbb.istipfc": 1, "bbb.isupcore": 1},
    "supersonic_plates": {"bbb.isupss": 1},
    "extrap_plates": {"bbb.isupss": -1},
    "albedo_like_recycling": {"bbb.recycp": -0.5},
    "flux_limits": {"bbb.isflxlde": 1, "bbb.flgam": 2.0, "bbb.isflxldi": 1, "bbb.flalfv": 0.5, "bbb.flalfgx": np.full(10, 1.0), "bbb.flalfgy": np.full(10, 1.0)},
    "ion_flux_limit_off": {"bbb.isflxldi": 0, "bbb.isplflxl": 1},
    "viscosity_options": {"bbb.isgxvon": 1, "bbb.ishavisy": 0, "bbb.isvhyha": 1},
    "log_radial_velocity": {"bbb.isvylog": 1, "bbb.difpr": 0.3},
    "cx_model_2_recomb": {"bbb.icnucx": 2, "bbb.isrecmon": 1, "com.istabon": 10},
    "const_rates": {"bbb.icnuiz": 1, "bbb.icnucx": 1, "bbb.islnlamcon": 1},
    "turbulent_kye": {"bbb.kyet": 0.5, "bbb.kyit": 0.5},
    "v81_defaults": {"bbb.oldseec": 0.0, "bbb.isoldalbarea": 0.0},
    "neutral_options": {"bbb.cngmom": 1.0, "bbb.cmwall": 0.5, "bbb.cngtgx": 1.0, "bbb.cngtgy": 1.0, "bbb.kxn": 1.0, "bbb.kyn": 1.0,
                        "bbb.isgasdc": 1, "bbb.cngflox": 1.0, "bbb.alftng": 0.5, "bbb.isdifxg_aug": 1, "bbb.isdifyg_aug": 1},
    "wide_jacobian_box": {"bbb.xlinc": 3, "bbb.xrinc": 2, "bbb.yinc": 3},
    "no_core_all": {"bbb.isjaccorall": 0},
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_switch_variants(built, variant):
    c, yl = make_case("d3dHsm", perturb=2e-3, overrides=VARIANTS[variant])
    gpu = bind(load_gpu(), c)
    ora = bind(oracle(), c)
    fg, fo = gpu.pandf1(yl), ora.pandf1(yl)
    assert np.isfinite(fo).all()
    assert np.array_equal(fg, fo), "%s: %d residual entries differ" % (variant, (fg != fo).sum())
    jg, jo, noise = _jac_pair(c, yl, gpu, ora)
    _check_jac(jg, jo, noise)


@pytest.mark.parametrize("seed", range(8))
def test_switch_combinations(built, seed):
    """Random combinations of 3-5 of the single-switch variants above (later ones win on a shared key), on a random state
    and with the time-step term and nufak on: the switches must also be right TOGETHER, bit for bit."""
    rng = np.random.default_rng(1000 + seed)
    names = list(rng.choice(sorted(VARIANTS), size=int(rng.integers(3, 6)), replace=False))
    ov = {}
    for nme in names:
        ov.update(VARIANTS[nme])
    c, yl = make_case("d3dHsm", perturb=float(rng.choice([1e-3, 5e-3, 2e-2])), seed=int(rng.integers(1 << 30)), overrides=ov)
    gpu = bind(load_gpu(), c)
    ora = bind(oracle(), c)
    n = c.bbb.neq
    for lib in (gpu, ora):
        lib.set_real("nufak", 1.0e3)
    fg, fo = gpu.pandf1(yl), ora.pandf1(yl)
    assert np.isfinite(fo).all(), names
    assert np.array_equal(fg, fo), "%s: %d residual entries differ" % (names, (fg != fo).sum())
    dt = 10.0 ** rng.uniform(-6, -2, n)
    jg, jo, noise = _jac_pair(c, yl, gpu, ora, dt=dt)
    assert np.array_equal(jg[2], jo[2]) and np.array_equal(jg[1], jo[1]), names
    assert np.array_equal(jg[0], jo[0]), names
    for lib in (gpu, ora):
        lib.set_real("nufak", 0.0)


_FUZZ_SKIP = {"ev", "qe", "me", "mp", "pi", "cutlo", "rt8opi", "difni2", "difpr2", "difax", "dif4order", "kye4order", "kyi4order", "l_parloss",
              "dtreal", "nufak", "delpert", "dylconst", "jaccliplim", "istgcon", "isoldalbarea", "sygytotc", "zi", "lgmax", "cne_sgvi",
              "lmfplim", "lxtimax", "lxtemax", "flalfipl", "flalfepl", "flalftf", "cfnus_i", "cfnus_e", "temp0", "n0", "n0g", "nnorm",
              "ennorm", "fnorm", "vpnorm", "tbmin", "recycm", "engbsr", "ebind", "cfnetap", "fracvgpgp", "fnnuiz"}
_FUZZ_ZERO_OK = ["vcony", "sigvi_floor", "fnuizx", "alfkxi", "alfkxe", "alfeqp", "nlimix", "nlimiy", "cmneut", "upcore", "tdiflim", "cngmom",
                 "cmwall", "cngtgx", "cngtgy", "kxn", "kyn", "alftng", "ccoldsor", "kyet", "kyit"]


@pytest.mark.parametrize("seed", range(12))
def test_coefficient_fuzz(built, seed):
    """Every real coefficient that crosses the ABI is an ordinary multiplier on both sides: scale ~60 of the non-zero ones
    by random factors and give a few zero-default ones small values; residual and Jacobian stay bit-identical.  Catches
    a term that one side evaluates and the other skips or hard-codes."""
    rng = np.random.default_rng(7000 + seed)
    c, yl = make_case("d3dHsm", perturb=2e-3, seed=50 + seed)
    s = c.static_inputs()
    names = [k for k, v in s["reals"].items() if k not in _FUZZ_SKIP and v != 0.0 and abs(v) < 1e15]
    for k in rng.choice(names, size=min(60, len(names)), replace=False):
        s["reals"][k] = float(s["reals"][k]) * float(rng.uniform(0.8, 1.25))
    for k in rng.choice(_FUZZ_ZERO_OK, size=6, replace=False):
        if s["reals"][k] == 0.0:
            s["reals"][k] = float(rng.uniform(0.05, 0.3))
    gpu, ora = load_gpu(), oracle()
    for lib in (gpu, ora):
        lib.load_static(s)
        lib.init()
    n = c.bbb.neq
    fg, fo = gpu.pandf1(yl), ora.pandf1(yl)
    assert np.isfinite(fo).all()
    assert np.array_equal(fg, fo), "%d residual entries differ" % (fg != fo).sum()
    y, su = psetnk_inputs(c, yl)
    for lib in (gpu, ora):
        lib.step_params(np.full(n, 1e20), y[:n], su, np.ones(n))
    f1, f2 = gpu.pandf1(y), ora.pandf1(y)
    jg = gpu.jac_calc(y, f1, c.bbb.lbw, c.bbb.ubw, c.bbb.nnzmx)
    jo = ora.jac_calc(y, f2, c.bbb.lbw, c.bbb.ubw, c.bbb.nnzmx)
    assert all(np.array_equal(p, q) for p, q in zip(jg, jo))


@pytest.mark.parametrize("seed", range(4))
def test_array_input_fuzz(built, seed):
    """Random per-element factors (+-10 %) on every geometry plane and on the 1-D coefficient arrays: physically
    inconsistent, but both sides must consume exactly the same numbers in the same way, and the dependency pruning of
    the Jacobian must not rely on any symmetry of the mesh."""
    rng = np.random.default_rng(9000 + seed)
    c, yl = make_case("d3dHsm", perturb=2e-3, seed=80 + seed)
    s = c.static_inputs()
    for k in s["planes"]:
        a = np.array(s["planes"][k], dtype=np.float64)
        s["planes"][k] = a * rng.uniform(0.9, 1.1, a.shape)
    for k in ("fgtdx", "fgtdy", "flalfea", "flalfia", "flalfva", "flalfgxa", "flalfgxya", "flalfgya", "tewalli", "tiwalli", "tewallo", "tiwallo",
              "alblb", "albrb", "albedoi", "albedoo"):
        a = np.array(s["lines"][k], dtype=np.float64)
        s["lines"][k] = np.minimum(a * rng.uniform(0.9, 1.1, a.shape), np.where(a <= 1.0, 1.0, np.inf))
    gpu, ora = load_gpu(), oracle()
    for lib in (gpu, ora):
        lib.load_static(s)
        lib.init()
    n = c.bbb.neq
    fg, fo = gpu.pandf1(yl), ora.pandf1(yl)
    assert np.isfinite(fo).all()
    assert np.array_equal(fg, fo), "%d residual entries differ" % (fg != fo).sum()
    y, su = psetnk_inputs(c, yl)
    for lib in (gpu, ora):
        lib.step_params(np.full(n, 1e20), y[:n], su, np.ones(n))
    f1, f2 = gpu.pandf1(y), ora.pandf1(y)
    jg = gpu.jac_calc(y, f1, c.bbb.lbw, c.bbb.ubw, c.bbb.nnzmx)
    jo = ora.jac_calc(y, f2, c.bbb.lbw, c.bbb.ubw, c.bbb.nnzmx)
    assert all(np.array_equal(p, q) for p, q in zip(jg, jo))


_INT_CHOICES = {
    "methn": [22, 23, 32, 33], "methu": [22, 23, 32, 33], "methe": [22, 23, 32, 33], "methi": [22, 23, 32, 33], "methg": [22, 23, 32, 33],
    "isjaccorall": [0, 1], "isbcwdt": [0, 1], "icnuiz": [0, 1, 2], "icnucx": [0, 1, 2], "isrecmon": [0, 1], "ingb": [1, 2, 3], "inflbg": [2, 4],
    "isgasdc": [0, 1], "isdifxg_aug": [0, 1], "isdifyg_aug": [0, 1], "isvylog": [0, 1], "isgxvon": [0, 1], "convis": [0, 1], "concap": [0, 1],
    "isflxlde": [0, 1], "isflxldi": [0, 1, 2], "isplflxl": [0, 1], "inkxc": [1, 2, 3], "ishavisy": [0, 1], "isvhyha": [0, 1],
    "islnlamcon": [0, 1], "iteb": [1, 2], "ifxnsgi": [0, 1], "isnicore": [0, 1], "isupcore": [0, 1], "iflcore": [0, 1], "ifluxni": [0, 1],
    "isupss": [-1, 0, 1], "isextrnp": [0, 1], "isbohmcalc": [0, 1], "newbcl": [0, 1], "newbcr": [0, 1], "xlinc": [2, 3], "xrinc": [1, 2],
    "yinc": [2, 3],
}


@pytest.mark.parametrize("seed", range(12))
def test_integer_switch_fuzz(built, seed):
    """Random values (from each switch's valid set) for a dozen integer switches at a time."""
    rng = np.random.default_rng(11000 + seed)
    c, yl = make_case("d3dHsm", perturb=2e-3, seed=120 + seed)
    s = c.static_inputs()
    picked = {}
    for k in rng.choice(sorted(_INT_CHOICES), size=12, replace=False):
        picked[str(k)] = int(rng.choice(_INT_CHOICES[str(k)]))
        s["ints"][str(k)] = picked[str(k)]
    if s["ints"]["iflcore"] == 1:
        s["reals"]["pcoree"] = s["reals"]["pcorei"] = 4.0e5
    gpu, ora = load_gpu(), oracle()
    try:
        for lib in (gpu, ora):
            lib.load_static(s)
            lib.init()
    except Exception as e:  # a combination both sides refuse is fine; they must refuse alike
        with pytest.raises(Exception):
            ora.load_static(s); ora.init()
        with pytest.raises(Exception):
            gpu.load_static(s); gpu.init()
        return
    n = c.bbb.neq
    fg, fo = gpu.pandf1(yl), ora.pandf1(yl)
    if not np.isfinite(fo).all():
        pytest.skip("non-physical combination: %s" % picked)
    assert np.array_equal(fg, fo), "%s: %d residual entries differ" % (picked, (fg != fo).sum())
    y, su = psetnk_inputs(c, yl)
    for lib in (gpu, ora):
        lib.step_params(np.full(n, 1e20), y[:n], su, np.ones(n))
    f1, f2 = gpu.pandf1(y), ora.pandf1(y)
    jg = gpu.jac_calc(y, f1, c.bbb.lbw, c.bbb.ubw, c.bbb.nnzmx)
    jo = ora.jac_calc(y, f2, c.bbb.lbw, c.bbb.ubw, c.bbb.nnzmx)
    assert all(np.array_equal(p, q) for p, q in zip(jg, jo)), picked


@pytest.mark.parametrize("name,seed", [("d3dHsm", 0), ("d3dHsm", 1), ("d3dHsm", 2), ("d3dHsm", 3), ("case2", 4), ("d3dHsm2x", 5)])
def test_rough_states(built, name, seed):
    """States far from any equilibrium: densities and temperatures scaled cell by cell by factors in [0.3, 3], gas by
    [0.1, 10], velocities by [-2, 2] (flow reversals).  Exercises the upwind switches, flux limiters and floors on both
    sides of every branch; residual and Jacobian stay bit-identical."""
    rng = np.random.default_rng(13000 + seed)
    c, yl = make_case(name)
    gpu, ora = bind(load_gpu(), c), bind(oracle(), c)
    n = c.bbb.neq
    y = yl.copy()
    Y = y[:n].reshape(-1, 5)
    Y[:, 0] *= np.exp(rng.uniform(np.log(0.3), np.log(3.0), len(Y)))
    Y[:, 1] *= rng.uniform(-2.0, 2.0, len(Y))
    Y[:, 2] *= np.exp(rng.uniform(np.log(0.3), np.log(3.0), len(Y)))
    Y[:, 3] *= np.exp(rng.uniform(np.log(0.3), np.log(3.0), len(Y)))
    Y[:, 4] *= np.exp(rng.uniform(np.log(0.1), np.log(10.0), len(Y)))
    fg, fo = gpu.pandf1(y), ora.pandf1(y)
    assert np.isfinite(fo).all()
    assert np.array_equal(fg, fo), "%d residual entries differ" % (fg != fo).sum()
    jg, jo, noise = _jac_pair(c, y, gpu, ora)
    assert all(np.array_equal(p, q) for p, q in zip(jg, jo))


def test_newton_on_gpu_recovers_reference_steady_state(built):
    """Newton driven entirely by the CUDA residual and Jacobian returns to the reference's converged state."""
    c, yref = make_case("d3dHsm")
    _, y0 = make_case("d3dHsm", perturb=1e-3, seed=7)
    gpu = bind(load_gpu(), c)
    y, hist = newton_solve(gpu, c, y0)
    assert hist[-1] < 1e-6 and hist[0] > 1.0
    n = c.bbb.neq
    assert np.abs((y[:n] - yref[:n]) * c.suscal(yref)).max() < 1e-8


def test_sfsetnk_on_device(built):
    """Row scaling chain of sfsetnk (oderhs.m:9815-9884) on the device vs the same chain on the oracle's CSR."""
    c, yl, gpu, ora = _pair("d3dHsm", 1e-3)
    b = c.bbb
    y, su = psetnk_inputs(c, yl)
    for lib in (gpu, ora):
        lib.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    sf_g, ymax_g = gpu.sfsetnk(yl, su, b.lbw, b.ubw)
    f0 = ora.pandf1(y)
    jac, ja, ia = ora.jac_calc(y, f0, b.lbw, b.ubw, b.nnzmx)
    rows = np.repeat(np.arange(b.neq), np.diff(ia))
    nrm = np.zeros(b.neq)
    np.maximum.at(nrm, rows, np.abs(jac * (1.0 / su)[ja - 1]))
    sf_o = 1.0 / nrm
    assert np.array_equal(sf_g, sf_o)
    assert ymax_g == max(np.abs(f0 * sf_o).max(), 1e-300)


def test_portable_math_bit_identical_on_device(built):
    """include/ue_math.h gives the same bits on the B200 and on the host over wide argument ranges."""
    import ctypes as C
    from tests.test_oracle_golden import _math_inputs
    glib, olib = load_gpu().lib, oracle().lib
    sig = [C.c_int64, C.c_int64] + [C.POINTER(C.c_double)] * 3
    glib.ue_gpu_math_probe.argtypes = sig; olib.ue_ora_math_probe.argtypes = sig
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for op, (x, y) in _math_inputs().items():
        yy = np.zeros_like(x) if y is None else y
        og, oo = np.zeros_like(x), np.zeros_like(x)
        assert glib.ue_gpu_math_probe(op, x.size, P(x), P(yy), P(og)) == 0
        assert olib.ue_ora_math_probe(op, x.size, P(x), P(yy), P(oo)) == 0
        assert np.array_equal(og, oo), (op, int((og != oo).sum()))


def test_page_locked_caller_arrays(built):
    """With page-locked caller arrays the kernels read yl and write yldot / jac / ja / ia directly (no copy nodes);
    the results are those of the copy path, for fresh pointers (un-captured) and repeated ones (graph replay)."""
    import torch
    c, yl, gpu, ora = _pair("d3dHsm", 1e-3)
    b = c.bbb; n = b.neq
    y, su = psetnk_inputs(c, yl)
    for lib in (gpu, ora):
        lib.step_params(np.full(n, 1e20), y[:n], su, np.ones(n))
    pin = lambda k, dt=torch.float64: torch.zeros(k, dtype=dt).pin_memory()
    hy, hf, hjac, hja, hia = pin(n + 2), pin(n + 2), pin(b.nnzmx), pin(b.nnzmx, torch.int64), pin(n + 1, torch.int64)
    for it in range(4):
        yk = y.copy(); yk[:n] *= 1 + 1e-6 * it
        hy.numpy()[:] = yk
        f = gpu.pandf1(hy.numpy(), out=hf.numpy())
        fo = ora.pandf1(yk)
        assert np.array_equal(f[:n], fo), "call %d" % it
        jg = gpu.jac_calc(hy.numpy(), hf.numpy(), b.lbw, b.ubw, b.nnzmx, out=(hjac.numpy(), hja.numpy(), hia.numpy()))
        jo = ora.jac_calc(yk, fo, b.lbw, b.ubw, b.nnzmx)
        assert all(np.array_equal(p, q) for p, q in zip(jg, jo)), "call %d" % it
    hy.numpy()[5 * 40] = -1.0
    with pytest.raises(Exception, match="ni is negative"):
        gpu.pandf1(hy.numpy(), out=hf.numpy())


def test_pin_host_array(built):
    """ue_gpu_pin_host_array on ordinary (pageable) arrays switches the same call to the direct path; same results."""
    import ctypes as C
    c, yl, gpu, ora = _pair("d3dHsm", 1e-3)
    n = c.bbb.neq
    ybuf, fbuf = yl.copy(), np.zeros(n)
    f_copy = gpu.pandf1(ybuf, out=fbuf).copy()
    lib = gpu.lib
    lib.ue_gpu_pin_host_array.argtypes = [C.c_void_p, C.c_int64]; lib.ue_gpu_unpin_host_array.argtypes = [C.c_void_p]
    assert lib.ue_gpu_pin_host_array(ybuf.ctypes.data, ybuf.nbytes) == 0
    assert lib.ue_gpu_pin_host_array(fbuf.ctypes.data, fbuf.nbytes) == 0
    assert lib.ue_gpu_pin_host_array(fbuf.ctypes.data, fbuf.nbytes) == 0  # idempotent
    for it in range(3):
        ybuf[:n] = yl[:n] * (1 + 1e-6 * it)
        fbuf[:] = 0.0
        gpu.pandf1(ybuf, out=fbuf)
        assert np.array_equal(fbuf, ora.pandf1(ybuf))
    assert lib.ue_gpu_unpin_host_array(ybuf.ctypes.data) == 0
    assert lib.ue_gpu_unpin_host_array(fbuf.ctypes.data) == 0
    ybuf[:n] = yl[:n]
    assert np.array_equal(gpu.pandf1(ybuf, out=fbuf), f_copy)


def test_fused_rhs_jac_equals_the_two_calls(built):
    c, yl, gpu, ora = _pair("d3dHsm", 1e-3)
    b = c.bbb
    y, su = psetnk_inputs(c, yl)
    gpu.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    for k in range(3):
        yk = y.copy(); yk[: b.neq] *= 1 + 1e-5 * k
        f1 = gpu.pandf1(yk)
        j1 = gpu.jac_calc(yk, f1, b.lbw, b.ubw, b.nnzmx)
        f2, j2 = gpu.rhs_jac(yk, b.lbw, b.ubw, b.nnzmx)
        assert np.array_equal(f1, f2) and all(np.array_equal(p, q) for p, q in zip(j1, j2))
    yk[5 * 40] = -1.0
    with pytest.raises(Exception, match="ni is negative"):
        gpu.rhs_jac(yk, b.lbw, b.ubw, b.nnzmx)


@pytest.mark.parametrize("normtype", [0, 1, 2])
def test_psetnk_scaling_chain_on_device(built, normtype):
    """amudia -> diamua -> roscal on the device-resident Jacobian (oderhs.m:9473-9485, svr/svrut4.m:954-1148) against
    the same chain in plain IEEE arithmetic on the oracle's CSR: bit-identical values and row factors."""
    c, yl, gpu, ora = _pair("d3dHsm", 1e-3)
    b = c.bbb
    y, su = psetnk_inputs(c, yl)
    rng = np.random.default_rng(11)
    sf = 10.0 ** rng.uniform(-3, 3, b.neq)
    for lib in (gpu, ora):
        lib.step_params(np.full(b.neq, 1e20), y[: b.neq], su, sf)
    f0 = gpu.pandf1(y)
    jg, jag, iag = gpu.jac_calc(y, f0, b.lbw, b.ubw, b.nnzmx)
    jo, jao, iao = ora.jac_calc(y, ora.pandf1(y), b.lbw, b.ubw, b.nnzmx)
    assert np.array_equal(jg, jo) and np.array_equal(jag, jao)
    scaled, fac = gpu.jac_scale(su, sf, len(jg), isrnorm=1, normtype=normtype)
    rows = np.repeat(np.arange(b.neq), np.diff(iao))
    a = (jo * (1.0 / su)[jao - 1]) * sf[rows]
    nrm = np.zeros(b.neq)
    for i in range(b.neq):  # serial sums in storage order, as rnrms does
        s = 0.0
        for v in a[iao[i] - 1 : iao[i + 1] - 1]:
            s = max(s, abs(v)) if normtype == 0 else (s + abs(v) if normtype == 1 else s + v * v)
        nrm[i] = np.sqrt(s) if normtype == 2 else s
    d = 1.0 / nrm
    assert np.array_equal(fac, d)
    assert np.array_equal(scaled, a * d[rows])
    with pytest.raises(Exception, match="nnz is not that of the last"):
        gpu.jac_scale(su, sf, len(jg) - 1)


@pytest.mark.parametrize("name,cuts", [("d3dHsm", [1, 301, 577, 901]), ("d3dHsm4x", [1, 5611, 11221])])
def test_jacobian_column_range_split(built, name, cuts):
    """ppp-style column split: the union of per-range CSRs equals the full CSR."""
    c, yl, gpu, ora = _pair(name, 1e-3)
    b = c.bbb
    y, su = psetnk_inputs(c, yl)
    gpu.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    f = gpu.pandf1(y)
    full = gpu.jac_calc(y, f, b.lbw, b.ubw, b.nnzmx)
    parts = []
    assert cuts[-1] == b.neq + 1
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        gpu.set_column_range(lo, hi - 1)
        parts.append(gpu.jac_calc(y, f, b.lbw, b.ubw, b.nnzmx))
    gpu.set_column_range(1, b.neq)
    vals, cols, ia = full
    assert sum(len(p[0]) for p in parts) == len(vals)
    for i in range(b.neq):
        cc = np.concatenate([p[1][p[2][i] - 1 : p[2][i + 1] - 1] for p in parts])
        vv = np.concatenate([p[0][p[2][i] - 1 : p[2][i + 1] - 1] for p in parts])
        assert np.array_equal(cc, cols[ia[i] - 1 : ia[i + 1] - 1])
        assert np.array_equal(vv, vals[ia[i] - 1 : ia[i + 1] - 1])


def test_jacobian_is_deterministic(built):
    c, yl, gpu, ora = _pair("d3dHsm", 1e-3)
    b = c.bbb
    y, su = psetnk_inputs(c, yl)
    gpu.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    f = gpu.pandf1(y)
    a = gpu.jac_calc(y, f, b.lbw, b.ubw, b.nnzmx)
    bb = gpu.jac_calc(y, f, b.lbw, b.ubw, b.nnzmx)
    assert all(np.array_equal(p, q) for p, q in zip(a, bb))


def test_coefficient_outside_the_built_path_is_refused_on_gpu(built):
    c, yl = make_case("d3dHsm", overrides={"bbb.cfyef": 1.0})
    gpu = load_gpu()
    with pytest.raises(Exception, match="cfyef must be 0"):  # at ue_gpu_init, or at once if the library is already initialised
        gpu.load_static(c.static_inputs())
        gpu.init()
    bind(gpu, make_case("d3dHsm")[0])  # and a clean set of inputs is accepted again


def test_assert_zero_guard(built):
    import ctypes as C
    lib = load_gpu().lib
    lib.ue_gpu_assert_zero.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int64]
    lib.ue_gpu_last_error.restype = C.c_char_p
    a = np.zeros(100)
    assert lib.ue_gpu_assert_zero(b"volpsor", a.ctypes.data_as(C.POINTER(C.c_double)), a.size) == 0
    a[37] = 1e-30
    assert lib.ue_gpu_assert_zero(b"volpsor", a.ctypes.data_as(C.POINTER(C.c_double)), a.size) == -5
    assert b"volpsor must be identically 0" in lib.ue_gpu_last_error()


def test_negative_density_is_trapped(built):
    c, yl, gpu, ora = _pair("d3dHsm", 0.0)
    y = yl.copy()
    y[5 * 40] = -1.0
    with pytest.raises(Exception, match="ni is negative"):
        gpu.pandf1(y)


def test_nnzmx_overflow_message(built):
    c, yl, gpu, ora = _pair("d3dHsm", 1e-3)
    b = c.bbb
    y, su = psetnk_inputs(c, yl)
    gpu.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    f = gpu.pandf1(y)
    with pytest.raises(Exception, match="More storage needed"):
        gpu.jac_calc(y, f, b.lbw, b.ubw, 1000)


def test_solver_style_buffer_reuse(built):
    """NKSOL calls rhsnk with the same work arrays every time: from the second call on the host entry point replays
    one graph holding upload, kernels and download.  Results must not depend on which path ran."""
    c, yl, gpu, ora = _pair("d3dHsm", 1e-3)
    n = c.bbb.neq
    rng = np.random.default_rng(5)
    ybuf, fbuf = yl.copy(), np.zeros(n)
    for it in range(5):
        ybuf[:n] = yl[:n] * (1 + 1e-3 * rng.uniform(-1, 1, n))
        gpu.pandf1(ybuf, out=fbuf)
        assert np.array_equal(fbuf, ora.pandf1(ybuf)), "call %d" % it
    # psetnk's second rhsnk: same unknowns, Jacobian flag off and the time-step term on (only the last phase reruns)
    dt = 10.0 ** rng.uniform(-6, -3, n)
    for lib in (gpu, ora):
        lib.set_real("dtreal", 1e-4)
        lib.step_params(dt, yl[:n], np.ones(n), np.ones(n))
    ybuf[n] = 1.0
    gpu.pandf1(ybuf, out=fbuf)
    assert np.array_equal(fbuf, ora.pandf1(ybuf))
    ybuf[n] = -1.0
    f_on = gpu.pandf1(ybuf, out=fbuf).copy()
    assert np.array_equal(f_on, ora.pandf1(ybuf))
    ybuf[n] = 1.0
    assert not np.array_equal(gpu.pandf1(ybuf), f_on)
    for lib in (gpu, ora):
        lib.set_real("dtreal", 1e20)
    ybuf[5 * 40] = -1.0
    with pytest.raises(Exception, match="ni is negative"):
        gpu.pandf1(ybuf, out=fbuf)


def test_jacobian_without_preceding_residual(built):
    """jac_calc for a yl other than the last residual call's: everything is uploaded, the residual is re-evaluated on
    the device and yldot00 is verified against it; the result is the same Jacobian."""
    c, yl, gpu, ora = _pair("d3dHsm", 1e-3)
    b = c.bbb
    y, su = psetnk_inputs(c, yl)
    for lib in (gpu, ora):
        lib.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    gpu.pandf1(y)
    y3 = y.copy()
    y3[: b.neq] *= 1 + 1e-4
    f3 = ora.pandf1(y3)
    jg = gpu.jac_calc(y3, f3, b.lbw, b.ubw, b.nnzmx)
    jo = ora.jac_calc(y3, f3, b.lbw, b.ubw, b.nnzmx)
    assert all(np.array_equal(p, q) for p, q in zip(jg, jo))
    # the base fields now describe y3: a second call takes the short path and must agree
    jg2 = gpu.jac_calc(y3, f3, b.lbw, b.ubw, b.nnzmx)
    assert all(np.array_equal(p, q) for p, q in zip(jg2, jo))


def test_foreign_yldot00_is_refused(built):
    """The dependency-pruned windows equal the reference's full windows only if yldot00 is pandf1(yl) bit for bit
    (rows outside the dependency set then difference to exactly zero).  Any other yldot00 fails loudly."""
    c, yl, gpu, ora = _pair("d3dHsm", 1e-3)
    b = c.bbb
    y, su = psetnk_inputs(c, yl)
    gpu.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    f = gpu.pandf1(y)
    f2 = f.copy()
    f2[7::11] *= 1 + 1e-9
    with pytest.raises(Exception, match="yldot00 is not pandf1"):
        gpu.jac_calc(y, f2, b.lbw, b.ubw, b.nnzmx)
    # changing the time-step vector between the two calls makes the cached residual stale: also verified
    gpu.pandf1(y)
    dt = np.full(b.neq, 1e-4)
    gpu.set_real("dtreal", 1e-4)
    gpu.step_params(dt, y[: b.neq] * 1.01, su, np.ones(b.neq))
    yneg = y.copy()
    yneg[b.neq] = -1.0
    fdt = gpu.pandf1(yneg)
    assert not np.array_equal(fdt, f)
    with pytest.raises(Exception, match="yldot00 is not pandf1"):
        gpu.jac_calc(yneg, f, b.lbw, b.ubw, b.nnzmx)
    gpu.set_real("dtreal", 1e20)


def test_large_grid_properties(built):
    """8x-refined grid (neq = 42 900, too large for the scalar oracle in a test): size-independent properties.
    (1) the column-split union equals the full Jacobian; (2) two evaluations are bit-identical; (3) J v equals the
    directional finite difference of the CUDA residual, (f(y + eps v) - f(y)) / eps, for random v."""
    import scipy.sparse as sp
    c, yl = make_case("d3dHsm8x", perturb=1e-3)
    gpu = bind(load_gpu(), c)
    b = c.bbb; n = b.neq
    y, su = psetnk_inputs(c, yl)
    y[n + 1] = 0.0
    gpu.set_real("nufak", 0.0)
    gpu.step_params(np.full(n, 1e20), y[:n], su, np.ones(n))
    f = gpu.pandf1(y)
    full = gpu.jac_calc(y, f, b.lbw, b.ubw, b.nnzmx)
    again = gpu.jac_calc(y, f, b.lbw, b.ubw, b.nnzmx)
    assert all(np.array_equal(p, q) for p, q in zip(full, again))
    parts = []
    cuts = [1, n // 3, (2 * n) // 3 + 7, n + 1]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        gpu.set_column_range(lo, hi - 1)
        parts.append(gpu.jac_calc(y, f, b.lbw, b.ubw, b.nnzmx))
    gpu.set_column_range(1, n)
    assert sum(len(p[0]) for p in parts) == len(full[0])
    mats = [sp.csr_matrix((p[0], p[1] - 1, p[2] - 1), shape=(n, n)) for p in parts]
    J = sp.csr_matrix((full[0], full[1] - 1, full[2] - 1), shape=(n, n))
    assert abs(sum(mats) - J).max() == 0.0
    # directional derivative; the -1/dtuse diagonal term (dtuse = 1e20) is far below the noise
    rng = np.random.default_rng(3)
    for _ in range(3):
        v = rng.uniform(-1, 1, n) / su            # a perturbation of relative size O(1) in every unknown
        eps = 1e-7
        y2 = y.copy(); y2[:n] += eps * v
        fd = (gpu.pandf1(y2) - f) / eps
        jv = J @ v
        scale = np.abs(J) @ np.abs(v) + 1e-300      # row-wise size of the terms
        assert (np.abs(fd - jv) / scale).max() < 2e-4


@pytest.mark.gpu
def test_nccl_split_jacobian(built):
    """ONE Jacobian assembled by 2 GPUs (ue_gpu_comm_init: column ranges + NCCL all-gather of the CSC fragments on the
    device): every rank returns the single-GPU CSR bit for bit (values, ja, ia), also against the CPU oracle.  Needs 2 GPUs."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, UE_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", os.path.join(ROOT, "tests", "nccl_worker.py"), "d3dHsm", "d3dHsm4x", "case1"],
                         env=env, capture_output=True, text=True, timeout=900)
    assert "NCCL_SPLIT_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
