"""Named configurations of the general hydrogen family (inputs of oracle/ue_oracle2.cpp)."""
import ctypes as C
import os

import numpy as np

from .case2 import Case2
from .cases import GOLDEN, load_grid_npz

SUBSETS = ("default", "ni", "ni-0", "ni-1", "up", "up-0", "up-1", "te", "ti", "phi")


def inputex_case(subset="default", mods=None):
    """pyexamples/input_example/input.py:19-110 on its 8x4 mesh; `subset` = which equations are on, as the groups
    pytests/<subset> of its solution.h5 record them (isnion/isupon/isteon/istion/isphion)."""
    g = load_grid_npz(os.path.join(GOLDEN, "inputex_8x4_grid.npz"))
    c = Case2(g)
    b, com = c.bbb, c.com
    com.nxleg = np.array([[2, 2]]); com.nxcore = np.array([[2, 2]]); com.nysol = np.array([3]); com.nycore = np.array([1])
    b.oldseec = 0.0; b.jhswitch = 0
    com.isnonog = 1
    b.methn = b.methu = b.methe = b.methi = 33; b.methg = 66
    b.isupwo[1] = 0; b.ineudif = 2; com.ngsp = 1; com.nhsp = 2; b.ziin[1] = 0; b.travis[1] = 0.0
    b.isnicore[0] = 1; b.ncore[0] = 2.0e19
    b.iflcore = 0; b.tcoree = 100.0; b.tcorei = 100.0
    b.recycp[0] = 0.9
    b.istewc = 1; b.tedge = 2.0; b.istepfc = 3; b.lyte = np.full_like(np.asarray(b.lyte, dtype=float), 0.03)
    b.matwso[0] = 1; b.isnwcono = np.ones_like(np.asarray(b.isnwcono)); b.isnwconi = np.ones_like(np.asarray(b.isnwconi))
    b.nwallo = 1.0e18; b.nwalli = 1.0e18
    b.recycw[0] = 0.9
    b.isngon = np.zeros_like(b.isngon); b.isupgon = np.zeros_like(b.isupgon); b.isupgon[0] = 1
    b.flalfe = 0.21; b.flalfi = 0.21; b.flalfv = 1.0
    b.flalfgx = np.full(10, 1.0); b.flalfgy = np.full(10, 1.0); b.flalfvgx = 1.0; b.flalfvgy = 1.0; b.flalftgx = 1.0; b.flalftgy = 1.0
    b.difni[1] = 1.0; b.kye = 1.0; b.kyi = 1.0; b.travis[1] = 1.0
    z = np.load(os.path.join(GOLDEN, "inputex_solution.npz"))
    p = "pytests__%s__" % subset
    b.isnion = z[p + "isnion"].astype(np.int64).copy(); b.isupon = z[p + "isupon"].astype(np.int64).copy()
    b.isteon = int(z[p + "isteon"]); b.istion = int(z[p + "istion"]); b.isphion = int(z[p + "isphion"]); b.isphiofft = 0
    if mods is not None:
        mods(b, com)
    c.setup()
    T = lambda a: np.ascontiguousarray(a.T)
    nis, ups = z["bbb__nis"], z["bbb__ups"]
    yl = c.set_state2([T(nis[:, :, 0]), T(nis[:, :, 1])], [T(ups[:, :, 0]), T(ups[:, :, 1])], T(z["bbb__tes"]), T(z["bbb__tis"]),
                      ng=T(z["bbb__ngs"][:, :, 0]), phi=T(z["bbb__phis"]), tg=T(z["bbb__tgs"][:, :, 0]))
    gold = {k[len(p):]: z[k] for k in z.files if k.startswith(p)}
    return c, yl, gold


def box2_inertial_case():
    """pyexamples/box2 exactly as its deck runs it (box2_in.py:12-140): the slab of box2_case with INERTIAL atoms -
    isupgon(1)=1, isngon=0, nhsp=2, ziin=(1,0), cngmom=cmwall=cngtgx=cngtgy=kxn=kyn=0 (box2_in.py:114-131)."""
    from .cases import box2_case
    c = box2_case(0, cls=Case2)
    b, com = c.bbb, c.com
    b.isupgon = np.zeros_like(b.isupgon); b.isupgon[0] = 1
    b.isngon = np.zeros_like(b.isngon)
    com.ngsp = 1; com.nhsp = 2
    b.ziin[0] = 1; b.ziin[1] = 0
    b.cngmom = np.zeros_like(np.asarray(b.cngmom, dtype=float)); b.cmwall = np.zeros_like(np.asarray(b.cmwall, dtype=float))
    b.cngtgx = np.zeros_like(np.asarray(b.cngtgx, dtype=float)); b.cngtgy = np.zeros_like(np.asarray(b.cngtgy, dtype=float))
    b.kxn = 0.0; b.kyn = 0.0
    b.isnion = np.asarray(b.isnion).copy(); b.isupon = np.asarray(b.isupon).copy()
    b.isnion[:2] = 1; b.isupon[:2] = 1
    c.setup()
    return c


def box2_initial_state(c):
    """restart=0 profiles of ueinit for a half-space slab (bbb/odesetup.m:1355-1452): densities nibeg(ifld)*proffacy, both
    parallel velocities sqrt(te(0,0)/mi(1))*proffacx*proffacy with up(nx+1)=up(nx), te = ttbeg*proffacx*proffacy, ti = tscal*te."""
    b, com = c.bbb, c.com
    nx, ny = com.nx, com.ny
    IY, IX = np.meshgrid(np.arange(ny + 2), np.arange(nx + 2), indexing="ij")
    px = (nx + 3 - IX) / float(nx + 3)
    py = (ny + 3 - IY) / float(ny + 3)
    ttbeg = float(b.tinit) * b.ev if "tinit" in b else 40.0 * b.ev
    te = ttbeg * px * py
    ti = float(b.tscal) * ttbeg * px * py
    ni = [float(b.nibeg[f]) * py for f in range(2)]
    up = np.sqrt(te[0, 0] / b.mi[0]) * px * py
    up[:, nx + 1] = up[:, nx]
    return c.set_state2(ni, [up, up.copy()], te, ti, tg=np.full_like(te, float(b.tscal) * ttbeg))


def d3d_full_physics_case(grid=None):
    """The DIII-D single-null 16x8 mesh of d3dHsm (or a refinement) with everything the general path has switched on: non-orthogonal
    stencils, inertial atoms (nhsp=2), log-interpolated gas flux (methg=66) and the potential equation, 7 unknowns per cell - the
    input_example switch set on the headline mesh.  State: the d3dHsm restart with ni(,,2) := ng, the atoms moving with 30 % of the
    ion velocity and phi = 3 Te/e.  No reference vector exists for this combination: oracle <-> CUDA parity and timing only."""
    from .cases import d3dhsm_case, load_grid_npz, load_state_npz, refine_state
    g = grid or load_grid_npz()
    c = d3dhsm_case(g, cls=Case2)
    b, com = c.bbb, c.com
    b.oldseec = 0.0; b.isoldalbarea = 0.0
    com.isnonog = 1
    b.methg = 66
    b.isupwo[1] = 0; b.ineudif = 2; com.ngsp = 1; com.nhsp = 2; b.ziin[1] = 0
    b.isngon = np.zeros_like(b.isngon); b.isupgon = np.zeros_like(b.isupgon); b.isupgon[0] = 1
    b.isnion = np.asarray(b.isnion).copy(); b.isupon = np.asarray(b.isupon).copy(); b.isnion[:2] = 1; b.isupon[:2] = 1
    b.flalfgx = np.full(10, 1.0); b.flalfgy = np.full(10, 1.0); b.flalfvgx = 1.0; b.flalfvgy = 1.0; b.flalftgx = 1.0; b.flalftgy = 1.0
    b.travis[1] = 0.0; b.difni[1] = 1.0
    b.isphion = 1; b.isphiofft = 0
    c.setup()
    ni, up, te, ti, ng = load_state_npz("d3dHsm_state.npz")
    fx, fy = g["nxm"] // 16, g["nym"] // 8
    if fx > 1 or fy > 1:
        ni, up, te, ti, ng = refine_state((ni, up, te, ti, ng), fx, fy)
    yl = c.set_state2([ni, ng], [up, 0.3 * up], te, ti, ng=ng, phi=3.0 * te / b.ev, tg=ti)
    return c, yl


def jupyter_case(mods=None, v8_0_defaults=True, grid=None):
    """jupyter/case_setup.py (the case of jupyter/PyUedge.ipynb; BASELINE configs[2]): the 16x8 DIII-D mesh, inertial atoms, potential
    equation with isnewpot=1, ExB and grad-B drifts (cfyef=cf2ef=cfybf=cf2bf=1), grad-B currents, Joule heating, sheath conditions
    from the current (newbcl=newbcr=1, isfdiax=1), iphibcc=3.  State: jupyter/d3d.hdf5 (tests/golden/jupyter_d3d_state.npz).  The
    notebook regenerates its mesh from aeqdsk/neqdsk (mesh generation is out of scope); the mesh here is builder/test/facets/gridue, built
    from the same equilibrium with the same flux-surface and poloidal distribution inputs.  `grid`: a refinement of that mesh (the
    state is then prolonged piecewise constant - timing and oracle <-> CUDA parity only)."""
    from .cases import refine_state
    g = grid or load_grid_npz()
    fx, fy = g["nxm"] // 16, g["nym"] // 8
    c = Case2(g)
    b, com = c.bbb, c.com
    com.nxleg = np.array([[4 * fx, 4 * fx]]); com.nxcore = np.array([[4 * fx, 4 * fx]]); com.nysol = np.array([6 * fy]); com.nycore = np.array([2 * fy])
    if v8_0_defaults:
        b.oldseec = 1.0; b.isoldalbarea = 1.0
    b.methn = b.methu = b.methe = b.methi = b.methg = 33
    b.ncore[0] = 2.0e19; b.iflcore = 0; b.tcoree = 100.0; b.tcorei = 100.0; b.tedge = 2.0
    b.istepfc = 3; b.lyte = np.full_like(np.asarray(b.lyte, dtype=float), 0.03)
    b.recycp[0] = 0.98; b.recycw[0] = 0.9; b.matwso[0] = 1
    b.isnwcono = np.ones_like(np.asarray(b.isnwcono)); b.isnwconi = np.ones_like(np.asarray(b.isnwconi))
    b.nwallo = 1.0e18; b.nwalli = 1.0e18
    b.difni[0] = 1.0; b.kye = 1.0; b.kyi = 1.0; b.travis[0] = 1.0
    b.flalfe = 0.21; b.flalfi = 0.21; b.flalfv = 1.0
    b.flalfgx = np.full(10, 1.0); b.flalfgy = np.full(10, 1.0); b.flalfvgx = 1.0; b.flalfvgy = 1.0; b.flalftgx = 1.0; b.flalftgy = 1.0
    b.ineudif = 2
    b.isupgon = np.zeros_like(b.isupgon); b.isupgon[0] = 1; b.isngon = np.zeros_like(b.isngon)
    com.ngsp = 1; com.nhsp = 2; b.ziin[1] = 0; b.travis[1] = 0.0
    b.cngmom = np.zeros_like(np.asarray(b.cngmom, dtype=float)); b.cmwall = np.zeros_like(np.asarray(b.cmwall, dtype=float))
    b.cngtgx = np.zeros_like(np.asarray(b.cngtgx, dtype=float)); b.cngtgy = np.zeros_like(np.asarray(b.cngtgy, dtype=float))
    b.kxn = 0.0; b.kyn = 0.0
    b.isnion = np.asarray(b.isnion).copy(); b.isupon = np.asarray(b.isupon).copy(); b.isnion[:2] = 1; b.isupon[:2] = 1
    b.isphion = 1; b.isphiofft = 0; b.b0 = 1.0; b.rsigpl = 1.0e-8; b.cfjhf = 1.0; b.cfjve = 1.0; b.jhswitch = 1; b.cfjpy = 0.0; b.cfjp2 = 0.0
    b.newbcl = np.ones_like(np.asarray(b.newbcl)); b.newbcr = np.ones_like(np.asarray(b.newbcr)); b.isfdiax = 1.0
    b.cfyef = 1.0; b.cf2ef = 1.0; b.cfydd = 0.0; b.cf2dd = 0.0; b.cfrd = 0.0; b.cfbgt = 0.0; b.cfybf = 1.0; b.cf2bf = 1.0; b.cfqybf = 1.0; b.cfq2bf = 1.0
    b.isnewpot = 1; b.rnewpot = 1.0; b.iphibcc = 3
    if mods is not None:
        mods(b, com)
    c.setup()
    z = np.load(os.path.join(GOLDEN, "jupyter_d3d_state.npz"))
    T = lambda a: np.ascontiguousarray(a.T)
    nis, ups = z["nis"], z["ups"]
    pl = [T(nis[:, :, 0]), T(nis[:, :, 1]), T(ups[:, :, 0]), T(ups[:, :, 1]), T(z["tes"]), T(z["tis"]), T(z["ngs"][:, :, 0]), T(z["phis"])]
    if fx > 1 or fy > 1:
        pl = refine_state(pl, fx, fy)
    yl = c.set_state2([pl[0], pl[1]], [pl[2], pl[3]], pl[4], pl[5], ng=pl[6], phi=pl[7], tg=pl[5])
    return c, yl


def all_drifts(b, com):
    """mods for jupyter_case: the diamagnetic and resistive parts of the drifts and currents on top of the deck's ExB / grad-B set
    (the decks keep them at zero: `cfydd`, `cf2dd` 'always = 0' in jupyter/case_setup.py:99-103)"""
    b.cfydd = 1.0; b.cf2dd = 1.0; b.cfrd = 1.0; b.cfbgt = 1.0; b.cfjpy = 1.0; b.cfjp2 = 1.0
    # the classical (Braginskii) momentum-transfer velocity and conductivities, the charge-exchange / neoclassical current
    b.cfvycr = 1.0; b.cfrtaue = 1.0; b.cfeta1 = 1.0; b.cfcl_e = 1.0; b.cfcl_i = 1.0; b.cfqyn = 1.0; b.nuneo = 1.0e3


def braginskii_current(b, com):
    """mods for jupyter_case: the radial current from the classical viscosity alone (cfvycf: fqy := e n vycf, potencur.m:405-408, and
    the core potential conditions of boundary.m:1095-1100)"""
    b.cfvycf = 1.0; b.cfeta1 = 1.0


def gas_energy_case(mods=None, grid=None, deck="jupyter"):
    """istgon = 1 (the gas energy equation engbalg, bbb/oderhs.m:7508-7878) on top of a deck with inertial atoms: `jupyter` = the drift
    case on the DIII-D mesh (numvar 8), `inputex` = pyexamples/input_example on its non-orthogonal mesh (the fegxy term).  As the
    reference's comments prescribe (bbb.v:345): cftiexclg = 0 (the atoms leave the ion energy equation), istgcon = -1 (tg is an unknown,
    not a multiple of ti).  State: the deck's restart with tg := 0.8 ti.  No reference vector exists for istgon = 1."""
    def on(b, com):
        b.istgon = np.asarray(b.istgon).copy(); b.istgon[0] = 1
        b.istgcon = np.asarray(b.istgcon, dtype=float).copy(); b.istgcon[0] = -1.0
        b.cftiexclg = 0.0
        if mods is not None:
            mods(b, com)
    if deck == "jupyter":
        c, yl = jupyter_case(on, grid=grid)
    else:
        c, yl, _ = inputex_case("default", mods=on)
    st = c.st
    yl = c.set_state2(st["ni"], st["up"], st["te"], st["ti"], ng=st["ng"], phi=st["phi"], tg=0.8 * st["ti"])
    return c, yl


def switch_variant(seed):
    """A random combination of the switches and coefficients the general path implements beyond the input_example deck (differencing
    schemes 0-8, flux limits, viscosity and conductivity options, rate models, boundary options, 4th-order terms): returns a function
    (b, com) -> None for inputex_case(mods=...), and a description."""
    rng = np.random.default_rng(seed)
    pick = lambda *a: a[int(rng.integers(len(a)))]
    ch = {}
    if rng.random() < 0.8:
        m = pick(0, 1, 2, 3, 4, 5, 6, 7) + 10 * pick(0, 1, 2, 3, 4, 5, 6, 7)
        ch["methn"] = pick(22, 33, 66, 23, 36, 62, 77); ch["methu"] = m; ch["methe"] = pick(m, 33, 22, 45, 54); ch["methi"] = pick(m, 33, 44, 55)
        ch["methg"] = pick(66, 33, 22, 77, 63, 26)
    for k, vals in (("isgxvon", (0, 1)), ("ishavisy", (1, 0)), ("isvhyha", (0, 1)), ("isvylog", (0, 1)), ("isintlog", (0, 1)), ("isflxlde", (0, 1)), ("isflxldi", (2, 0, 1)),
                    ("convis", (0, 1)), ("concap", (0, 1)), ("isgpye", (0, 1)), ("isnupdot1sd", (0, 1)), ("islnlamcon", (0, 1)), ("icnuiz", (0, 1)), ("icnucx", (0, 1, 2)),
                    ("isrecmon", (0, 1)), ("isplflxl", (0, 1)), ("isgasdc", (0, 1)), ("isdifxg_aug", (0, 1)), ("isdifyg_aug", (0, 1)), ("ifluxni", (1, 0)),
                    ("isrefluxclip", (1, 0)), ("iflcore", (0, 1, -1)), ("isexunif", (0, 1)), ("isugfm1side", (0, 1)), ("oldseec", (0.0, 1.0)), ("isoldalbarea", (0.0, 1.0)),
                    ("kye4order", (0.0, 1e-3)), ("kyi4order", (0.0, 2e-3)), ("isbcwdt", (0, 1)), ("inkxc", (0, 1, 2)), ("ingb", (2, 0, 1)), ("inflbg", (4, 2))):
        if rng.random() < 0.35:
            ch[k] = pick(*vals)
    idx = {}
    for k, i, vals in (("isupss", 0, (0, 1, -1)), ("isnicore", 0, (1, 0)), ("isupcore", 0, (0, 1, 2, 3)), ("isupcore", 1, (0, 1)), ("dif4order", 0, (0.0, 1e-3)),
                       ("newbcl", 0, (0, 1)), ("newbcr", 0, (0, 1)), ("difpr", 0, (0.0, 0.3)), ("difni2", 0, (0.0, 0.2)), ("difpr2", 0, (0.0, 0.1)), ("difax", 0, (0.0, 0.5)),
                       ("nlimix", 0, (0.0, 0.1)), ("nlimiy", 0, (0.0, 1e17)), ("vcony", 0, (0.0, 2.0)), ("cfvisxy", 0, (1.0, 0.0)), ("cfvisxy", 1, (1.0, 0.5))):
        if rng.random() < 0.3:
            idx[(k, i)] = pick(*vals)
    istab = pick(0, 0, 7) if rng.random() < 0.4 else None
    # cross-field drifts, grad-B currents, the new potential model with its core conditions, Joule heating (jupyter/case_setup.py:87-110)
    for k, vals in (("cfyef", (1.0, 0.5)), ("cf2ef", (1.0, 0.5)), ("cfybf", (1.0,)), ("cf2bf", (1.0,)), ("cfqybf", (1.0,)), ("cfq2bf", (1.0,)), ("isnewpot", (1,)), ("rnewpot", (1.0, 0.5)),
                    ("jhswitch", (1, 2)), ("isfdiax", (1.0,)), ("iphibcc", (1, 2, 3)), ("cfcurv", (0.5,)), ("cfgradb", (0.5,)), ("eycore", (10.0,)), ("icoreelec", (5.0,)),
                    ("cfqybbo", (1.0,)), ("cfqydbo", (1.0,)), ("cfniybbo", (1.0,)), ("cfeeybbo", (1.0,)), ("ExtendedJacPhi", (0,)),
                    ("cfydd", (1.0,)), ("cf2dd", (1.0,)), ("cfrd", (1.0, 0.5)), ("cfbgt", (1.0,)), ("cfjpy", (1.0,)), ("cfjp2", (1.0,)),
                    ("isybdrywd", (1,)), ("isphilbc", (1,)), ("isphirbc", (1,)), ("isfqpave", (1,)), ("cfniydbo", (1.0,)), ("cfeeydbo", (1.0,)),
                    ("cfvycr", (1.0,)), ("cfrtaue", (1.0,)), ("cfvycf", (1.0,)), ("cfeta1", (1.0,)), ("cfcl_e", (1.0,)), ("cfcl_i", (1.0,)), ("cfqyn", (1.0,)), ("nuneo", (1e3,))):
        if rng.random() < 0.4:
            ch[k] = pick(*vals)

    # gas energy equation (engbalg) with its boundary options
    tgopt = None
    if rng.random() < 0.35:
        tgopt = dict(istgpfc=pick(0, 1, 2, 3, 4, 5), istgwc=pick(0, 1, 2, 3, 4, 5), istgcore=pick(0, 1, 2, 3), istglb=pick(0, 1, 3, 4, 5), istgrb=pick(0, 1, 3, 4, 5),
                     cftiexclg=pick(0.0, 0.0, 1.0), recyce=pick(0.0, 0.3), recycwe=pick(0.0, 0.2), isfegxyqflave=pick(0, 1), lytg=pick(1e20, 0.05))

    def mods(b, com):
        for k, v in ch.items():
            setattr(b, k, v)
        if tgopt is not None:
            b.istgon = np.asarray(b.istgon).copy(); b.istgon[0] = 1
            b.istgcon = np.asarray(b.istgcon, dtype=float).copy(); b.istgcon[0] = -1.0
            for k in ("istgpfc", "istgwc", "istgcore"):
                a = np.asarray(getattr(b, k)).copy(); a[0] = tgopt[k]; setattr(b, k, a)
            b.istglb = tgopt["istglb"]; b.istgrb = tgopt["istgrb"]; b.cftiexclg = tgopt["cftiexclg"]; b.recyce = tgopt["recyce"]; b.recycwe = tgopt["recycwe"]
            b.isfegxyqflave = tgopt["isfegxyqflave"]; b.lytg = np.full(12, tgopt["lytg"])
        for (k, i), v in idx.items():
            a = np.asarray(getattr(b, k)).copy()
            a[i] = v
            setattr(b, k, a)
        if istab is not None:
            com.istabon = istab

    return mods, dict(ch, **{"%s[%d]" % k: v for k, v in idx.items()}, istabon=istab, **({"istgon": 1, **tgopt} if tgopt else {}))


class Lib2:
    """ctypes binding of one library that exports the generic-setter API <prefix>clear/set/init/step_params/pandf1/jac_calc/
    get_plane/last_error: the product's general path (libuegpu.so, prefix ue_gen_, include/ue_gen.h), its host build for the
    CPU logic check (tests/hostcheck, ue_genh_) or the oracle (oracle/libue_oracle2.so, ue_or2_)."""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self._f("last_error").restype = C.c_char_p
        self._f("set").argtypes = [C.c_char_p, C.c_void_p, C.c_int64]

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, self._f("last_error")().decode()))

    def bind(self, c):
        self.c = c
        self._f("clear")()
        self.keep = c.inputs2()
        for k, v in self.keep.items():
            self._f("set")(k.encode(), v.ctypes.data_as(C.c_void_p), v.size)
        self._ck(self._f("init")(), "init")
        self.neq = int(c.bbb.neq); self.NC = (c.com.nx + 2) * (c.com.ny + 2)
        return self

    def step_params(self, dtuse, ylodt, su, sf):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (dtuse, ylodt, su, sf)]
        self._f("step_params").argtypes = [C.c_int64] + [C.c_void_p] * 4
        self._ck(self._f("step_params")(self.neq, *[x.ctypes.data_as(C.c_void_p) for x in a]), "step_params")

    def pandf1(self, yl, xc=-1, yc=-1, out=None):
        yl = np.ascontiguousarray(yl, dtype=np.float64)
        yd = np.zeros(self.neq) if out is None else out
        if xc < 0 and yc < 0:
            self._f("pandf1").argtypes = [C.c_int64, C.c_double, C.c_void_p, C.c_void_p]
            self._ck(self._f("pandf1")(self.neq, 0.0, yl.ctypes.data_as(C.c_void_p), yd.ctypes.data_as(C.c_void_p)), "pandf1")
        else:  # (windowed evaluation: oracle only)
            self._f("pandf1_win").argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
            self._ck(self._f("pandf1_win")(xc, yc, self.neq, yl.ctypes.data_as(C.c_void_p), yd.ctypes.data_as(C.c_void_p)), "pandf1")
        return yd

    def jac_calc(self, yl, f0, ml, mu, nnzmx):
        yl = np.ascontiguousarray(yl, dtype=np.float64)
        y0 = np.zeros(self.neq + 2); y0[: self.neq] = f0[: self.neq]
        jac = np.zeros(nnzmx); ja = np.zeros(nnzmx, dtype=np.int64); ia = np.zeros(self.neq + 1, dtype=np.int64); nnz = C.c_int64(0)
        self._f("jac_calc").argtypes = [C.c_int64, C.c_double] + [C.c_void_p] * 2 + [C.c_int64] * 3 + [C.c_void_p] * 3 + [C.POINTER(C.c_int64)]
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        self._ck(self._f("jac_calc")(self.neq, 0.0, P(yl), P(y0), int(ml), int(mu), int(nnzmx), P(jac), P(ja), P(ia), C.byref(nnz)), "jac_calc")
        n = nnz.value
        return jac[:n].copy(), ja[:n].copy(), ia

    def jac_calc_raw(self, yl, y0, ml, mu, nnzmx, bufs=None):
        """The same C-ABI call into caller-owned arrays (what a host code does): returns nnz and the buffers (jac, ja, ia)."""
        if bufs is None:
            bufs = (np.zeros(nnzmx), np.zeros(nnzmx, dtype=np.int64), np.zeros(self.neq + 1, dtype=np.int64))
        jac, ja, ia = bufs
        nnz = C.c_int64(0)
        self._f("jac_calc").argtypes = [C.c_int64, C.c_double] + [C.c_void_p] * 2 + [C.c_int64] * 3 + [C.c_void_p] * 3 + [C.POINTER(C.c_int64)]
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        self._ck(self._f("jac_calc")(self.neq, 0.0, P(yl), P(y0), int(ml), int(mu), int(nnzmx), P(jac), P(ja), P(ia), C.byref(nnz)), "jac_calc")
        return nnz.value, bufs

    def plane(self, name):
        out = np.zeros(self.NC)
        self._f("get_plane").argtypes = [C.c_char_p, C.c_void_p]
        self._ck(self._f("get_plane")(name.encode(), out.ctypes.data_as(C.c_void_p)), "get_plane " + name)
        return out.reshape(self.c.com.ny + 2, self.c.com.nx + 2)


_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Oracle2(Lib2):
    """oracle/libue_oracle2.so (checker only)."""

    def __init__(self, path=None):
        super().__init__(path or os.path.join(_HERE, "oracle", "libue_oracle2.so"), "ue_or2_")


def load_gen():
    """The product's general path: libuegpu.so, entry points ue_gen_* (include/ue_gen.h).  Fails loudly if the library is missing;
    ue_gen_init fails without a CUDA device (there is no CPU fallback)."""
    path = os.path.join(_HERE, "uedge_b200", "csrc", "libuegpu.so")
    if not os.path.exists(path):
        raise RuntimeError("libuegpu.so is not built: run __graft_entry__.build()")
    return Lib2(path, "ue_gen_")
