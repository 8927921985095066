"""Named configurations of BASELINE.json, built on `Case`.

d3dHsm: switch set of pyexamples/d3dHsm/rd_d3dHsm_in.py (reference file:line
23-64): DIII-D single null 16x8, hydrogen + diffusive atoms, upwind (33)
differencing, fixed core density/temperatures, recycling plates.
"""
import os
import re

import numpy as np

from .case import Case
from .gridue import read_gridue
from .h5lite import read_h5
from .slabgrid import idealgrd

_HERE = os.path.dirname(os.path.abspath(__file__))


def d3dhsm_case(grid, istabon=0, v8_0_defaults=True, cls=Case):
    """v8_0_defaults: oldseec=1, isoldalbarea=1 — the defaults of UEDGE 8.0.x that wrote the
    d3dHsm.h5 restart file (changed in 8.1: src/uedge/defaults.yaml)."""
    c = cls(grid)
    b, com = c.bbb, c.com
    f = grid["nxm"] // 16
    fy = grid["nym"] // 8
    com.nxleg = np.array([[4 * f, 4 * f]]); com.nxcore = np.array([[4 * f, 4 * f]])
    com.nysol = np.array([6 * fy]); com.nycore = np.array([2 * fy])
    if v8_0_defaults:
        b.oldseec = 1.0
        b.isoldalbarea = 1.0
    b.methn = b.methu = b.methe = b.methi = b.methg = 33
    b.ncore[0] = 2.5e19
    b.tcoree = 100.0; b.tcorei = 100.0; b.tedge = 2.0
    b.recycp[0] = 0.8
    b.difni[0] = 1.0; b.kye = 1.0; b.kyi = 1.0; b.travis[0] = 1.0
    b.flalfe = 0.21; b.flalfi = 0.21; b.flalfv = 1.0
    b.flalfgx = np.full(10, 1.0e20); b.flalfgy = np.full(10, 1.0e20)
    com.istabon = istabon
    return c


def forthon_case1(cls=Case):
    """builder/test/Forthon_cases/Forthon_case1/rd_forthon_case1.py: slab (mhdgeo=-1) 6x10, symmetry plane at ix=0
    (isfixlb=2), no core region (nycore=0), four unknowns per cell (ni, up, te, ti; isngon=0: frozen atom density)."""
    g = idealgrd(nxleg2=2, nxcore2=4, nycore=0, nysol=10, zax=1.0, zaxpt=0.75, alfyt=-1.0e-5)
    c = cls(g)
    b, com = c.bbb, c.com
    com.nxleg = np.array([[0, 2]]); com.nxcore = np.array([[0, 4]])
    com.nysol = np.array([10]); com.nycore = np.array([0])
    b.isngon = np.zeros_like(b.isngon)
    b.isfixlb = np.array([2, 0])
    b.ncore[0] = 2.0e19; b.tcoree = 100.0; b.tcorei = 100.0
    b.recycp[0] = 0.9
    b.n0g = np.full_like(np.asarray(b.n0g, dtype=float), 1.0e16)
    b.difni[0] = 1.0; b.kye = 1.0
    b.flalfe = 0.21; b.flalfi = 0.21
    b.flalfgx = np.full(10, 1.0e10); b.flalfgy = np.full(10, 1.0e10)
    return c


def box2_case(isupgon=0, cls=Case):
    """pyexamples/box2/box2_in.py:12-140: slab 6x6 with a core region (nycore=2), symmetry plane at ix=0, core power
    boundary condition.  isupgon=0 gives the diffusive-atom variant of the deck (isngon=1, the d3dHsm-family kernels);
    isupgon=1 is the deck as it runs (inertial atoms as ion species 2, nhsp=2, box2_in.py:114-131): a Case2 for the general
    path (include/ue_gen.h)."""
    if isupgon != 0:
        from .cases2 import box2_inertial_case
        return box2_inertial_case()
    g = idealgrd(nxleg2=3, nxcore2=3, nycore=2, nysol=4, radx=4.0e-2, rad0=0.0, radm=-1.0e-2, za0=0.0, zax=3.0, zaxpt=2.25,
                 alfyt=-2.0, alfxt=2.76, btfix=2.0, bpolfix=0.2)
    c = cls(g)
    b, com = c.bbb, c.com
    com.nxleg = np.array([[0, 3]]); com.nxcore = np.array([[0, 3]])
    com.nysol = np.array([4]); com.nycore = np.array([2])
    b.isfixlb = np.array([2, 0])
    b.isnicore[0] = 1; b.ncore[0] = 1.1e19; b.iflcore = 1
    b.tcoree = 25.0; b.tcorei = 25.0; b.pcoree = 2.5e4; b.pcorei = 2.5e4
    b.recycp[0] = 0.98; b.albdsi[0] = 0.99; b.albdso[0] = 0.99
    b.istepfc = 0; b.istipfc = 0; b.istewc = 0; b.istiwc = 0
    b.bcee = 4.0; b.bcei = 2.5; b.bcen = 0.0
    b.isupss[0] = 0; b.isupcore[0] = 0
    b.difni[0] = 0.5; b.kye = 0.7; b.kyi = 0.7; b.travis[0] = 1.0; b.parvis[0] = 1.0
    b.flalfe = 0.2; b.flalfi = 0.2; b.flalfv = 0.5
    b.flalfgx = np.full(10, 1.0); b.flalfgy = np.full(10, 1.0); b.flalfgxy = np.full(10, 1.0)
    b.methn = b.methu = b.methe = b.methi = b.methg = 33
    b.cngfx[0] = 1.0; b.cngfy[0] = 1.0; b.cngflox[0] = 1.0; b.cngfloy[0] = 0.0
    b.cngmom[0] = 1.0; b.eion = 5.0; b.ediss = 10.0; b.isrecmon = 1
    b.cfupcx = 1.0; b.cfticx = 1.0
    com.istabon = 0
    return c


def initial_profiles(c):
    """restart=0 profiles of ueinit (bbb/odesetup.m:1335-1470) for a half-space (isfixlb>0) slab: flat density,
    linearly shaped temperatures and parallel velocity.  Used as a smooth, deterministic test/bench state."""
    b, com = c.bbb, c.com
    nx, ny = com.nx, com.ny
    IY, IX = np.meshgrid(np.arange(ny + 2), np.arange(nx + 2), indexing="ij")
    px = (nx + 3 - IX) / float(nx + 3)
    py = (ny + 3 - IY) / float(ny + 3)
    ttbeg = float(b.tinit) * b.ev if "tinit" in b else 40.0 * b.ev
    te = ttbeg * px * py
    ti = float(b.tscal) * ttbeg * px * py
    ni = float(b.nibeg[0]) * (1.0 + 0.0 * px) * (0.5 + 0.5 * py)
    cs = np.sqrt((te + ti) / (b.minu[0] * b.mp))
    up = 0.3 * cs * (1.0 - px)
    return ni, up, te, ti, c.initial_ng()


GOLDEN = os.path.join(os.path.dirname(_HERE), "tests", "golden")


def load_grid_npz(path=None):
    """Mesh fixture written by tools/make_golden.py (same content as a gridue file)."""
    z = np.load(path or os.path.join(GOLDEN, "d3d_16x8_grid.npz"))
    nxm, nym, ixpt1, ixpt2, iys = [int(v) for v in z["dims"]]
    g = dict(nxm=nxm, nym=nym, ixpt1=ixpt1, ixpt2=ixpt2, iysptrx1=iys, runid="fixture")
    for k in ("rm", "zm", "psi", "br", "bz", "bpol", "bphi", "b"):
        g[k] = z[k]
    return g


def load_state_npz(name):
    z = np.load(os.path.join(GOLDEN, name))
    return z["ni"], z["up"], z["te"], z["ti"], z["ng"]


def load_rate_tables_npz():
    z = np.load(os.path.join(GOLDEN, "ehr2_tables.npz"))
    return [z["wsveh"], z["wsveh0"], z["welms1"], z["welms2"]]


def refine_grid(g, fx=4, fy=4):
    """Synthetic refinement of a single-null mesh: every interior cell is split fx x fy by bilinear
    subdivision of its four corners in (R,Z); field components are interpolated bilinearly at the new
    corners/centres.  Stand-in for the reference's griddubl/grid-sequencing (bbb/griddubl.m), which
    needs the flx/grd generators; used only for the "4x refined" scaling configuration (BASELINE.json)."""
    nxm, nym = g["nxm"], g["nym"]
    NX, NY = nxm * fx, nym * fy
    out = dict(nxm=NX, nym=NY, ixpt1=g["ixpt1"] * fx, ixpt2=g["ixpt2"] * fx, iysptrx1=g["iysptrx1"] * fy, runid="synthetic %dx%d" % (fx, fy))
    for k in ("rm", "zm", "psi", "br", "bz", "bpol", "bphi", "b"):
        a = g[k]
        o = np.zeros((5, NY + 2, NX + 2))
        for iy in range(1, nym + 1):
            for ix in range(1, nxm + 1):
                c1, c2, c3, c4 = a[1, iy, ix], a[2, iy, ix], a[3, iy, ix], a[4, iy, ix]

                def bil(s, t):
                    return (1 - s) * (1 - t) * c1 + s * (1 - t) * c2 + (1 - s) * t * c3 + s * t * c4

                for jy in range(fy):
                    for jx in range(fx):
                        s0, s1, t0, t1 = jx / fx, (jx + 1) / fx, jy / fy, (jy + 1) / fy
                        X, Y = (ix - 1) * fx + jx + 1, (iy - 1) * fy + jy + 1
                        o[1, Y, X], o[2, Y, X], o[3, Y, X], o[4, Y, X] = bil(s0, t0), bil(s1, t0), bil(s0, t1), bil(s1, t1)
                        o[0, Y, X] = bil(0.5 * (s0 + s1), 0.5 * (t0 + t1))
        out[k] = o
    return out


def refine_state(planes, fx=4, fy=4):
    """Piecewise-constant prolongation of cell planes [iy, ix] (guards kept one cell wide)."""
    res = []
    for a in planes:
        ny, nx = a.shape[0] - 2, a.shape[1] - 2
        o = np.zeros((ny * fy + 2, nx * fx + 2))
        o[1:-1, 1:-1] = np.repeat(np.repeat(a[1:-1, 1:-1], fy, axis=0), fx, axis=1)
        o[0, 1:-1] = np.repeat(a[0, 1:-1], fx); o[-1, 1:-1] = np.repeat(a[-1, 1:-1], fx)
        o[1:-1, 0] = np.repeat(a[1:-1, 0], fy); o[1:-1, -1] = np.repeat(a[1:-1, -1], fy)
        o[0, 0], o[0, -1], o[-1, 0], o[-1, -1] = a[0, 0], a[0, -1], a[-1, 0], a[-1, -1]
        res.append(o)
    return res


def state_from_h5(path):
    """(ni, up, te, ti, ng) planes [iy, ix] from a reference save file; both the
    new `bbb/nis` and the old flat `nis@bbb` dataset names (src/uedge/hdf5.py:31-55)."""
    d = read_h5(path)

    def get(n):
        for k in ("bbb/%s" % n, "%s@bbb" % n):
            if k in d:
                a = d[k]
                if a.ndim == 3:
                    a = a[:, :, 0]
                return np.ascontiguousarray(a.T)  # file is [ix, iy]
        raise KeyError(n)

    return get("nis"), get("ups"), get("tes"), get("tis"), get("ngs")


# ---- named test / benchmark cases (the fixtures live in tests/golden) --------------------------------------------
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
        z = np.load(os.path.join(GOLDEN, "case1_state.npz"))
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


def psetnk_inputs(c, yl):
    """(yl with Jacobian flag, suscal) as psetnk/sfsetnk prepare them (bbb/oderhs.m:9453-9468, 9848-9857)."""
    y = yl.copy()
    y[c.bbb.neq] = 1.0
    return y, c.suscal(yl)


