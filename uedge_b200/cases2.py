"""Named configurations of the general hydrogen family (inputs of oracle/ue_oracle2.cpp)."""
import ctypes as C
import os

import numpy as np

from .case2 import Case2
from .cases import GOLDEN, load_grid_npz

SUBSETS = ("default", "ni", "ni-0", "ni-1", "up", "up-0", "up-1", "te", "ti", "phi")


def inputex_case(subset="default"):
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
    c.setup()
    T = lambda a: np.ascontiguousarray(a.T)
    nis, ups = z["bbb__nis"], z["bbb__ups"]
    yl = c.set_state2([T(nis[:, :, 0]), T(nis[:, :, 1])], [T(ups[:, :, 0]), T(ups[:, :, 1])], T(z["bbb__tes"]), T(z["bbb__tis"]),
                      ng=T(z["bbb__ngs"][:, :, 0]), phi=T(z["bbb__phis"]), tg=T(z["bbb__tgs"][:, :, 0]))
    gold = {k[len(p):]: z[k] for k in z.files if k.startswith(p)}
    return c, yl, gold


class Oracle2:
    """ctypes binding of oracle/libue_oracle2.so (checker only)."""

    def __init__(self, path=None):
        here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        self.lib = C.CDLL(path or os.path.join(here, "oracle", "libue_oracle2.so"))
        self.lib.ue_or2_last_error.restype = C.c_char_p
        self.lib.ue_or2_set.argtypes = [C.c_char_p, C.c_void_p, C.c_int64]

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, self.lib.ue_or2_last_error().decode()))

    def bind(self, c):
        self.c = c
        self.lib.ue_or2_clear()
        self.keep = c.inputs2()
        for k, v in self.keep.items():
            self.lib.ue_or2_set(k.encode(), v.ctypes.data_as(C.c_void_p), v.size)
        self._ck(self.lib.ue_or2_init(), "init")
        self.neq = int(c.bbb.neq); self.NC = (c.com.nx + 2) * (c.com.ny + 2)
        return self

    def step_params(self, dtuse, ylodt, su, sf):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (dtuse, ylodt, su, sf)]
        self.lib.ue_or2_step_params.argtypes = [C.c_int64] + [C.c_void_p] * 4
        self._ck(self.lib.ue_or2_step_params(self.neq, *[x.ctypes.data_as(C.c_void_p) for x in a]), "step_params")

    def pandf1(self, yl, xc=-1, yc=-1, out=None):
        yl = np.ascontiguousarray(yl, dtype=np.float64)
        yd = np.zeros(self.neq) if out is None else out
        self.lib.ue_or2_pandf1_win.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        self._ck(self.lib.ue_or2_pandf1_win(xc, yc, self.neq, yl.ctypes.data_as(C.c_void_p), yd.ctypes.data_as(C.c_void_p)), "pandf1")
        return yd

    def jac_calc(self, yl, f0, ml, mu, nnzmx):
        yl = np.ascontiguousarray(yl, dtype=np.float64)
        y0 = np.zeros(self.neq + 2); y0[: self.neq] = f0[: self.neq]
        jac = np.zeros(nnzmx); ja = np.zeros(nnzmx, dtype=np.int64); ia = np.zeros(self.neq + 1, dtype=np.int64); nnz = C.c_int64(0)
        self.lib.ue_or2_jac_calc.argtypes = [C.c_int64, C.c_double] + [C.c_void_p] * 2 + [C.c_int64] * 3 + [C.c_void_p] * 3 + [C.POINTER(C.c_int64)]
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        self._ck(self.lib.ue_or2_jac_calc(self.neq, 0.0, P(yl), P(y0), int(ml), int(mu), int(nnzmx), P(jac), P(ja), P(ia), C.byref(nnz)), "jac_calc")
        n = nnz.value
        return jac[:n].copy(), ja[:n].copy(), ia

    def plane(self, name):
        out = np.zeros(self.NC)
        self.lib.ue_or2_get_plane.argtypes = [C.c_char_p, C.c_void_p]
        self._ck(self.lib.ue_or2_get_plane(name.encode(), out.ctypes.data_as(C.c_void_p)), "get_plane " + name)
        return out.reshape(self.c.com.ny + 2, self.c.com.nx + 2)
