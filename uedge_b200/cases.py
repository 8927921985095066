"""Named configurations of BASELINE.json, built on `Case`.

d3dHsm: switch set of pyexamples/d3dHsm/rd_d3dHsm_in.py (reference file:line
23-64): DIII-D single null 16x8, hydrogen + diffusive atoms, upwind (33)
differencing, fixed core density/temperatures, recycling plates.
"""
import os

import numpy as np

from .case import Case
from .gridue import read_gridue
from .h5lite import read_h5

_HERE = os.path.dirname(os.path.abspath(__file__))


def d3dhsm_case(grid, istabon=0):
    c = Case(grid)
    b, com = c.bbb, c.com
    com.nxleg = np.array([[4, 4]]); com.nxcore = np.array([[4, 4]])
    com.nysol = np.array([6]); com.nycore = np.array([2])
    b.methn = b.methu = b.methe = b.methi = b.methg = 33
    b.ncore[0] = 2.5e19
    b.tcoree = 100.0; b.tcorei = 100.0; b.tedge = 2.0
    b.recycp[0] = 0.8
    b.difni[0] = 1.0; b.kye = 1.0; b.kyi = 1.0; b.travis[0] = 1.0
    b.flalfe = 0.21; b.flalfi = 0.21; b.flalfv = 1.0
    b.flalfgx = np.full(10, 1.0e20); b.flalfgy = np.full(10, 1.0e20)
    com.istabon = istabon
    return c


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
