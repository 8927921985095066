"""Reader for UEDGE ASCII ``gridue`` mesh files.

Format follows the reference's ``readgrid``/``rdgrid`` (grd/grdread.m:199-274):
header ``5i4`` = nxm, nym, ixpt1, ixpt2, iysptrx1 (single-null), then eight
blocks ``rm, zm, psi, br, bz, bpol, bphi, b``, each ``(0:nxm+1, 0:nym+1, 0:4)``
in Fortran order written ``3e23.15`` with a blank line in front, then runid.

Arrays are returned with shape ``(5, nym+2, nxm+2)`` (C order), i.e. indexed
``a[n, iy, ix]`` so that ``ix`` is the fastest-varying index exactly as in the
Fortran storage ``a(ix,iy,n)``.
"""
import re

import numpy as np

_NUM = re.compile(r"[-+]?\d*\.\d+(?:[DdEe][-+]?\d+)?")
FIELDS = ("rm", "zm", "psi", "br", "bz", "bpol", "bphi", "b")


def read_gridue(path):
    with open(path, "r") as f:
        lines = f.read().splitlines()
    hdr = lines[0]
    ints = [int(hdr[i : i + 4]) for i in range(0, 20, 4)]
    nxm, nym, ixpt1, ixpt2, iysptrx1 = ints
    n = (nxm + 2) * (nym + 2) * 5
    body = " ".join(lines[1:])
    toks = _NUM.findall(body)
    need = n * len(FIELDS)
    if len(toks) < need:
        raise ValueError("gridue: expected %d reals, found %d" % (need, len(toks)))
    vals = np.array([float(t.replace("D", "E").replace("d", "e")) for t in toks[:need]])
    out = {
        "nxm": nxm,
        "nym": nym,
        "ixpt1": ixpt1,
        "ixpt2": ixpt2,
        "iysptrx1": iysptrx1,
        "runid": lines[-1].strip(),
    }
    for k, name in enumerate(FIELDS):
        out[name] = vals[k * n : (k + 1) * n].reshape(5, nym + 2, nxm + 2).copy()
    return out
