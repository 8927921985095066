"""Analytic slab / cylinder mesh of the reference (`mhdgeo = -1 / 0`, `gengrid = 1`).

Restates grd/grdcomp.m:102-365 (`idealgrd`/`idlcomp`) for the options the `box2` and
`Forthon_case1` decks use (tnoty = 0, no tilt, no distortion, isgrdsym = 0, isadjalfxt = 0):
uniform poloidal cells from the symmetry plane `za0` to the X-point position `zaxpt`, exponentially
shrinking cells from there to the plate at `zax`, exponentially stretched radial cells on both sides of
the "separatrix" `rad0`, constant total and poloidal field.  Index bookkeeping for this geometry:
com/comutil.m:3-30 (`nxm`, `nym`) and bbb/odesetup.m:128-137 (`ixpt1 = -1`, `ixpt2 = nxcore(1,2)`,
`iysptrx1 = nycore`).

The result has the layout of `gridue.read_gridue` (arrays [5, nym+2, nxm+2], interior cells filled)
so that `Case` treats both mesh sources alike.
"""
import math

import numpy as np

ANALGRD_DEFAULTS = dict(  # grd/grd.v, group Analgrd
    radm=-1.0e-4, radx=0.04, rad0=0.0, rscalcore=1.0, za0=0.0, zax=1.0, zaxpt=0.75,
    alfyt=-2.0, tnoty=0.0, sratiopf=0.0, alfxt=4.0, tctr=0.0,
    bpolfix=0.3, btfix=5.0, rmajfix=1.0, sigma_bpol=0.0, sigma_btor=0.0,
)


def idealgrd(nxleg2, nxcore2, nycore, nysol, **kw):
    p = dict(ANALGRD_DEFAULTS)
    for k, v in kw.items():
        if k not in p:
            raise KeyError("unknown Analgrd parameter %s" % k)
        p[k] = float(v)
    if p["tnoty"] != 0.0:
        raise NotImplementedError("tnoty != 0 (tanh radial mesh)")
    nxm, nym = nxleg2 + nxcore2, nycore + nysol  # nxleg(1,1) = nxcore(1,1) = 0 (com/com.v:242-244)
    ixpt2, iys = nxcore2, nycore
    radm, radx, rad0 = p["radm"], p["radx"], p["rad0"]
    za0, zax, zaxpt, alfyt, alfxt, tctr = p["za0"], p["zax"], p["zaxpt"], p["alfyt"], p["alfxt"], p["tctr"]
    if iys == 0:
        rad0 = radm  # grdcomp.m:157
    dznu = (zaxpt - za0) / float(ixpt2) if ixpt2 > 0 else 0.0
    dznl = tctr * (zax - zaxpt - za0) / (1 - math.exp(+alfxt * tctr)) if tctr > 0.0000001 else 0.0
    dznr = (1 - tctr) * (zax - zaxpt - za0) / (1 - math.exp(-alfxt * (1 - tctr))) if tctr < 0.9999999 else 0.0
    ixm = int(tctr * float(nxm - ixpt2) + 0.5)
    sratiopf = p["sratiopf"]
    rm = np.zeros((5, nym + 2, nxm + 2))
    zm = np.zeros((5, nym + 2, nxm + 2))
    for iy in range(nym, 0, -1):
        if iy > iys:  # grdcomp.m:175-191
            t1y = float(iy - 1 - iys) / float(nym - iys)
            t2y = float(iy - iys) / float(nym - iys)
            drn1 = (radx - rad0) * (1 - math.exp(-alfyt * t1y)) / (1 - math.exp(-alfyt))
            drn2 = (radx - rad0) * (1 - math.exp(-alfyt * t2y)) / (1 - math.exp(-alfyt))
        else:  # grdcomp.m:192-213
            if sratiopf == 0.0:
                sratiopf = (rad0 - radm) / (radx - rad0)
            t1y = -sratiopf * float(iy - 1 - iys) / float(iys)
            t2y = -sratiopf * float(iy - iys) / float(iys)
            drn1 = (radm - rad0) * (1 - math.exp(-alfyt * t1y)) / (1 - math.exp(-sratiopf * alfyt))
            drn2 = (radm - rad0) * (1 - math.exp(-alfyt * t2y)) / (1 - math.exp(-sratiopf * alfyt))
        for ix in range(1, nxm + 1):
            sc = p["rscalcore"] if (ix <= ixpt2 and iy <= iys) else 1.0
            rm[1, iy, ix] = rad0 + sc * drn1
            rm[2, iy, ix] = rm[1, iy, ix]
            rm[3, iy, ix] = rad0 + sc * drn2
            rm[4, iy, ix] = rm[3, iy, ix]
            rm[0, iy, ix] = 0.25 * (rm[1, iy, ix] + rm[2, iy, ix] + rm[3, iy, ix] + rm[4, iy, ix])
            if ix <= ixpt2:  # uniform region
                z1 = za0 + dznu * (ix - 1)
                z2 = za0 + dznu * ix
            else:
                t1x = float(ix - ixpt2 - 1) / float(nxm - ixpt2)
                t2x = float(ix - ixpt2) / float(nxm - ixpt2)
                if ix > ixm:  # decreasing dx towards the plate
                    z1 = zaxpt + dznr * (1 - math.exp(-alfxt * (t1x - tctr)))
                    z2 = zaxpt + dznr * (1 - math.exp(-alfxt * (t2x - tctr)))
                else:
                    z1 = zaxpt + dznl * (1 - math.exp(+alfxt * t1x))
                    z2 = zaxpt + dznl * (1 - math.exp(+alfxt * t2x))
            zm[1, iy, ix], zm[2, iy, ix], zm[3, iy, ix], zm[4, iy, ix] = z1, z2, z1, z2
            zm[0, iy, ix] = 0.25 * (z1 + z2 + z1 + z2)
    # constant fields, grdcomp.m:341-361
    btorfix = math.sqrt(p["btfix"] ** 2 - p["bpolfix"] ** 2)
    rr = rm / p["rmajfix"]
    with np.errstate(all="ignore"):
        bphi = btorfix * (rr ** p["sigma_btor"] if p["sigma_btor"] != 0.0 else np.ones_like(rr))
        bpol = p["bpolfix"] * (rr ** p["sigma_bpol"] if p["sigma_bpol"] != 0.0 else np.ones_like(rr))
        psi = ((p["bpolfix"] * p["rmajfix"] ** 2) / (p["sigma_bpol"] + 2)) * np.abs(rr) ** (p["sigma_bpol"] + 2)
    b = np.sqrt(bphi ** 2 + bpol ** 2)
    g = dict(nxm=nxm, nym=nym, ixpt1=-1, ixpt2=ixpt2, iysptrx1=iys, runid="ideal geometry",
             rm=rm, zm=zm, psi=psi, br=np.zeros_like(rm), bz=-bpol, bpol=bpol, bphi=bphi, b=b, slab=True)
    return g
