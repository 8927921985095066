"""Non-orthogonal mesh geometry: angles of the x-faces and the 5-point interpolation stencils.

Host-side restatement (one-time set-up, not on the hot path) of the reference's
  nonorthg   bbb/geometry.m:1879-2397   vtag, angfx, fxm/fx0/fxp/fxmy/fxpy, fym/fy0/fyp/fymx/fypx, dxnog, dynog
  lindis     bbb/geometry.m:2399-2481   intersection of a face normal with a line between two mesh points
  nphygeo    bbb/geometry.m:872-885, 1385-1534, 1546-1579   boundary values, X-point resets, velocity-cell stencils
Arrays are [iy, ix] (and [k, iy, ix] for the two sides k = 0, 1 of a stencil), single null, no limiter.
"""
import numpy as np


def _lindis(rm, zm, i1, j1, i2, j2, ipos2, ish, r0, z0, slp1):
    r1, z1 = rm[0, j1, i1], zm[0, j1, i1]
    if ipos2 == 3:  # 2nd point is the y-face of the 2nd cell
        r2 = 0.5 * (rm[3 - 2 * ish, j2, i2] + rm[4 - 2 * ish, j2, i2])
        z2 = 0.5 * (zm[3 - 2 * ish, j2, i2] + zm[4 - 2 * ish, j2, i2])
    elif ipos2 == 4:  # x-face of the 2nd cell
        r2 = 0.5 * (rm[2 - ish, j2, i2] + rm[4 - ish, j2, i2])
        z2 = 0.5 * (zm[2 - ish, j2, i2] + zm[4 - ish, j2, i2])
    else:
        raise NotImplementedError
    slp2 = (z1 - z2) / (r1 - r2 + 1.0e-20)
    if abs(slp1) > 1.0e-9:
        rx = (z0 + r0 / slp1 - z1 + slp2 * r1) / (slp2 + 1 / slp1)
        zx = z0 - (rx - r0) / slp1
    else:
        rx = r0
        zx = z1 - slp2 * (r1 - r0)
    d1 = np.sqrt((rx - r1) ** 2 + (zx - z1) ** 2)
    d2 = np.sqrt((rx - r2) ** 2 + (zx - z2) ** 2)
    d3 = np.sqrt((r1 - r2) ** 2 + (z1 - z2) ** 2)
    return rx, zx, d1, d2, d3


def nonorthg(case):
    b, c = case.bbb, case.com
    nx, ny = c.nx, c.ny
    rm, zm = case.rz["rm"], case.rz["zm"]
    ixp1, ixm1 = case.ixp1, case.ixm1
    ixpt1, ixpt2, iys1, iys2 = c.ixpt1, c.ixpt2, c.iysptrx1, c.iysptrx2
    pi = b.pi
    shp = (ny + 2, nx + 2)
    vtag = np.zeros(shp)
    for iy in range(0, ny + 1):
        for ix in range(0, nx + 1):
            aa = np.hypot(rm[3, iy, ix] - rm[4, iy, ix], zm[3, iy, ix] - zm[4, iy, ix])
            bb = np.hypot(rm[2, iy, ix] - rm[4, iy, ix], zm[2, iy, ix] - zm[4, iy, ix])
            cc = np.hypot(rm[3, iy, ix] - rm[2, iy, ix], zm[3, iy, ix] - zm[2, iy, ix])
            ang4 = np.arccos((aa**2 + bb**2 - cc**2) / (2 * aa * bb))
            i2 = ixp1[iy + 1, ix]
            aa = np.hypot(rm[3, iy + 1, i2] - rm[1, iy + 1, i2], zm[3, iy + 1, i2] - zm[1, iy + 1, i2])
            bb = np.hypot(rm[2, iy + 1, i2] - rm[1, iy + 1, i2], zm[2, iy + 1, i2] - zm[1, iy + 1, i2])
            cc = np.hypot(rm[3, iy + 1, i2] - rm[2, iy + 1, i2], zm[3, iy + 1, i2] - zm[2, iy + 1, i2])
            ang1 = np.arccos((aa**2 + bb**2 - cc**2) / (2 * aa * bb))
            vtag[iy, ix] = 0.5 * pi - 0.5 * (ang4 + ang1)
        vtag[iy, nx + 1] = vtag[iy, nx]
    angfx = np.zeros(shp)
    for iy in range(ny + 2):
        iym1 = max(0, iy - 1)
        angfx[iy, :] = 0.5 * (vtag[iy, :] + vtag[iym1, :])
    if int(c.redopltvtag) == 1:
        for iy in range(ny + 2):
            angfx[iy, 0] = 2 * angfx[iy, 1] - angfx[iy, 2]
            vtag[iy, 0] = 2 * vtag[iy, 1] - vtag[iy, 2]
            angfx[iy, nx] = 2 * angfx[iy, nx - 1] - angfx[iy, nx - 2]
            vtag[iy, nx] = 2 * vtag[iy, nx - 1] - vtag[iy, nx - 2]
            angfx[iy, nx + 1] = angfx[iy, nx]
            vtag[iy, nx + 1] = vtag[iy, nx]
    angfx[0, :] = angfx[1, :]
    angfx[ny + 1, :] = angfx[ny, :]
    vtag[0, :] = vtag[1, :]
    vtag[ny + 1, :] = vtag[ny, :]
    if ixpt1 > 0:
        for d in (-1, 0, 1):
            angfx[iys1 + d, ixpt1] = 0.0
    if ixpt2 > 0:
        for d in (-1, 0, 1):
            angfx[iys2 + d, ixpt2] = 0.0
    if int(b.isfixlb[0]) == 2 and ixpt2 > 0:
        angfx[0 : iys1 + 1, ixpt2] = 0.0
    S = lambda v: np.full((2,) + shp, v)
    fx0, fxm, fxp, fxmy, fxpy = S(1.0), S(0.0), S(0.0), S(0.0), S(0.0)
    fy0, fym, fyp, fymx, fypx = S(1.0), S(0.0), S(0.0), S(0.0), S(0.0)
    dynog = np.zeros(shp)
    dxnog = np.zeros(shp)
    bigslp = 1.0e20
    fails = []
    itry = 1  # (the reference leaves itry undefined before the first successful search)
    # ---- x-stencil on the y-faces (geometry.m:2052-2148)
    for iy in range(0, ny + 1):
        for ix in range(1, nx + 1):
            if rm[4, iy, ix] == rm[3, iy, ix]:
                slp1 = bigslp
            elif zm[4, iy, ix] == zm[3, iy, ix]:
                slp1 = 1 / bigslp
            else:
                slp1 = (zm[4, iy, ix] - zm[3, iy, ix]) / (rm[4, iy, ix] - rm[3, iy, ix])
            zmid = 0.5 * (zm[4, iy, ix] + zm[3, iy, ix])
            rmid = 0.5 * (rm[4, iy, ix] + rm[3, iy, ix])
            rints, zints = [0.0, 0.0], [0.0, 0.0]
            for ishy in (0, 1):
                ixu1 = ix
                iyu1 = iy + ishy
                iyu2 = iy + 1 - ishy
                if (vtag[iy, ix] + vtag[iy, ixm1[iy, ix]]) * (1 - 2 * ishy) >= 0:
                    ishx, ixu2 = 1, ixp1[iyu2, ix]
                else:
                    ishx, ixu2 = 0, ixm1[iyu2, ix]
                isht = 1 - ishy
                while True:
                    rint, zint, d1, d2, d3 = _lindis(rm, zm, ixu1, iyu1, ixu2, iyu2, 3, isht, rmid, zmid, slp1)
                    if d1 <= d3 * 1.0001 and d2 <= d3 * 1.0001:
                        rints[ishy], zints[ishy] = rint, zint
                        fx0[ishy, iy, ix] = d2 / d3
                        fxm[ishy, iy, ix] = (1 - ishx) * 0.5 * d1 / d3
                        fxp[ishy, iy, ix] = ishx * 0.5 * d1 / d3
                        fxmy[ishy, iy, ix] = (1 - ishx) * 0.5 * d1 / d3
                        fxpy[ishy, iy, ix] = ishx * 0.5 * d1 / d3
                        itry = 1
                        break
                    elif itry == 1:
                        if ishx == 1:
                            ishx, ixu2 = 0, ixm1[iyu2, ix]
                        else:
                            ishx, ixu2 = 1, ixp1[iyu2, ix]
                        itry = 2
                        continue
                    else:
                        fails.append(("fx", ix, iy, ishy))
                        break
            dynog[iy, ix] = np.sqrt((rints[1] - rints[0]) ** 2 + (zints[1] - zints[0]) ** 2)
    for iy in range(0, ny + 1):
        dynog[iy, 0] = dynog[iy, 1]
        dynog[iy, nx + 1] = dynog[iy, nx]
    dynog[ny + 1, :] = 0.1 * dynog[ny, :]
    # ---- y-stencil on the x-faces (geometry.m:2200-2330)
    for ix in range(0, nx + 1):
        for iy in range(1, ny + 1):
            if rm[4, iy, ix] == rm[2, iy, ix]:
                slp1 = bigslp
            elif zm[4, iy, ix] == zm[2, iy, ix]:
                slp1 = 1 / bigslp
            else:
                slp1 = (zm[4, iy, ix] - zm[2, iy, ix]) / (rm[4, iy, ix] - rm[2, iy, ix])
            zmid = 0.5 * (zm[4, iy, ix] + zm[2, iy, ix])
            rmid = 0.5 * (rm[4, iy, ix] + rm[2, iy, ix])
            rints, zints = [0.0, 0.0], [0.0, 0.0]
            for ishx in (0, 1):
                iyu1 = iy
                ixu1 = (1 - ishx) * ix + ishx * ixp1[iyu1, ix]
                if angfx[iy, ix] * (1 - 2 * ishx) >= 0:
                    ishy, iyu2 = 1, iy + 1
                else:
                    ishy, iyu2 = 0, iy - 1
                ixu2 = ishx * ix + (1 - ishx) * ixp1[iyu2, ix]
                isht = 1 - ishx
                while True:
                    rint, zint, d1, d2, d3 = _lindis(rm, zm, ixu1, iyu1, ixu2, iyu2, 4, isht, rmid, zmid, slp1)
                    if d1 <= d3 * 1.0001 and d2 <= d3 * 1.0001:
                        rints[ishx], zints[ishx] = rint, zint
                        fy0[ishx, iy, ix] = d2 / d3
                        fym[ishx, iy, ix] = (1 - ishy) * 0.5 * d1 / d3
                        fyp[ishx, iy, ix] = ishy * 0.5 * d1 / d3
                        fymx[ishx, iy, ix] = (1 - ishy) * 0.5 * d1 / d3
                        fypx[ishx, iy, ix] = ishy * 0.5 * d1 / d3
                        itry = 1
                        break
                    elif itry == 1:
                        if ishy == 1:
                            ishy, iyu2 = 0, iy - 1
                        else:
                            ishy, iyu2 = 1, iy + 1
                        ixu2 = ishx * ix + (1 - ishx) * ixp1[iyu2, ix]
                        itry = 2
                        continue
                    else:
                        fails.append(("fy", ix, iy, ishx))
                        break
            dxnog[iy, ix] = np.sqrt((rints[1] - rints[0]) ** 2 + (zints[1] - zints[0]) ** 2)
    # plate guard cells: orthogonal fy stencil (geometry.m:2366-2392)
    ixlb, ixrb = c.ixlb, c.ixrb
    for a, v in ((fym, 0.0), (fy0, 1.0), (fyp, 0.0), (fymx, 0.0), (fypx, 0.0)):
        a[0, :, ixlb] = v
        a[1, :, ixrb] = v
    # nphygeo after the call (geometry.m:876-884)
    dxnog[0, :] = dxnog[1, :]
    dxnog[ny + 1, :] = dxnog[ny, :]
    dxnog[:, nx + 1] = 0.1 * dxnog[:, nx]
    return dict(vtag=vtag, angfx=angfx, fx0=fx0, fxm=fxm, fxp=fxp, fxmy=fxmy, fxpy=fxpy, fy0=fy0, fym=fym, fyp=fyp, fymx=fymx, fypx=fypx,
                dxnog=dxnog, dynog=dynog, fails=fails)


def finish_nonog(case, gx, gy, gxf, gyf):
    """nphygeo, geometry.m:1385-1534 and 1546-1579: X-point / cut resets (these also reset gyf) and the velocity-cell stencils."""
    b, c = case.bbb, case.com
    nx, ny = c.nx, c.ny
    g = case.nog
    ixpt1, ixpt2, iys1, iys2 = c.ixpt1, c.ixpt2, c.iysptrx1, c.iysptrx2
    dxnog, dynog = g["dxnog"], g["dynog"]
    fixlb = int(b.isfixlb[0])
    if fixlb == 0 and int(b.isfixrb[0]) == 0:
        for ij in (0, 1):
            for ixp, iysp in ((ixpt1, iys1), (ixpt2, iys2)):
                for k in (0, 1):
                    g["fxm"][k, iysp, ixp + ij] = 0.0; g["fx0"][k, iysp, ixp + ij] = 1.0; g["fxp"][k, iysp, ixp + ij] = 0.0
                    g["fxmy"][k, iysp, ixp + ij] = 0.0; g["fxpy"][k, iysp, ixp + ij] = 0.0
                gyf[iysp, ixp + ij] = 2 * gy[iysp, ixp + ij] * gy[iysp + 1, ixp + ij] / (gy[iysp, ixp + ij] + gy[iysp + 1, ixp + ij])
                dynog[iysp, ixp + ij] = 1.0 / gyf[iysp, ixp + ij]
        for ij in (0, 1):
            for ixp, iysp in ((ixpt1, iys1), (ixpt2, iys2)):
                for k in (0, 1):
                    g["fym"][k, iysp + ij, ixp] = 0.0; g["fy0"][k, iysp + ij, ixp] = 1.0; g["fyp"][k, iysp + ij, ixp] = 0.0
                    g["fypx"][k, iysp + ij, ixp] = 0.0; g["fymx"][k, iysp + ij, ixp] = 0.0
            # (the reference indexes dxnog with iysptrx2 for both X-point columns)
            dxnog[iys2 + ij, ixpt1] = 1.0 / gxf[iys1 + ij, ixpt1]
            dxnog[iys2 + ij, ixpt2] = 1.0 / gxf[iys2 + ij, ixpt2]
    if fixlb == 2:
        ix = ixpt2
        for k in (0, 1):
            for iy in range(0, iys1 + 2):
                for ij in (0, 1):
                    g["fxm"][k, iy, ix + ij] = 0.0; g["fx0"][k, iy, ix + ij] = 1.0; g["fxp"][k, iy, ix + ij] = 0.0
                    g["fxmy"][k, iy, ix + ij] = 0.0; g["fxpy"][k, iy, ix + ij] = 0.0
                    gyf[iy, ix + ij] = 2 * gy[iy, ix + ij] * gy[iy + 1, ix + ij] / (gy[iy, ix + ij] + gy[iy + 1, ix + ij])
        ix2 = 0
        for k in (0, 1):
            for iy in range(0, iys1 + 2):
                g["fym"][k, iy, ix] = 0.0; g["fy0"][k, iy, ix] = 1.0; g["fyp"][k, iy, ix] = 0.0; g["fypx"][k, iy, ix] = 0.0; g["fymx"][k, iy, ix] = 0.0
                dxnog[iy, ix] = 1.0 / gxf[iy, ix]
            for iy in range(0, ny + 2):
                g["fym"][k, iy, ix2] = 0.0; g["fy0"][k, iy, ix2] = 1.0; g["fyp"][k, iy, ix2] = 0.0; g["fypx"][k, iy, ix2] = 0.0; g["fymx"][k, iy, ix2] = 0.0
                dxnog[iy, ix2] = 1.0 / gxf[iy, ix2]
    # velocity-cell stencils (geometry.m:1546-1579)
    ixp1, ixm1 = case.ixp1, case.ixm1
    for nm in ("fxm", "fx0", "fxp", "fxmy", "fxpy"):
        a = g[nm]; v = np.zeros_like(a)
        for iy in range(ny + 2):
            for ix in range(nx + 2):
                v[:, iy, ix] = 0.5 * (a[:, iy, ix] + a[:, iy, ixp1[iy, ix]])
        g[nm + "v"] = v
    for nm in ("fym", "fy0", "fyp", "fymx", "fypx"):
        a = g[nm]; v = np.zeros_like(a)
        for iy in range(ny + 2):
            for ix in range(nx + 2):
                v[:, iy, ix] = 0.5 * (a[:, iy, ix] + a[:, iy, ixm1[iy, ix]])
        g[nm + "v"] = v
    g["fxmv"][:, :, c.ixlb] = 0.0; g["fxmyv"][:, :, c.ixlb] = 0.0
    g["fxpv"][:, :, c.ixrb] = 0.0; g["fxpyv"][:, :, c.ixrb] = 0.0
    return dxnog, dynog
