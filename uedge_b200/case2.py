"""Host-side case setup for the general hydrogen family (inertial atoms as ion species 2, non-orthogonal meshes,
potential equation, arbitrary subsets of equations): the inputs of the general oracle `oracle/ue_oracle2.cpp`.

`Case2` extends `Case` (same geometry restatements) with the species-indexed parts of `allocate`/`ueinit`/`convert`/`idalg`
(bbb/odesetup.m:272-365, 900-1045; bbb/convert.m:25-155; bbb/boundary.m:3701-3870; bbb/boundary.m:4537-4645 recyprof).
Every scalar of the `bbb`/`com` namespaces and every array crosses to the oracle as doubles under its own name.
"""
import numpy as np

from .case import Case


class Case2(Case):
    def setup(self):
        b, c = self.bbb, self.com
        nx, ny = c.nx, c.ny
        c.nzspt = 0
        nhsp = int(c.nhsp)
        c.nisp = nhsp
        c.nusp = nhsp
        ngsp = int(c.ngsp)
        if nhsp not in (1, 2) or ngsp != 1:
            raise NotImplementedError("nhsp must be 1 or 2, ngsp 1")
        nisp = nusp = nhsp
        isn = [int(b.isnion[f]) for f in range(nisp)]
        isu = [int(b.isupon[f]) for f in range(nusp)]
        self.isn, self.isu = isn, isu
        b.numvar = int(b.isteon) + int(b.istion) + sum(isn) + sum(isu) + int(b.isngon[0]) + int(b.istgon[0]) + int(b.isphion)
        numvar = b.numvar
        b.neq = numvar * (nx + 2) * (ny + 2)
        self.hasg = int(b.isngon[0]) == 1
        self._nphygeo()
        b.ubw = (numvar + b.numvarbwpad) * (nx + c.ixpt2 - max(0, c.ixpt1) + 4)
        b.lbw = b.ubw
        b.nnzmx = (9 * numvar if b.lenpfac < 9 * numvar else b.lenpfac) * b.neq
        # ueinit, odesetup.m:900-991
        b.mi = np.array([b.minu[f] * b.mp for f in range(nisp)])
        b.zi = np.array([float(b.ziin[f]) for f in range(nisp)])
        b.mg = np.array([b.facmg[0] * b.mi[0]])
        b.nnorm = float(b.n0[0])
        b.fnorm = np.array([b.n0[f] * np.sqrt(b.mi[f] * b.temp0 * b.ev) for f in range(nusp)])
        b.ennorm = 1.5 * b.n0[0] * b.temp0 * b.ev
        b.vpnorm = float(np.sqrt(b.temp0 * b.ev / b.mi[0]))
        b.sigbar0 = float(b.cfsigm * b.sigma1 * b.temp0 ** 1.5)
        nc = [b.n0[f] for f in range(nisp) if isn[f]] + [b.fnorm[f] for f in range(nusp) if isu[f]]
        isup = [0] * sum(isn) + [1] * sum(isu)
        isphi = []
        for on in (int(b.isteon), int(b.istion)):
            if on:
                nc.append(b.ennorm); isup.append(0)
        if self.hasg:
            nc.append(b.n0g[0]); isup.append(0)
        if int(b.istgon[0]):
            nc.append(b.ennorm); isup.append(0)
        isphi = [0] * len(nc)
        if int(b.isphion):
            nc.append(b.temp0); isup.append(0); isphi.append(1)
        norm_cons = np.array(nc, dtype=float)
        floor_cons = b.var_scale_floor * norm_cons
        for i in range(len(nc)):
            if isup[i]:
                floor_cons[i] = b.vsf_up * norm_cons[i]
            if isphi[i]:
                floor_cons[i] = b.vsf_phi * norm_cons[i]
        self.norm_cons, self.floor_cons = norm_cons, floor_cons
        # flux-limit profile arrays (odesetup.m:1258-1330), incl. the inertial-atom ones
        ixlb, ixrb = c.ixlb, c.ixrb
        a1 = {}
        a1["fgtdx"] = np.ones(nx + 2); a1["fgtdy"] = np.ones(ny + 2)
        a1["fgtdx"][ixlb] = b.gcfacgtx; a1["fgtdx"][ixrb] = b.gcfacgtx
        a1["fgtdy"][0] = b.gcfacgty; a1["fgtdy"][ny] = b.gcfacgty
        f0 = lambda v: float(np.atleast_1d(v)[0])
        for k, src, n in (("flalfea", b.flalfe, nx), ("flalfia", b.flalfi, nx), ("flalfva", b.flalfv, nx), ("flalfgxa", b.flalfgx, nx),
                          ("flalfgxya", b.flalfgxy, nx), ("flalfgya", b.flalfgy, ny), ("flalfvgxa", b.flalfvgx, nx), ("flalfvgya", b.flalfvgy, ny),
                          ("flalfvgxya", b.flalfvgxy, nx), ("flalftgxa", b.flalftgx, nx), ("flalftgya", b.flalftgy, ny)):
            a1[k] = np.full(n + 2, f0(src))
        if b.isplflxl == 0:
            for k in ("flalfea", "flalfia"):
                a1[k][ixlb] = 1e20; a1[k][ixrb] = 1e20
        if b.isplflxlv == 0:
            a1["flalfva"][ixlb + 1] = 1e20; a1["flalfva"][ixrb + 1] = 1e20
        if b.isplflxlgx == 0:
            for k in ("flalfgxa", "flalfgxya"):
                a1[k][ixlb] = 1e20; a1[k][ixrb] = 1e20
        if b.iswflxlgy == 0:
            a1["flalfgya"][0] = 1e20; a1["flalfgya"][ny] = 1e20
        if b.isplflxlvgx == 0:  # staggered mesh: ixlb+1 is the boundary viscosity
            for k in ("flalfvgxa", "flalfvgxya"):
                a1[k][ixlb + 1] = 1e20; a1[k][ixrb + 1] = 1e20
        if b.iswflxlvgy == 0:
            a1["flalfvgya"][0] = 1e20; a1["flalfvgya"][ny] = 1e20
        if b.isplflxltgx == 0:
            a1["flalftgxa"][ixlb] = 1e20; a1["flalftgxa"][ixrb] = 1e20
        if b.iswflxltgy == 0:
            a1["flalftgya"][0] = 1e20; a1["flalftgya"][ny] = 1e20
        self.a1 = a1
        # setwallbcarrays, odesetup.m:557-606
        I = lambda v: int(np.atleast_1d(v)[0])
        bc = dict(
            istepfcix=np.full(nx + 2, I(b.istepfc)), istipfcix=np.full(nx + 2, I(b.istipfc)),
            istewcix=np.full(nx + 2, I(b.istewc)), istiwcix=np.full(nx + 2, I(b.istiwc)),
            iphibcwiix=np.full(nx + 2, I(b.iphibcwi)), iphibcwoix=np.full(nx + 2, I(b.iphibcwo)),
        )
        # species-indexed wall switches (ix, ifld): [ifld][ix]
        sp = lambda v: np.concatenate([np.full(nx + 2, int(np.atleast_1d(v)[f])) for f in range(2)])
        bc["isnwconiix"] = sp(b.isnwconi); bc["isnwconoix"] = sp(b.isnwcono); bc["isupwiix"] = sp(b.isupwi); bc["isupwoix"] = sp(b.isupwo)
        for k in ("nwalli", "nwallo"):
            bc[k] = np.asarray(getattr(b, k), dtype=float) * np.ones(nx + 2) if k in b else np.zeros(nx + 2)
        for k, src, i in (("lytepf", "lyte", 0), ("lytewc", "lyte", 1), ("lytipf", "lyti", 0), ("lytiwc", "lyti", 1), ("lynipf", "lyni", 0), ("lyniwc", "lyni", 1),
                          ("lyphiix1", "lyphi", 0), ("lyphiix2", "lyphi", 1)):
            bc[k] = np.full(nx + 2, float(getattr(b, src)[i]))
        for k in ("tewalli", "tiwalli", "tewallo", "tiwallo"):
            bc[k] = np.full(nx + 2, float(b.tedge))
        # recyprof, boundary.m:4558-4642 (no user profiles): plate recycling / albedo / momentum recycling of gas species 1
        bc["recylb"] = np.full(ny + 2, float(b.recycp[0])); bc["recyrb"] = np.full(ny + 2, float(b.recycp[0]))
        bc["alblb"] = np.full(ny + 2, 1.0); bc["albrb"] = np.full(ny + 2, 1.0)
        bc["recycmlb"] = np.full(ny + 2, float(b.recycm)); bc["recycmrb"] = np.full(ny + 2, float(b.recycm))
        bc["recycwot"] = np.full(nx + 2, float(b.recycw[0])); bc["recycwit"] = np.full(nx + 2, float(b.recycw[0]))
        for k in ("fngysi", "fngyso", "fngyi_use", "fngyo_use"):
            bc[k] = np.zeros(nx + 2)
        for k in ("fngxslb", "fngxsrb", "fngxlb_use", "fngxrb_use", "phi0l", "phi0r", "bctype"):
            bc[k] = np.zeros(ny + 2)
        if int(b.nwsor) != 1 or float(b.wgaso[0]) < 50.0 or float(b.wgasi[0]) < 50.0 or float(b.igaso[0]) != 0.0 or float(b.igasi[0]) != 0.0:
            raise NotImplementedError("wall gas-source regions other than the default all-covering, zero-current one")
        bc["albedoi"] = np.ones(nx + 2); bc["albedoo"] = np.ones(nx + 2)
        bc["matwalli"] = np.zeros(nx + 2); bc["matwallo"] = np.zeros(nx + 2)
        bc["albedoi"][1 : nx + 1] = float(b.albdsi[0]); bc["albedoo"][1 : nx + 1] = float(b.albdso[0])
        bc["matwalli"][1 : nx + 1] = int(b.matwsi[0]); bc["matwallo"][1 : nx + 1] = int(b.matwso[0])
        self.bc1 = bc
        for k, cf in (("nurlxn", "cnurn"), ("nurlxu", "cnuru"), ("nurlxe", "cnure"), ("nurlxi", "cnuri"), ("nurlxg", "cnurg"), ("nurlxp", "cnurp")):
            setattr(b, k, float(getattr(b, cf)) * float(b.nurlx))
        self._index_maps2()
        if not hasattr(self, "rate_tables"):
            self.rate_tables = (1, 1, [np.ones(1)] * 4)

    def _index_maps2(self):
        """convert (bbb/convert.m:25-155) ordering and idalg (boundary.m:3701-3870) for any subset of equations."""
        b, c = self.bbb, self.com
        nx, ny = c.nx, c.ny
        shp = (ny + 2, nx + 2)
        idxn = np.zeros((2,) + shp, dtype=np.int64); idxu = np.zeros((2,) + shp, dtype=np.int64)
        idxte = np.zeros(shp, dtype=np.int64); idxti = np.zeros(shp, dtype=np.int64); idxg = np.zeros(shp, dtype=np.int64); idxtg = np.zeros(shp, dtype=np.int64); idxphi = np.zeros(shp, dtype=np.int64)
        igyl = np.zeros((b.neq, 2), dtype=np.int64)
        iv = 0
        for iy in range(ny + 2):
            for ix in range(nx + 2):
                def put(arr, *k):
                    nonlocal iv
                    iv += 1
                    arr[k + (iy, ix)] = iv
                    igyl[iv - 1] = (ix, iy)
                for f in range(c.nisp):
                    if self.isn[f]:
                        put(idxn, f)
                for f in range(c.nusp):
                    if self.isu[f]:
                        put(idxu, f)
                if int(b.isteon):
                    put(idxte)
                if int(b.istion):
                    put(idxti)
                if self.hasg:
                    put(idxg)
                if int(b.istgon[0]):  # gas temperature follows the gas density (convert.m:120-133)
                    put(idxtg)
                if int(b.isphion):
                    put(idxphi)
        assert iv == b.neq
        self.idx = dict(idxn=idxn, idxu=idxu, idxte=idxte, idxti=idxti, idxg=idxg, idxtg=idxtg, idxphi=idxphi)
        self.igyl = igyl
        alg = np.zeros(b.neq, dtype=np.int64)
        def mark(a):
            v = a[a > 0]
            alg[v - 1] = 1
        for arr in [idxn[0], idxn[1], idxu[0], idxu[1], idxte, idxti, idxg, idxtg, idxphi]:
            mark(arr[0, :]); mark(arr[ny + 1, :]); mark(arr[1 : ny + 1, 0]); mark(arr[1 : ny + 1, nx + 1])
        for f in range(2):
            mark(idxu[f][1 : ny + 1, nx])  # boundary.m:3812-3815
        mark(idxphi[1 : ny + 1, 0 : nx + 1])  # boundary.m:3796-3800: the potential equation is algebraic everywhere
        if int(b.isfixlb[0]) == 2 and c.ixpt2 >= 0:
            for f in range(2):
                mark(idxu[f][0 : c.iysptrx1 + 1, c.ixpt2])
        self.iseqalg = alg

    def set_state2(self, ni, up, te, ti, ng=None, phi=None, tg=None):
        """restart=1 path of ueinit: ni, up are lists of planes per species [iy, ix]; builds yl for the equations that are on."""
        b, c = self.bbb, self.com
        nx, ny = c.nx, c.ny
        shp = (ny + 2, nx + 2)
        f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(shp))
        self.st = dict(ni=[f(a) for a in ni], up=[f(a) for a in up], te=f(te), ti=f(ti),
                       ng=f(ng) if ng is not None else np.zeros(shp), phi=f(phi) if phi is not None else np.zeros(shp),
                       tg=f(tg) if tg is not None else f(ti))
        if int(b.isupgon[0]) == 1:
            self.st["ng"] = self.st["ni"][1].copy()
        yl = np.zeros(b.neq + 2)
        ix_ = self.idx
        def put(idx, val):
            m = idx > 0
            yl[idx[m] - 1] = val[m]
        for s in range(c.nisp):
            put(ix_["idxn"][s], self.st["ni"][s] / b.n0[s])
        for s in range(c.nusp):
            put(ix_["idxu"][s], self.st["up"][s] * (b.mi[s] * b.n0[s]) / b.fnorm[s])
        put(ix_["idxte"], 1.5 * b.nnorm * self.st["te"] / b.ennorm)
        put(ix_["idxti"], 1.5 * b.nnorm * self.st["ti"] / b.ennorm)
        put(ix_["idxg"], self.st["ng"] / b.n0g[0])
        put(ix_["idxtg"], 1.5 * b.n0g[0] * self.st["tg"] / b.ennorm)  # convert.m:123-129 with isflxvar = 0
        put(ix_["idxphi"], self.st["phi"] / b.temp0)
        yl[b.neq] = -1.0
        yl[b.neq + 1] = float(b.nufak) if b.inufaknk == 1 else 0.0
        self.yl = yl
        return yl

    def suscal(self, yl):
        b = self.bbb
        nv = b.numvar
        y = np.abs(yl[: b.neq]).reshape(-1, nv)
        fl = (self.floor_cons / self.norm_cons)[None, :]
        return (1.0 / np.maximum(y, fl)).reshape(-1)

    def drift_geometry(self):
        """Geometrical factors of the curvature and grad-B drifts on the y- and x-faces (bbb/geometry.m:1155-1182; b0 = 1, so
        the s2scal of odesetup.m:1198-1201 is the identity); zero in a slab (mhdgeo < 0)."""
        b = self.bbb
        rm, zm, bb, bphi = self.rz["rm"], self.rz["zm"], self.rz["b"], self.rz["bphi"]
        g = self.geo
        shp = rm[0].shape
        if b.mhdgeo < 0:
            return {k: np.zeros(shp) for k in ("curvrby", "curvrb2", "gradby", "gradb2")}
        s_bphi = np.sign(bphi[1][0, 0]) if bphi[1][0, 0] != 0 else 1.0
        with np.errstate(divide="ignore", invalid="ignore"):
            cossr = -s_bphi * (rm[4] - rm[3]) / ((rm[4] - rm[3]) ** 2 + (zm[4] - zm[3]) ** 2) ** 0.5
            curvrby = -2 * cossr / (rm[4] * bb[4] + rm[3] * bb[3])
            cossp = (rm[4] - rm[2]) / ((rm[4] - rm[2]) ** 2 + (zm[4] - zm[2]) ** 2) ** 0.5
            curvrb2 = -2 * cossp / (rm[4] * bb[4] + rm[2] * bb[2])
            gradby = -0.5 * (bphi[4] / bb[4] ** 3 + bphi[3] / bb[3] ** 3) * (bb[4] - bb[3]) * g["gxc"]
            gradb2 = 0.5 * (1 / bb[4] ** 2 + 1 / bb[2] ** 2) * (bb[4] - bb[2]) * g["gyc"] * np.cos(self.angfx)
        f = lambda a: np.ascontiguousarray(np.nan_to_num(a, nan=0.0, posinf=0.0, neginf=0.0), dtype=np.float64)
        return dict(curvrby=f(curvrby), curvrb2=f(curvrb2), gradby=f(gradby), gradb2=f(gradb2))

    def inputs2(self):
        """name -> float64 array for ue_or2_set: every numeric scalar/array of the namespaces, then the computed data."""
        b, c = self.bbb, self.com
        out = {}
        for ns in (self.aph, c, b):
            for k, v in ns._d.items():
                if isinstance(v, (bool, int, float, np.integer, np.floating)):
                    out[k] = np.array([float(v)])
                elif isinstance(v, np.ndarray) and v.dtype.kind in "fiub" and v.size > 0:
                    out[k] = np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
        nx, ny = c.nx, c.ny
        shp = (ny + 2, nx + 2)
        out["sygytotc"] = np.array([self.sygytotc])
        out["mpe"] = np.array([float(self.rate_tables[0])]); out["mpd"] = np.array([float(self.rate_tables[1])])
        for k, v in zip(("wsveh", "wsveh0", "welms1", "welms2"), self.rate_tables[2]):
            out[k] = np.asarray(v, dtype=float).reshape(-1)
        g = self.geo
        for k in ("vol", "gx", "gy", "gxf", "gyf", "gxc", "gyc", "sx", "sxnp", "sy", "rr", "rrv", "volv", "syv", "dxnog", "dynog", "btot", "rbfbt", "rbfbt2", "lcone", "lconi", "isxptx", "isxpty"):
            out[k] = np.asarray(g[k], dtype=float).reshape(-1)
        out["angfx"] = self.angfx.reshape(-1).astype(float)
        out.update({k: v.reshape(-1) for k, v in self.drift_geometry().items()})
        out["b_c"] = self.rz["b"][0].reshape(-1).astype(float); out["rm_c"] = self.rz["rm"][0].reshape(-1).astype(float)
        out["ixm1"] = self.ixm1.reshape(-1).astype(float); out["ixp1"] = self.ixp1.reshape(-1).astype(float)
        one = np.ones((2,) + shp); zero = np.zeros((2,) + shp)
        for nm in ("fxm", "fx0", "fxp", "fxmy", "fxpy", "fym", "fy0", "fyp", "fymx", "fypx", "fymv", "fy0v", "fypv", "fymxv", "fypxv"):
            if int(c.isnonog) >= 1:
                out[nm] = self.nog[nm].reshape(-1).astype(float)
            else:
                out[nm] = (one if nm in ("fx0", "fy0", "fy0v") else zero).reshape(-1).copy()
        st = getattr(self, "st", None)
        out["ngfix"] = (st["ng"] if st is not None else np.zeros(shp)).reshape(-1).astype(float)
        if st is not None:
            for s in range(c.nisp):
                out["ni%d_init" % (s + 1)] = st["ni"][s].reshape(-1); out["up%d_init" % (s + 1)] = st["up"][s].reshape(-1)
            for k in ("te", "ti", "ng", "tg", "phi"):
                out[k + "_init"] = st[k].reshape(-1)
        for k, v in self.a1.items():
            out[k] = np.asarray(v, dtype=float)
        out["yyf"] = self.geo1d["yyf"]
        out["isixcore"] = self.geo1d["isixcore"].astype(float)
        for k, v in self.bc1.items():
            out[k] = np.asarray(v, dtype=float)
        out["iseqalg"] = self.iseqalg.astype(float)
        out["igyl"] = self.igyl.T.copy().reshape(-1).astype(float)
        for k, v in self.idx.items():
            out[k] = v.reshape(-1).astype(float)
        for k in ("ixpt1", "ixpt2", "iysptrx1", "iysptrx2", "iysptrx", "ixlb", "ixrb", "ixmp", "nisp", "nusp"):
            out[k] = np.array([float(getattr(c, k))])
        return out
