"""Comparison of one library's field planes and pandf1 output with the reference's OWN stored vectors
(pyexamples/input_example/solution.h5, groups pytests/<subset>, fixture tests/golden/inputex_solution.npz).
Used for the oracle (tests/test_oracle2_golden.py) and, unchanged, for the CUDA general path (tests/test_gpu_general.py)."""
import numpy as np

TOL = 5.0e-9  # of the plane's largest value: the noise floor of the reference's -Ofast stencil geometry is 1e-11..1e-10


def planes_to_compare(o, gold, c):
    b = c.bbb
    isu = c.isu
    out = []  # (name, ours, gold[iy, ix])
    T = lambda a: a.T
    for nm in ("fnix", "fniy", "visx", "visy", "hcxij", "hcyij"):
        for s in range(2):
            out.append((nm + str(s + 1), o.plane(nm + str(s + 1)), T(gold[nm][:, :, s])))
    for nm in ("hcxe", "hcye", "hcxi", "hcyi", "conxe", "conye", "conxi", "conyi", "floxe", "floye", "floxi", "floyi", "conxg", "conyg", "floxg", "floyg"):
        out.append((nm, o.plane(nm), T(gold[nm])))
    for nm in ("fngx", "fngy"):
        out.append((nm, o.plane(nm), T(gold[nm][:, :, 0])))
    for s in range(2):
        if isu[s]:
            for nm in ("fmix", "fmiy"):
                out.append((nm + str(s + 1), o.plane(nm + str(s + 1)), T(gold[nm][:, :, s])))
    if any(isu):  # scratch planes of the momentum equation hold the last species evaluated
        for nm in ("conx", "cony", "flox", "floy"):
            out.append((nm, o.plane(nm), T(gold[nm])))
    if int(b.isteon):
        out += [("feex", o.plane("feex"), T(gold["feex"])), ("feey", o.plane("feey"), T(gold["feey"]))]
    if int(b.istion):
        out += [("feix", o.plane("feix"), T(gold["feix"])), ("feiy", o.plane("feiy"), T(gold["feiy"]))]
    return out


def check_against_reference(o, c, gold, subset, yl):
    """o: a bound library (Lib2).  Evaluates pandf1(yl) twice (the module state, e.g. upi at ix = nx+1 set by bouncon, is then
    that of a running code) and asserts every stored plane, the residual planes and - for `te` - the stored yldot and fnrm."""
    o.pandf1(yl)
    f = o.pandf1(yl)
    for name, ours, g in planes_to_compare(o, gold, c):
        scale = np.abs(g).max()
        if scale == 0:  # e.g. the conductivities of the atoms: identically zero in the reference too
            assert np.abs(ours).max() == 0, name
            continue
        err = np.abs(ours - g).max() / scale
        assert err <= TOL, "%s/%s: %.3g" % (subset, name, err)
    # residual planes = divergences of the fluxes above: compare on the flux scale
    flux = dict(resco=max(np.abs(gold["fnix"]).max(), np.abs(gold["fniy"]).max()), resmo=max(np.abs(gold["fmix"]).max(), np.abs(gold["fmiy"]).max()),
                resee=max(np.abs(gold["feex"]).max(), np.abs(gold["feey"]).max()), resei=max(np.abs(gold["feix"]).max(), np.abs(gold["feiy"]).max()))
    for s in range(2):
        if c.isn[s]:
            assert np.abs(o.plane("resco%d" % (s + 1)) - gold["resco"][:, :, s].T).max() <= TOL * flux["resco"], subset
        if c.isu[s]:
            assert np.abs(o.plane("resmo%d" % (s + 1)) - gold["resmo"][:, :, s].T).max() <= 2 * TOL * flux["resmo"], subset
    if int(c.bbb.isteon):
        assert np.abs(o.plane("resee") - gold["resee"].T).max() <= TOL * flux["resee"], subset
    if int(c.bbb.istion):
        assert np.abs(o.plane("resei") - gold["resei"].T).max() <= TOL * flux["resei"], subset
    if int(c.bbb.isphion):  # resphi = factor * (sum of currents): compare on the scale of the currents
        fac = c.bbb.nurlxp * c.bbb.dx0 ** 2 / c.bbb.sigbar0
        cur = max(np.abs(o.plane("fqx")).max(), np.abs(o.plane("fqy")).max())
        d = np.abs(o.plane("resphi") - gold["resphi"].T) / fac
        assert d.max() <= TOL * cur, "%s resphi %.3g of %.3g A" % (subset, d.max(), cur)
        if subset == "phi":  # away from the large parallel currents (rows 2, 3) the stored residual itself is reproduced
            a, g = o.plane("resphi")[2:4, 1:9], gold["resphi"].T[2:4, 1:9]
            assert np.corrcoef(a.ravel(), g.ravel())[0, 1] > 0.95
    gy = gold["yldot"][: c.bbb.neq]
    if subset == "te":  # the one subset whose stored output vector is not a converged ~0: 43 non-zero rows up to 2.2e4
        assert np.abs(gy).max() > 2.0e4 and np.count_nonzero(gy) >= 40
        assert np.abs(f - gy).max() <= 2.0e-8 * np.abs(gy).max()
        assert abs(np.sqrt(np.sum(f * f)) - float(gold["fnrm"])) <= 1e-8 * float(gold["fnrm"])  # fnrm as stored (sfscal = 1)
    return f
