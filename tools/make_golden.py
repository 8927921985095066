#!/usr/bin/env python
"""Generate the committed fixtures under tests/golden/ from the reference tree.

Run ONCE in the build container (where /root/reference exists); the GPU box has
no reference tree, so tests read only the .npz files written here.

  d3d_16x8_grid.npz   mesh of builder/test/facets/gridue (same mesh as the one stored in
                      pyexamples/d3dHsmNew/d3dHsm.h5: com/rm, com/zm agree to 3e-14)
  d3dHsm_state.npz    converged state of pyexamples/d3dHsmNew/d3dHsm.h5 (written by UEDGE 8.0.4.1)
                      + the guard-cell rm/zm stored in that file (pins guardc)
  case2_state.npz     restart state builder/test/Forthon_cases/Forthon_case2/h5d3d_ex.16x8
  ehr2_tables.npz     DEGAS2 hydrogen tables of Forthon_case2/ehr2.dat (istabon=10), SI units
  case2_golden.json   numbers printed in Forthon_case2/output_forthon_case2.rtf
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uedge_b200.aphdata import read_ehr  # noqa: E402
from uedge_b200.cases import state_from_h5  # noqa: E402
from uedge_b200.gridue import FIELDS, read_gridue  # noqa: E402
from uedge_b200.h5lite import read_h5  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = "tests/golden"
os.makedirs(OUT, exist_ok=True)

g = read_gridue(os.path.join(REF, "builder/test/facets/gridue"))
np.savez_compressed(os.path.join(OUT, "d3d_16x8_grid.npz"), **{k: g[k] for k in FIELDS},
                    dims=np.array([g["nxm"], g["nym"], g["ixpt1"], g["ixpt2"], g["iysptrx1"]]))

h5 = os.path.join(REF, "pyexamples/d3dHsmNew/d3dHsm.h5")
ni, up, te, ti, ng = state_from_h5(h5)
d = read_h5(h5)
np.savez_compressed(os.path.join(OUT, "d3dHsm_state.npz"), ni=ni, up=up, te=te, ti=ti, ng=ng,
                    rm=d["com/rm"].transpose(2, 1, 0), zm=d["com/zm"].transpose(2, 1, 0))

ni, up, te, ti, ng = state_from_h5(os.path.join(REF, "builder/test/Forthon_cases/Forthon_case2/h5d3d_ex.16x8"))
np.savez_compressed(os.path.join(OUT, "case2_state.npz"), ni=ni, up=up, te=te, ti=ti, ng=ng)

t = read_ehr(os.path.join(REF, "builder/test/Forthon_cases/Forthon_case2/ehr2.dat"))
np.savez_compressed(os.path.join(OUT, "ehr2_tables.npz"), wsveh=t[0], wsveh0=t[1], welms1=t[2], welms2=t[3])

json.dump({
    "source": "builder/test/Forthon_cases/Forthon_case2/output_forthon_case2.rtf",
    "fnrm": [0.7926655291535246, 0.3962737862594912, 0.3718767383113360e-01, 0.7156208214503779e-03,
             0.8680565629223895e-07, 0.1176826431444834e-10],
    "yyc": [-0.00965086, -0.00741261, -0.00258718, 0.0028399, 0.00850614, 0.01416487, 0.01984442, 0.02554944,
            0.03127981, 0.03415233],
    # converged outer-midplane profiles printed at the end of the same file (ni [m^-3], up [m/s], te, ti [eV])
    "midplane_ni": [2.50000000e+19, 2.33721451e+19, 2.00349826e+19, 1.65067380e+19, 1.34076191e+19, 1.09788036e+19,
                    9.18858968e+18, 8.00094888e+18, 7.40016731e+18, 7.40016731e+18],
    "midplane_up": [0.0, -32.62027825, 730.93989383, 3915.89582457, 6102.24137545, 7199.82503778, 7672.80715756,
                    7789.14414326, 7734.00426862, 7734.00426862],
    "midplane_te": [100.0, 97.49739971, 85.78344749, 59.91034579, 43.35906036, 33.5923757, 26.3980251, 19.09156522,
                    8.93259829, 2.0],
    "midplane_ti": [100.0, 95.59396205, 86.51726616, 76.10545898, 65.54926912, 54.82427298, 43.07235412, 29.09750451,
                    11.89098495, 2.0],
}, open(os.path.join(OUT, "case2_golden.json"), "w"), indent=1)
print("wrote fixtures to", OUT)


def rtf_arrays(path):
    """Arrays printed by the 2007 Forthon test decks (`print bbb.ni` ... inside an RTF capture):
    name -> nested-list text -> ndarray in the printed index order [ix, iy(, ifld)]."""
    import ast
    import re
    txt = open(path).read().replace("\\\n", "\n")
    out = {}
    for m in re.finditer(r"^(\w+) = \n(.*?)(?=^\*{5,}|^>>>)", txt, flags=re.S | re.M):
        body = re.sub(r",\s*\]", "]", m.group(2))
        body = re.sub(r"\]\s*\[", "], [", body)
        out[m.group(1)] = np.array(ast.literal_eval(body.strip()))
    return out


# Forthon_case1: slab, 4 unknowns per cell (isngon=0), evolved to steady state by vodpk; final ni, up, te, ti
a = rtf_arrays(os.path.join(REF, "builder/test/Forthon_cases/Forthon_case1/output_forthon_case1.rtf"))
np.savez_compressed(os.path.join(OUT, "case1_state.npz"),
                    ni=a["ni"][:, :, 0].T.copy(), up=a["up"][:, :, 0].T.copy(), te=a["te"].T.copy(), ti=a["ti"].T.copy())
print("case1:", {k: v.shape for k, v in a.items()})


# pyexamples/input_example: 8x4 non-orthogonal single-null mesh, inertial atoms, potential; converged state and the
# reference's own stored pandf1 output + intermediate planes for eleven equation subsets (solution.h5: pytests/<subset>)
g = read_gridue(os.path.join(REF, "pyexamples/input_example/gridue"))
np.savez_compressed(os.path.join(OUT, "inputex_8x4_grid.npz"), **{k: g[k] for k in FIELDS},
                    dims=np.array([g["nxm"], g["nym"], g["ixpt1"], g["ixpt2"], g["iysptrx1"]]))
d = read_h5(os.path.join(REF, "pyexamples/input_example/solution.h5"))
out = {}
for k, v in d.items():
    if k.startswith("bbb/") or k.startswith("pytests/"):
        out[k.replace("/", "__")] = np.asarray(v)
np.savez_compressed(os.path.join(OUT, "inputex_solution.npz"), **out)
print("input_example:", len(out), "arrays")


# jupyter/PyUedge.ipynb: the restart state of its drift + potential case (case_setup.py; cell 17 prints fnrm0 = 2.134077960622300)
d = read_h5(os.path.join(REF, "jupyter/d3d.hdf5"))
np.savez_compressed(os.path.join(OUT, "jupyter_d3d_state.npz"), **{k.split("@")[0]: np.asarray(v) for k, v in d.items()})
print("jupyter:", sorted(k.split("@")[0] for k in d))
