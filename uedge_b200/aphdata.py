"""Hydrogen rate tables (DEGAS2 `ehr2.dat`) for com.istabon=10.

Restates `readehr1` (aph/aphread.m:370-468): four blocks wsveh (ionisation),
wsveh0 (recombination), welms1, welms2 (radiation), each mpd=15 density rows of
mpe=60 temperatures written `6(1x,e12.5)` — including Fortran's exponent form
without the letter (``1.20091-102``) — then converted to SI with a 1e-50 floor.
Returned arrays are Fortran-ordered (mpe fastest) flat vectors.
"""
import re

import numpy as np

_TOK = re.compile(r"[-+]?\d\.\d{5}(?:[EeDd][-+]?\d+|[-+]\d+)")


def _val(t):
    t = t.replace("D", "E").replace("d", "e")
    if "E" not in t and "e" not in t:
        m = re.match(r"([-+]?\d\.\d+)([-+]\d+)$", t)
        t = m.group(1) + "E" + m.group(2)
    return float(t)


def read_ehr(path, mpe=60, mpd=15):
    lines = open(path).read().splitlines()
    pos = 0
    blocks = []
    for _ in range(4):
        pos += 1  # block title
        tab = np.zeros((mpd, mpe))
        for jd in range(mpd):
            pos += 1  # "jn = ..." line
            vals = []
            while len(vals) < mpe:
                vals += [_val(t) for t in _TOK.findall(lines[pos])]
                pos += 1
            tab[jd, :] = vals[:mpe]
        blocks.append(tab)
    scale = (1.0e-6, 1.0e-6, 1.0e-7, 1.0e-7)
    return [np.maximum(1.0e-50, t).reshape(-1) * s for t, s in zip(blocks, scale)]
