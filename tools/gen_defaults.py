#!/usr/bin/env python
"""Extract variable defaults from the reference's Forthon variable-description
files (bbb/bbb.v, com/com.v, aph/aph.v, api/api.v, grd/grd.v) into
uedge_b200/data/defaults.json.

The .v files are *data declarations* (name, dims, type, /default/), not code.
This script is run once in the build container (where /root/reference exists);
its JSON output is committed so nothing reads /root/reference at run time.

Usage:  python tools/gen_defaults.py [/root/reference]
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
PKGS = ["com", "bbb", "aph", "api", "grd"]

NAME_RE = re.compile(r"^(?P<name>[A-Za-z_][A-Za-z0-9_]*)\s*(?P<dims>\([^)]*\))?(?P<rest>(\s|/).*)?$")
TYPEKW_RE = re.compile(
    r"(?<![A-Za-z0-9_])(_?real|_?integer|_?logical|_?double|character\*\d+|_?complex|real\(Size4\)|function|subroutine)(?![A-Za-z0-9_])"
)


def strip_comment(line):
    # '#' starts a comment unless inside quotes
    out = []
    q = None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "\"'":
            q = ch
            out.append(ch)
        elif ch == "#":
            break
        else:
            out.append(ch)
    return "".join(out)


def parse_params(text):
    """The leading { name = expr } block of compile-time parameters."""
    params = {}
    m = re.search(r"\{(.*?)\}", text, re.S)
    if not m:
        return params
    for line in m.group(1).splitlines():
        line = strip_comment(line).strip()
        if "=" not in line:
            continue
        k, v = line.split("=", 1)
        k = k.strip()
        v = v.strip()
        try:
            params[k] = int(eval(v, {}, dict(params)))
        except Exception:
            pass
    return params


def to_num(tok, typ, env):
    tok = tok.strip()
    if not tok:
        return None
    if tok[0] in "\"'":
        return tok.strip("\"'")
    t = tok.lower()
    if t in (".true.", "true"):
        return True
    if t in (".false.", "false"):
        return False
    t2 = re.sub(r"(?<=[0-9.])d(?=[+-]?[0-9])", "e", t)
    try:
        if "integer" in typ:
            return int(float(t2))
        return float(t2)
    except ValueError:
        try:
            return eval(tok, {}, dict(env))
        except Exception:
            return tok


def parse_default(s, typ, env):
    """'/a, n*b, .../' -> list or scalar"""
    items = []
    # split on commas not inside quotes
    parts = re.split(r",(?=(?:[^\"']*[\"'][^\"']*[\"'])*[^\"']*$)", s)
    for p in parts:
        p = p.strip()
        if not p:
            continue
        m = re.match(r"^([A-Za-z0-9_+\-*() ]+?)\*(?![*])(.+)$", p)
        if m and not p[0] in "\"'":
            cnt_s, val_s = m.group(1), m.group(2)
            try:
                cnt = int(eval(cnt_s, {}, dict(env)))
                v = to_num(val_s, typ, env)
                items.extend([v] * cnt)
                continue
            except Exception:
                pass
        items.append(to_num(p, typ, env))
    if len(items) == 1:
        return items[0]
    return items


def parse_v(path, pkg, out, env):
    text = open(path, errors="replace").read()
    env.update(parse_params(text))
    group = None
    # join continuation: a default may span lines until closing '/'
    lines = text.splitlines()
    i = 0
    while i < len(lines):
        raw = lines[i]
        i += 1
        if raw.startswith("*****"):
            g = raw.strip("* \t").split(":")[0].split()[0] if raw.strip("* \t") else None
            group = g
            continue
        line = strip_comment(raw).rstrip()
        if not line.strip() or raw[:1] in " \t{}":
            continue
        m = NAME_RE.match(line)
        if not m:
            continue
        name = m.group("name")
        dims = m.group("dims")
        rest = m.group("rest") or ""
        rest_nb0 = re.sub(r"\[[^\]]*\]", "", rest)
        # the type keyword must appear outside the /default/ segment
        outside = re.sub(r"/[^/]*/", " ", rest_nb0)
        tm = TYPEKW_RE.search(outside)
        if tm:
            typ = tm.group(1)
        elif "/" in rest_nb0:
            typ = "real"
        else:
            continue
        if typ in ("function", "subroutine"):
            continue
        default = None
        if "/" in rest:
            seg = rest[rest.index("/") + 1:]
            # units like [1/s] may precede the default: drop bracketed text
            rest_nb = re.sub(r"\[[^\]]*\]", "", rest)
            if "/" in rest_nb:
                seg = rest_nb[rest_nb.index("/") + 1:]
                while "/" not in seg and i < len(lines):
                    nxt = strip_comment(lines[i]).strip()
                    i += 1
                    seg += " " + nxt
                if "/" in seg:
                    seg = seg[: seg.index("/")]
                    default = parse_default(seg, typ, env)
        rec = {"pkg": pkg, "group": group, "type": typ}
        if dims:
            rec["dims"] = dims.strip("()")
        if default is not None:
            rec["default"] = default
        out[f"{pkg}.{name}"] = rec


def main():
    out = {}
    env = {}
    for pkg in PKGS:
        p = os.path.join(REF, pkg, pkg + ".v")
        if os.path.exists(p):
            parse_v(p, pkg, out, env)
    dst = os.path.join(os.path.dirname(__file__), "..", "uedge_b200", "data", "defaults.json")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as f:
        json.dump({"params": env, "vars": out}, f, indent=0, sort_keys=True)
    print(f"wrote {len(out)} variables, {len(env)} params -> {dst}")


if __name__ == "__main__":
    main()
