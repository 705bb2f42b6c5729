#!/usr/bin/env python
"""Golden vectors for the Rys roots/weights, taken from the REFERENCE's own tables.

Run in the build container only (reads /root/reference, which does not exist on the GPU
box).  Parses jqc/backend/rys/rys_root{1..9}.cu, evaluates ROOT_SMALLX_DATA /
ROOT_LARGEX_DATA / ROOT_RW_DATA exactly the way jqc/backend/rys/rys_roots.cu:29-160 does
(including its two-step Clenshaw and the erf closed form for one root) and writes
tests/golden/rys_ref_samples.json = {nroots: [[x, [roots...], [weights...]], ...]} with
roots sorted ascending.
"""
import json
import math
import os
import re

import numpy as np

REF = "/root/reference/jqc/backend/rys"


def parse(n):
    s = open(f"{REF}/rys_root{n}.cu").read()
    out = {}
    for name, body in re.findall(r"DataType\s+(\w+)\[\]\s*=\s*\{(.*?)\};", s, re.S):
        out[name] = np.array([float(v) for v in re.findall(r"[-+]?\d+\.\d+e[-+]?\d+|[-+]?\d+\.\d+", body)])
    return out


def ref_eval(n, x, d):
    if x < 3e-7:
        sx = d["ROOT_SMALLX_DATA"].reshape(n, 4)
        return [(sx[i, 0] + sx[i, 1] * x, sx[i, 2] + sx[i, 3] * x) for i in range(n)]
    if x > 35 + 5 * n:
        lx = d["ROOT_LARGEX_DATA"].reshape(n, 2)
        t = 0.8862269254527580136 / math.sqrt(x)
        return [(lx[i, 0] / x, lx[i, 1] * t) for i in range(n)]
    if n == 1:
        tt = math.sqrt(x)
        fmt0 = 0.8862269254527580136 / tt * math.erf(tt)
        fmt1 = 0.5 / x * (fmt0 - math.exp(-x))
        return [(fmt1 / fmt0, fmt0)]
    rw = d["ROOT_RW_DATA"].reshape(n, 40, 14, 2)
    it = int(x * 0.4)
    u = (x - it * 2.5) * 0.8 - 1.0
    u2 = 2 * u
    res = []
    for i in range(n):
        pair = []
        for t in range(2):
            a = rw[i, it, :, t]
            c0, c1 = a[13], a[12]
            for m in range(11, 0, -2):
                c2 = a[m] - c1
                c3 = c0 + c1 * u2
                c1 = c2 + c3 * u2
                c0 = a[m - 1] - c3
            pair.append(c0 + c1 * u)
        res.append(tuple(pair))
    return res


def main():
    rng = np.random.default_rng(20261017)
    out = {}
    for n in range(1, 10):
        d = parse(n)
        xs = sorted(set([0.0, 1e-8, 2.9e-7, 1e-3, 0.5, 2.49, 2.51, 3.7, 17.0, 35 + 5 * n - 0.01, 35 + 5 * n + 0.01, 120.0]
                        + [round(float(v), 6) for v in rng.uniform(0, 35 + 5 * n, 24)]))
        rows = []
        for x in xs:
            rw = sorted(ref_eval(n, x, d))
            rows.append([x, [float(r) for r, _ in rw], [float(w) for _, w in rw]])
        out[str(n)] = rows
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rys_ref_samples.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path)


if __name__ == "__main__":
    main()
