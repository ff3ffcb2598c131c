#!/usr/bin/env python3
"""Generate tests/golden/tableaux.json from the reference source.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):  python tests/golden/make_tableaux_golden.py

The rktp64 decimals are parsed out of /root/reference/src/rk.rs:313-429 and
converted with Python's correctly-rounded float() (Rust's literal parser is
correctly rounded too); the parametric families (rk.rs:79-98, 116-203) and the
small fixed tableaux are evaluated with the same IEEE-754 double expressions the
reference uses.  Every entry is stored as a C99 hex float so the comparison in
the tests is bit-exact.
"""
import json
import os
import re

REF = "/root/reference/src/rk.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tableaux.json")


def hexm(rows):
    return [[float(x).hex() for x in r] for r in rows]


def hexv(v):
    return [float(x).hex() for x in v]


def pack(order, oe, oi, a, b, b2, c, bi):
    return dict(S=len(b), I=len(bi[0]), order=order, order_embedded=oe,
                order_interpolant=oi, a=hexm(a), b=hexv(b), b2=hexv(b2), c=hexv(c), bi=hexm(bi))


def generic_order_2(c2):
    a = [[0., 0.], [c2, 0.]]
    b = [1. - 1. / (2. * c2), 1. / (2. * c2)]
    return pack(2, 1, 1, a, b, [1., 0.], [0., c2], [[0., b[0]], [0., b[1]]])


def generic_order_3(c2, c3):  # case I only (all named tableaux use b3 = None)
    a = [[0., 0., 0.], [c2, 0., 0.],
         [(c3 * (3. * c2 * (1. - c2) - c3)) / (c2 * (2. - 3. * c2)),
          (c3 * (c3 - c2)) / (c2 * (2. - 3. * c2)), 0.]]
    b = [1. - (1. * c2 + 3. * c3 - 2.) / (6. * c2 * c3),
         (3. * c3 - 2.) / (6. * c2 * (c3 - c2)),
         (2. - 3. * c2) / (6. * c3 * (c3 - c2))]
    b2 = [1. - 1. / (2. * c2), 1. / (2. * c2), 0.]
    return pack(2, 1, 1, a, b, b2, [0., c2, c3], [[0., x] for x in b])


def rktp64_from_source():
    src = open(REF).read()
    body = src[src.index("pub fn rktp64()"):]
    body = body[:body.index("ButcherTableu {")]
    num = r"-?\d+\.\d*"
    a = [[0.0] * 7 for _ in range(7)]
    for m in re.finditer(r"a\[(\d)\]\[0\.\.(\d)\]\.copy_from_slice\(&\[(.*?)\]\)", body, re.S):
        i, n = int(m.group(1)), int(m.group(2))
        vals = [float(x) for x in re.findall(num, m.group(3))]
        assert len(vals) == n
        a[i][:n] = vals

    def vec(name):
        m = re.search(r"let %s = \[(.*?)\];" % name, body, re.S)
        return [float(x) for x in re.findall(num, m.group(1))]

    b, b2, c = vec("b"), vec("b2"), vec("c")
    assert len(b) == len(b2) == len(c) == 7
    m = re.search(r"let bi = \[(.*)\];", body, re.S)
    rows = re.findall(r"\[([^\[\]]*)\]", m.group(1))
    bi = [[float(x) for x in re.findall(num, r)] for r in rows]
    assert len(bi) == 7 and all(len(r) == 5 for r in bi)
    return pack(6, 4, 4, a, b, b2, c, bi)


def main():
    t = {}
    t["euler"] = pack(1, 0, 1, [[0.]], [1.], [0.], [0.], [[0., 1.]])
    t["midpoint"] = generic_order_2(0.5)
    t["heun2"] = generic_order_2(1.)
    t["ralston2"] = generic_order_2(2. / 3.)
    t["kutta3"] = generic_order_3(0.5, 1.)
    t["heun3"] = generic_order_3(1. / 3., 2. / 3.)
    t["ralston3"] = generic_order_3(1. / 2., 3. / 4.)
    t["wray3"] = generic_order_3(8. / 15., 2. / 3.)
    t["ssp3"] = generic_order_3(1.0, 0.5)
    b = [1. / 6., 1. / 3., 1. / 3., 1. / 6.]
    t["rk4"] = pack(4, 2, 1,
                    [[0., 0., 0., 0.], [0.5, 0., 0., 0.], [0., 0.5, 0., 0.], [0., 0., 1., 0.]],
                    b, [0., 1., 0., 0.], [0., 0.5, 0.5, 1.], [[0., x] for x in b])
    b = [1. / 8., 3. / 8., 3. / 8., 1. / 8.]
    t["three_eights"] = pack(4, 2, 1,
                             [[0., 0., 0., 0.], [1. / 3., 0., 0., 0.], [-1. / 3., 1., 0., 0.],
                              [1., -1., 1., 0.]],
                             b, [-0.5, 1.5, 0., 0.], [0., 1. / 3., 2. / 3., 1.],
                             [[0., x] for x in b])
    t["rk43"] = pack(4, 3, 3,
                     [[0., 0., 0., 0., 0.], [0.5, 0., 0., 0., 0.], [0., 0.5, 0., 0., 0.],
                      [0., 0., 1., 0., 0.], [5. / 32., 7. / 32., 13. / 32., -1. / 32., 0.]],
                     [1. / 6., 1. / 3., 1. / 3., 1. / 6., 0.],
                     [-1. / 2., 7. / 3., 7. / 3., 13. / 6., -16. / 3.],
                     [0., 0.5, 0.5, 1., 0.75],
                     [[0., 1., -1.5, 2. / 3.], [0., 0., 1., -2. / 3.], [0., 0., 1., -2. / 3.],
                      [0., 0., -0.5, 2. / 3.], [0., 0., 0., 0.]])
    t["rktp64"] = rktp64_from_source()
    with open(OUT, "w") as f:
        json.dump(dict(source="/root/reference/src/rk.rs:55-430", tableaux=t), f, indent=1)
    print("wrote", OUT, "with", len(t), "tableaux")


if __name__ == "__main__":
    main()
