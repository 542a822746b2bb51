#!/usr/bin/env python
"""tests/golden/spline.json from the reference's examples/spline.json (data means + covariance as stored by gvar.dump)
and examples/spline.out (the printed fit).  Run in the build container, where /root/reference exists:
    python tests/golden/make_spline_golden.py
"""
import json
import os
import re

REF = "/root/reference/examples"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    d = json.load(open(os.path.join(REF, "spline.json")))
    assert d["keys"] == "['A', 'B', 'C']"
    idx, cov = d["bcovs"][1]
    assert idx == list(range(12)) and d["bcovs"][0] == [[], []]
    out = open(os.path.join(REF, "spline.out")).read()
    m = re.search(r"chi2/dof \[dof\] = (\S+) \[(\d+)\]\s+Q = (\S+)\s+logGBF = (\S+)", out)
    rows = re.findall(r"^\s+(?:mknot|fknot|c)?\s*\d\s+(-?[\d.]+ \(\d[\d.]*\))\s+\[", out, flags=re.M)
    assert len(rows) == 13, rows
    nit = int(re.search(r"itns/time = (\d+)/", out).group(1))
    golden = dict(
        source="examples/spline.json, examples/spline.out, examples/spline.py:50-80",
        # examples/spline.py:69-73: (ainv, am) per data set; m = am * ainv
        param=dict(A=[10.0, [0.1, 0.3, 0.5, 0.7, 0.9]], B=[5.0, [0.3, 0.5, 0.7, 0.9]], C=[2.5, [0.5, 0.7, 0.9]]),
        ymean=d["means"], ycov=cov,
        # examples/spline.py:59-65
        prior_mean=[1.0, 1.5, 3.0, 9.0, 0.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0],
        prior_sdev=[0.01, 0.5, 1.0, 0.01, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0],
        out=dict(chi2_dof=m.group(1), dof=int(m.group(2)), Q=m.group(3), logGBF=m.group(4), nit=nit,
                 params=[r.replace(" ", "") for r in rows]))
    with open(os.path.join(HERE, "spline.json"), "w") as f:
        json.dump(golden, f, indent=1)
    print("wrote spline.json:", golden["out"])


if __name__ == "__main__":
    main()
