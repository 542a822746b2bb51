#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference tree.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

Writes
  tests/golden/nist.json      inputs of all 27 NIST StRD fits exactly as
                              examples/nist.py builds them (x, y, sigma_y,
                              priors, start), NIST certified values from
                              examples/nist/*.txt, lsqfit's expected strings
                              (the assert_equal arguments) and the chi2/dof, Q,
                              logGBF, iteration counts printed in
                              examples/nist.out; plus values of the reference's
                              own fit functions at the start and certified
                              points (to check functor implementations).
  tests/golden/examples.json  inputs + printed results of examples/simple.py,
                              y-vs-x.py, p-corr.py, x-err.py, empbayes.py,
                              y-noerr.py and numeric pins from
                              tests/test_lsqfit.py.

The reference scripts import gvar and lsqfit, neither of which is installed;
examples/nist.py is therefore executed against two tiny stub modules that only
*record* what the script passes to ``lsqfit.nonlinear_fit``.  No reference code
is copied: only data and printed numbers are extracted.
"""
import json
import os
import re
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------- gvar strings
def parse_gvar(s):
    """'0.253(32)', '238.9(2.7)', '0.0(2.5)e-05', '1 +- 2', '0 ± 4.8e+04' -> (mean, sdev)."""
    s = s.strip()
    m = re.match(r"^(.*?)\s*(\+-|±)\s*(.*)$", s)
    if m:
        return float(m.group(1)), float(m.group(3))
    m = re.match(r"^([-+]?[0-9]*\.?[0-9]*)\(([0-9.]+)\)(e[-+]?[0-9]+)?$", s)
    if not m:
        raise ValueError("cannot parse gvar string %r" % s)
    mean_s, err_s, exp_s = m.group(1), m.group(2), m.group(3)
    scale = float("1" + exp_s) if exp_s else 1.0
    mean = float(mean_s)
    if "." in err_s:
        sdev = float(err_s)
    else:
        ndec = len(mean_s.split(".")[1]) if "." in mean_s else 0
        sdev = float(err_s) * 10.0 ** (-ndec)
    return mean * scale, sdev * scale


class G(object):
    def __init__(self, mean, sdev):
        self.mean, self.sdev = float(mean), float(sdev)

    def __add__(self, o):
        return G(self.mean + float(o), self.sdev)

    __radd__ = __add__


def _gvar(*args):
    if len(args) == 2:
        return G(*args)
    a = args[0]
    if isinstance(a, str):
        return G(*parse_gvar(a))
    return np.array([_gvar(ai) for ai in a], dtype=object)


# ---------------------------------------------------------------- NIST capture
def nist_txt(name):
    """Certified values etc. from examples/nist/<name>.txt."""
    with open(os.path.join(REF, "examples/nist", name + ".txt")) as f:
        lines = f.readlines()
    start1, start2, cert, cert_sd = [], [], [], []
    rss = rsd = None
    for line in lines:
        m = re.match(r"^\s*b\d+\s*=\s*(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s*$", line)
        if m:
            start1.append(float(m.group(1)))
            start2.append(float(m.group(2)))
            cert.append(float(m.group(3)))
            cert_sd.append(float(m.group(4)))
        m = re.match(r"^Residual Sum of Squares:\s*(\S+)", line)
        if m:
            rss = float(m.group(1))
        m = re.match(r"^Residual Standard Deviation:\s*(\S+)", line)
        if m:
            rsd = float(m.group(1))
    return dict(start1=start1, start2=start2, certified=cert,
                certified_sdev=cert_sd, rss=rss, rsd=rsd)


def nist_out():
    """chi2/dof, dof, Q, logGBF, nit per problem from examples/nist.out."""
    with open(os.path.join(REF, "examples/nist.out")) as f:
        text = f.read()
    ans = {}
    for chunk in text.split("=" * 20)[1:]:
        name = chunk.split()[0]
        m = re.search(r"chi2/dof \[dof\] = (\S+) \[(\d+)\]\s+Q = (\S+)\s+logGBF = (\S+)", chunk)
        it = re.search(r"itns/time = (\d+)/", chunk)
        ans[name] = dict(chi2_dof=m.group(1), dof=int(m.group(2)), Q=m.group(3),
                         logGBF=m.group(4), nit=int(it.group(1)))
    return ans


def capture_nist():
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.models import NIST_FORM
    records = []

    gv = types.ModuleType("gvar")
    gv.gvar = _gvar
    lsq = types.ModuleType("lsqfit")

    class _fit(object):
        def __init__(self, prior, data, fcn, p0, tol):
            records.append(dict(prior=prior, data=data, fcn=fcn, p0=p0, tol=tol))
            self.p = None

        def __str__(self):
            return ""

    lsq.nonlinear_fit = _fit
    sys.modules["gvar"], sys.modules["lsqfit"] = gv, lsq
    src = open(os.path.join(REF, "examples/nist.py")).read()
    ns = {"__name__": "nist_ref"}
    exec(compile(src, "nist.py", "exec"), ns)
    expected = []
    ns["assert_equal"] = lambda x, s: expected.append(s)
    ns["print"] = lambda *a, **k: None
    names = [
        'misra1a', 'chwirut2', 'chwirut1', 'lanczos3', 'gauss1', 'gauss2',
        'danwood', 'misra1b', 'kirby2', 'hahn1', 'nelson', 'mgh17', 'lanczos1',
        'lanczos2', 'gauss3', 'misra1c', 'misra1d', 'roszman1', 'enso', 'mgh09',
        'thurber', 'boxbod', 'rat42', 'mgh10', 'eckerle4', 'rat43', 'bennett5',
        ]
    outinfo = nist_out()
    problems = []
    for name in names:
        ns[name]()
        rec, exp = records[-1], expected[-1]
        x, y = rec["data"]
        x = np.asarray(x, dtype=float)
        xrows = x.T if x.ndim == 2 else x[:, None]            # (ny, nx)
        txt = nist_txt(name)
        fcn = rec["fcn"]
        p0 = np.asarray(rec["p0"], dtype=float)
        cert = np.asarray(txt["certified"])
        prob = dict(
            name=name, form=NIST_FORM[name],
            x=xrows.tolist(),
            y=[g.mean for g in y], ysdev=[g.sdev for g in y],
            prior_mean=[g.mean for g in rec["prior"]],
            prior_sdev=[g.sdev for g in rec["prior"]],
            p0=p0.tolist(), tol=rec["tol"],
            expected=exp, out=outinfo[name],
            f_p0=np.asarray(fcn(x, p0), dtype=float).tolist(),
            f_cert=np.asarray(fcn(x, cert), dtype=float).tolist(),
            )
        prob.update(txt)
        assert np.allclose(p0, txt["start2"]), name
        problems.append(prob)
    with open(os.path.join(HERE, "nist.json"), "w") as f:
        json.dump(dict(source="examples/nist.py, examples/nist/*.txt, examples/nist.out "
                               "(gplepage/lsqfit v13.3.1)", problems=problems), f)
    print("nist.json:", len(problems), "problems")


# ---------------------------------------------------------------- other examples
def capture_examples():
    ex = {}
    # examples/simple.py:28-48, simple.out:2-6,21-31
    ex["simple"] = dict(
        model="simple",
        x=[[0.1, 0], [1.0, 0], [0.1, 0], [0.5, 0], [0, 1]],
        ymean=[1.376, 2.010, 1.329, 1.582, 2.0],
        ycov_blocks=[[[0.0047, 0.01], [0.01, 0.056]],
                     [[0.0047, 0.0067], [0.0067, 0.0136]], [[0.25]]],
        prior_mean=[0.5, 0.5], prior_sdev=[0.5, 0.5],
        out=dict(chi2_dof="0.17", dof=5, Q="0.97", logGBF="0.65538", nit=8, svdn=0,
                 p=["0.253(32)", "0.449(65)"],
                 fit=["1.347(46)", "2.02(16)", "1.347(46)", "1.612(82)", "1.78(30)"],
                 # Partial % Errors table (simple.out:26-31): rows y, prior, total; cols a, b/a, b
                 budget=dict(y=[12.75, 16.72, 14.30], prior=[0.92, 1.58, 1.88],
                             total=[12.78, 16.80, 14.42])),
        )
    # examples/y-vs-x.py:69-102, y-vs-x.out
    ycov = [
        [2.1537808808e-09, 8.8161794696e-10, 3.6237356558e-10, 1.4921344875e-10,
         6.1492842463e-11, 2.5353714617e-11, 4.3137593878e-12, 7.3465498888e-13],
        [8.8161794696e-10, 3.6193461816e-10, 1.4921610813e-10, 6.1633547703e-11,
         2.5481570082e-11, 1.0540958082e-11, 1.8059692534e-12, 3.0985581496e-13],
        [3.6237356558e-10, 1.4921610813e-10, 6.1710468826e-11, 2.5572230776e-11,
         1.0608148954e-11, 4.4036448945e-12, 7.6008881270e-13, 1.3146405310e-13],
        [1.4921344875e-10, 6.1633547703e-11, 2.5572230776e-11, 1.0632830128e-11,
         4.4264622187e-12, 1.8443245513e-12, 3.2087725578e-13, 5.5986403288e-14],
        [6.1492842463e-11, 2.5481570082e-11, 1.0608148954e-11, 4.4264622187e-12,
         1.8496194125e-12, 7.7369196122e-13, 1.3576009069e-13, 2.3914810594e-14],
        [2.5353714617e-11, 1.0540958082e-11, 4.4036448945e-12, 1.8443245513e-12,
         7.7369196122e-13, 3.2498644263e-13, 5.7551104112e-14, 1.0244738582e-14],
        [4.3137593878e-12, 1.8059692534e-12, 7.6008881270e-13, 3.2087725578e-13,
         1.3576009069e-13, 5.7551104112e-14, 1.0403917951e-14, 1.8976295583e-15],
        [7.3465498888e-13, 3.0985581496e-13, 1.3146405310e-13, 5.5986403288e-14,
         2.3914810594e-14, 1.0244738582e-14, 1.8976295583e-15, 3.5672355835e-16]]
    ex["y-vs-x"] = dict(
        model="multiexp",
        x=[5., 6., 7., 8., 9., 10., 12., 14.],
        ymean=[4.5022829417e-03, 1.8170543788e-03, 7.3618847843e-04, 2.9872730036e-04,
               1.2128831367e-04, 4.9256559129e-05, 8.1263644483e-06, 1.3415253536e-06],
        ycov=ycov, prior_a="0.5(4)", prior_E_sdev=0.4,
        # p0 of each fit is the previous fit's result for the shared parameters
        # (y-vs-x.py: p0 = fit.pmean), first fit: p0 = None
        out={
            "1": dict(chi2_dof="1.2e+03", dof=8, Q="0", logGBF="-4837.2", svdn=1, nit=11,
                      p=["0.00735(59)", "1.1372(49)"]),
            "2": dict(chi2_dof="2.2", dof=8, Q="0.024", logGBF="111.69", svdn=1, nit=9,
                      p=["0.4024(40)", "0.4471(46)", "0.90104(51)", "1.8282(14)"]),
            "3": dict(chi2_dof="0.63", dof=8, Q="0.76", logGBF="116.29", svdn=1, nit=27,
                      p=["0.4019(40)", "0.406(14)", "0.61(36)",
                         "0.90039(54)", "1.8026(82)", "2.83(19)"]),
            "4": dict(chi2_dof="0.63", dof=8, Q="0.76", logGBF="116.3", svdn=1, nit=6,
                      p=["0.4019(40)", "0.406(14)", "0.61(36)", "0.50(40)",
                         "0.90039(54)", "1.8026(82)", "2.83(19)", "4.00(40)"]),
            },
        ratios=dict(E1_E0="2.0020(87)", E2_E0="3.14(21)", a1_a0="1.011(33)", a2_a0="1.52(89)"),
        )
    # examples/p-corr.py:44-61, p-corr.out  (correlated prior: p1 = 20 p0 + 0.0(1))
    ex["p-corr"] = dict(
        model="mgh09",
        x=[4., 2., 1., 0.5, 0.25, 0.167, 0.125, 0.1, 0.0833, 0.0714, 0.0625],
        y=['0.198(14)', '0.216(15)', '0.184(23)', '0.156(44)', '0.099(49)',
           '0.142(40)', '0.108(32)', '0.065(26)', '0.044(22)', '0.041(19)', '0.044(16)'],
        prior_mean=[0, 0, 0, 0],
        prior_cov=[[1., 20., 0, 0], [20., 400. + 0.01, 0, 0], [0, 0, 1., 0], [0, 0, 0, 1.]],
        out=dict(chi2_dof="0.61", dof=11, Q="0.82", logGBF="19.129", nit=18, svdn=0,
                 p=["0.149(17)", "2.97(34)", "1.23(61)", "0.59(15)"],
                 p1_p0="19.97(67)", p3_p2="0.48(22)", corr_p0_p1="0.9571"),
        )
    # examples/x-err.py:21-43, x-err.out
    ex["x-err"] = dict(
        model="xerr_logistic",
        xprior=['0.73(50)', '2.25(50)', '3.07(50)', '3.62(50)', '4.86(50)',
                '6.41(50)', '6.39(50)', '7.89(50)', '9.32(50)', '9.78(50)',
                '10.83(50)', '11.98(50)', '13.37(50)', '13.84(50)', '14.89(50)'],
        y=['3.85(70)', '5.5(1.7)', '14.0(2.6)', '21.8(3.4)', '47.0(5.2)',
           '79.8(4.6)', '84.9(4.6)', '95.2(2.2)', '97.65(79)', '98.78(55)',
           '99.41(25)', '99.80(12)', '100.127(77)', '100.202(73)', '100.203(71)'],
        bprior=['0(500)', '0(5)', '0(5)', '0(5)'],
        out=dict(chi2_dof="0.35", dof=15, Q="0.99", logGBF="-40.156", nit=13, svdn=0,
                 p=["100.238(60)", "3.5(1.2)", "0.797(87)", "0.77(35)",
                    "1.26(41)", "1.87(34)", "2.84(28)", "3.42(29)", "4.72(32)",
                    "6.45(33)", "6.69(35)", "8.15(36)", "9.30(35)", "9.91(37)",
                    "10.77(37)", "11.70(38)", "13.34(46)", "13.91(48)", "14.88(50)"]),
        )
    # tests/test_lsqfit.py:1887-1901 gammaQ table
    ex["gammaQ"] = _grab_gammaq()
    # tests/test_lsqfit.py:581-589 svdcut pins: y = [1(1), 1(1)] fully correlated?  see test
    with open(os.path.join(HERE, "examples.json"), "w") as f:
        json.dump(dict(source="examples/*.py, examples/*.out, tests/test_lsqfit.py "
                               "(gplepage/lsqfit v13.3.1)", examples=ex), f, indent=1)
    print("examples.json:", sorted(ex))


def _grab_gammaq():
    src = open(os.path.join(REF, "tests/test_lsqfit.py")).read()
    m = re.search(r"def test_gammaQ.*?cases = \[(.*?)\]\s*\n\s*for", src, re.S)
    rows = re.findall(r"\(\s*([-0-9.e+]+)\s*,\s*([-0-9.e+]+)\s*,\s*([-0-9.e+]+)\s*,\s*([-0-9.e+]+)\s*\)", m.group(1))
    return [[float(v) for v in r] for r in rows]


if __name__ == "__main__":
    capture_nist()
    capture_examples()
