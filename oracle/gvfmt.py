"""Compare numbers with gvar-formatted golden strings such as ``0.253(32)``.

The reference's goldens (examples/*.out, assert strings in examples/nist.py)
print results with gvar's ``mean(sdev)`` notation.  Rather than re-implement
gvar's formatter, ``agrees`` checks that (mean, sdev) round to the printed
digits.

TEST INFRASTRUCTURE ONLY -- never imported by lsqfit_b200.
"""
import re


def parse(s):
    """'238.9(2.7)' -> (mean, sdev, quantum) ; quantum = unit of the last printed digit."""
    s = s.strip()
    m = re.match(r"^(.*?)\s*(\+-|±)\s*(.*)$", s)
    if m:
        mean_s, err_s = m.group(1), m.group(3)
        mm = re.match(r"^[-+]?[0-9]*\.?([0-9]*)(e[-+]?[0-9]+)?$", err_s)
        q = 10.0 ** (-len(mm.group(1))) * (float("1" + mm.group(2)) if mm.group(2) else 1.0)
        return float(mean_s), float(err_s), q
    m = re.match(r"^([-+]?[0-9]*\.?[0-9]*)\(([0-9.]+)\)(e[-+]?[0-9]+)?$", s)
    if not m:
        raise ValueError("cannot parse gvar string %r" % s)
    mean_s, err_s, exp_s = m.group(1), m.group(2), m.group(3)
    scale = float("1" + exp_s) if exp_s else 1.0
    ndec = len(mean_s.split(".")[1]) if "." in mean_s else 0
    q = 10.0 ** (-ndec)
    mean = float(mean_s)
    sdev = float(err_s) if "." in err_s else float(err_s) * q
    return mean * scale, sdev * scale, q * scale


def agrees(mean, sdev, s, slack=0.51):
    """True if (mean, sdev) print as ``s`` (to within rounding of the last digit)."""
    m, e, q = parse(s)
    return abs(mean - m) <= slack * q and abs(sdev - e) <= slack * q


def agrees_g(value, s, ndigit):
    """True if ``'%.{ndigit}g' % value`` equals golden string ``s`` numerically."""
    return float(("%." + str(ndigit) + "g") % value) == float(s)
