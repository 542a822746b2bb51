"""Forward-mode automatic differentiation on numpy arrays (oracle only).

Stands in for ``gvar.valder`` + GVar operator overloading, which the reference
uses to obtain Jacobians (src/lsqfit/_scipy.py:144-154: ``f(_valder + x)`` then
``fx[i].der``).  A ``Dual`` carries a value array ``v`` of any shape and a
derivative array ``d`` of shape ``v.shape + (n,)``.

TEST INFRASTRUCTURE ONLY -- never imported by lsqfit_b200.
"""
import numpy as np


def _lift(x):
    return x if isinstance(x, Dual) else None


class Dual(object):
    __array_priority__ = 1000.0     # numpy defers to our reflected operators
    __array_ufunc__ = None
    __slots__ = ("v", "d")

    def __init__(self, v, d):
        self.v = np.asarray(v, dtype=float)
        self.d = np.asarray(d, dtype=float)

    # ---- construction ----------------------------------------------------
    @staticmethod
    def variables(p):
        """Independent variables p[0..n) (cf. gvar.valder)."""
        p = np.asarray(p, dtype=float)
        return Dual(p, np.eye(p.size).reshape(p.shape + (p.size,)))

    @property
    def n(self):
        return self.d.shape[-1]

    @property
    def shape(self):
        return self.v.shape

    def __len__(self):
        return len(self.v)

    def __getitem__(self, idx):
        return Dual(self.v[idx], self.d[idx])

    def __iter__(self):
        for i in range(len(self.v)):
            yield self[i]

    # ---- arithmetic ------------------------------------------------------
    def _b(self, o):
        """Broadcast helper: returns (ov, od or None)."""
        if isinstance(o, Dual):
            return o.v, o.d
        return np.asarray(o, dtype=float), None

    def __neg__(self):
        return Dual(-self.v, -self.d)

    def __pos__(self):
        return self

    def __add__(self, o):
        ov, od = self._b(o)
        v = self.v + ov
        d = np.broadcast_to(self.d, v.shape + (self.n,)) if od is None else self.d + od
        return Dual(v, d)

    __radd__ = __add__

    def __sub__(self, o):
        ov, od = self._b(o)
        v = self.v - ov
        d = np.broadcast_to(self.d, v.shape + (self.n,)) if od is None else self.d - od
        return Dual(v, d)

    def __rsub__(self, o):
        return (-self).__add__(o)

    def __mul__(self, o):
        ov, od = self._b(o)
        v = self.v * ov
        d = self.d * np.asarray(ov)[..., None]
        if od is not None:
            d = d + od * self.v[..., None]
        return Dual(v, d)

    __rmul__ = __mul__

    def __truediv__(self, o):
        ov, od = self._b(o)
        v = self.v / ov
        d = self.d / np.asarray(ov)[..., None]
        if od is not None:
            d = d - od * (v / ov)[..., None]
        return Dual(v, d)

    def __rtruediv__(self, o):
        o = np.asarray(o, dtype=float)
        v = o / self.v
        return Dual(v, -self.d * (v / self.v)[..., None])

    def __pow__(self, o):
        if isinstance(o, Dual):
            return exp(log(self) * o)
        o = np.asarray(o, dtype=float)
        v = self.v ** o
        return Dual(v, self.d * (o * self.v ** (o - 1.0))[..., None])

    def __rpow__(self, o):
        # o ** self
        o = np.asarray(o, dtype=float)
        v = o ** self.v
        return Dual(v, self.d * (v * np.log(o))[..., None])


def _unary(f, df):
    def g(x):
        if isinstance(x, Dual):
            return Dual(f(x.v), x.d * df(x.v)[..., None])
        return f(np.asarray(x, dtype=float))
    return g


exp = _unary(np.exp, np.exp)
log = _unary(np.log, lambda v: 1.0 / v)
sqrt = _unary(np.sqrt, lambda v: 0.5 / np.sqrt(v))
sin = _unary(np.sin, np.cos)
cos = _unary(np.cos, lambda v: -np.sin(v))
arctan = _unary(np.arctan, lambda v: 1.0 / (1.0 + v * v))


def concatenate(parts):
    """np.concatenate for a mix of Dual / float 1-d arrays."""
    if not any(isinstance(p, Dual) for p in parts):
        return np.concatenate([np.asarray(p, dtype=float) for p in parts])
    n = next(p.n for p in parts if isinstance(p, Dual))
    vs, ds = [], []
    for p in parts:
        if isinstance(p, Dual):
            vs.append(p.v)
            ds.append(p.d)
        else:
            p = np.asarray(p, dtype=float)
            vs.append(p)
            ds.append(np.zeros(p.shape + (n,)))
    return Dual(np.concatenate(vs), np.concatenate(ds))


def stack(items):
    """Stack scalar Duals / floats into a 1-d Dual."""
    if not any(isinstance(p, Dual) for p in items):
        return np.array([float(p) for p in items])
    n = next(p.n for p in items if isinstance(p, Dual))
    v = np.array([p.v if isinstance(p, Dual) else float(p) for p in items], dtype=float)
    d = np.array([p.d if isinstance(p, Dual) else np.zeros(n) for p in items], dtype=float)
    return Dual(v, d)


def cumsum(x):
    if isinstance(x, Dual):
        return Dual(np.cumsum(x.v, axis=0), np.cumsum(x.d, axis=0))
    return np.cumsum(x, axis=0)


def value(x):
    return x.v if isinstance(x, Dual) else np.asarray(x, dtype=float)


def deriv(x, n):
    return x.d if isinstance(x, Dual) else np.zeros(np.shape(x) + (n,))
