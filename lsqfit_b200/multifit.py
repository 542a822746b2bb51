"""Device path for simultaneous (MultiFitter) fits.

``lsqfit.MultiFitter`` hands ``nonlinear_fit`` a Python closure built from its models
(``_multifitfcn(flatmodels)``, reference src/lsqfit/_extras.py:1018-1028, 1816-1829): parameters arrive as a
dictionary, every model returns the values of its own data set.  A closure cannot run on the device, but a list of
models that are all device-backed can be mapped onto ONE composite device functor: here the shared-energy
correlator family (``multiexp_shared2/3``: M data sets, each  sum_k a^(m)_k exp(-E_k t), common energies E).

``SharedExpModel`` implements the reference's model interface (``datatag``, ``ncg``, ``fitfcn``, ``builddata``,
``buildprior``; src/lsqfit/_extras.py:518-640) with numpy, so it also works with the CPU fitters; ``composite``
turns the closure back into (functor, x rows, parameter permutation) for the ``_build_chiv_chivw`` hook.
"""
import collections

import numpy as np

from .functors import Functor


class SharedExpModel(object):
    """One correlator  G(t) = sum_k p[a][k] exp(-p[E][k] t)  of a simultaneous fit; ``a`` and ``E`` are the keys of
    its amplitudes and of the (shared) energies in the parameter dictionary."""

    def __init__(self, datatag, t, a, E, ncg=1, exp=np.exp):
        """``exp``: the exponential used by the CPU evaluation (``gvar.exp`` when a CPU fitter differentiates the
        model with GVars; the device path never calls it)."""
        self.datatag, self.t, self.a, self.E, self.ncg = datatag, np.asarray(t, dtype=float), a, E, ncg
        self.exp = exp

    def fitfcn(self, p):                                     # _extras.py:552-565
        a, E = p[self.a], p[self.E]
        ans = 0.0
        for k in range(len(a)):
            ans = ans + a[k] * self.exp(-E[k] * self.t)
        return ans

    def builddata(self, data):                               # _extras.py:595-606
        return data[self.datatag]

    def buildprior(self, prior, mopt=None):                  # _extras.py:608-640
        return collections.OrderedDict((k, prior[k]) for k in (self.a, self.E))


def _slices(buf):
    """key -> slice of a flat buffer, for a gvar.BufferDict (``.slice(k)``) or anything with ``.slices``."""
    if hasattr(buf, "slices"):
        return collections.OrderedDict((k, buf.slices[k][0] if isinstance(buf.slices[k], tuple) else buf.slices[k])
                                       for k in getattr(buf, "keys_", buf.slices.keys()))
    return collections.OrderedDict((k, buf.slice(k)) for k in buf.keys())


def composite(multifcn, po, yo):
    """(Functor, x rows [ny, 2], pperm) for a ``_multifitfcn`` whose models are all ``SharedExpModel`` with one common
    energy key, or None.  ``po`` / ``yo`` are the parameter / data buffers of the reference's flatfcn
    (src/lsqfit/__init__.py:2037-2042); ``pperm[k]`` is the position in lsqfit's flat parameter buffer of device
    parameter k (device order: amplitudes of data set 0, 1, ..., then the energies)."""
    models = getattr(multifcn, "flatmodels", None)
    if not models or not all(isinstance(m, SharedExpModel) for m in models):
        return None
    M = len(models)
    if M not in (2, 3) or len(set(m.E for m in models)) != 1 or len(set(m.a for m in models)) != M:
        return None
    ps, ys = _slices(po), _slices(yo)
    def idx(s):
        return np.arange(s.start, s.stop) if isinstance(s, slice) else np.atleast_1d(np.asarray(s))
    E = idx(ps[models[0].E])
    K = E.size
    amps = [idx(ps[m.a]) for m in models]
    if any(a.size != K for a in amps) or sum(a.size for a in amps) + K != sum(idx(s).size for s in ps.values()):
        return None                                          # parameters the models do not use: no device mapping
    by_tag = dict((m.datatag, (i, m)) for i, m in enumerate(models))
    rows = []
    for tag, s in ys.items():
        if tag not in by_tag:
            return None
        i, m = by_tag[tag]
        if idx(s).size != m.t.size:
            return None
        rows.append(np.stack([m.t, np.full(m.t.size, float(i))], axis=1))
    x = np.concatenate(rows, axis=0)
    pperm = np.concatenate(amps + [E]).astype(np.int64)
    return Functor("multiexp_shared%d" % M), x, pperm
