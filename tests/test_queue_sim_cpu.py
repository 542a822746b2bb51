"""tools/queue_sim.py (the schedule simulation DESIGN.md section 5.1c quotes) on the committed evaluation counts of the
10^4 C3 copies (CPU oracle): the simulated makespans respect their lower bounds, longest-first is the best order and
within 5 % of the bound, and the statistics quoted in DESIGN.md are those of the data."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("queue_sim", os.path.join(ROOT, "tools", "queue_sim.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_schedule_bounds_and_quoted_statistics():
    qs = _load()
    nf = np.load(qs.DATA)["nfev"].astype(int)
    assert len(nf) == 10000 and nf.max() == 384 and int(np.median(nf)) == 19
    assert abs(nf.mean() - 23.34) < 0.01 and int(np.percentile(nf, 99)) == 117
    t, teams = 22.8e-6, 592
    bound = 1e3 * max(nf.sum() * t / teams, nf.max() * t)
    natural = qs.list_schedule(nf, range(len(nf)), teams, t)
    lpt = qs.list_schedule(nf, np.argsort(-nf), teams, t)
    spt = qs.list_schedule(nf, np.argsort(nf), teams, t)
    assert bound <= lpt <= natural <= spt
    assert lpt < 1.05 * bound
    total, wave_end, resumed = qs.wave_then_teams(nf, 40, 16e-6)
    assert resumed == int((nf > 40).sum()) and wave_end < total < natural
