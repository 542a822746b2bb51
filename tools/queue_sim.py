"""Queue simulation behind DESIGN.md section 5.1c: what a batch of 10^4 C3 fits costs under different schedules, from the
evaluation count of every copy (CPU oracle, profiles/c3_nfev_oracle.npz; regenerated with --regen: ~30 s on 8 cores) and the
measured trial / pass times of the kernels.  Pure numpy: runs anywhere.

  python tools/queue_sim.py [--regen]
"""
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "profiles", "c3_nfev_oracle.npz")
GHZ = 1.965e9


def regen():
    sys.path.insert(0, ROOT)
    import multiprocessing as mp
    import bench
    configs, cfg, pdf = bench._oracle_setup()
    means = configs.bootstrap_means(cfg, 10000, cfg["seed"])
    with mp.get_context("fork").Pool(os.cpu_count() or 1, initializer=bench._oracle_init) as pool:
        nit = pool.map(bench._oracle_fit, [(m, "jac") for m in means], chunksize=50)
    np.savez_compressed(DATA, nfev=np.asarray(nit, dtype=np.int16))


def list_schedule(nf, order, teams, t_trial):
    """makespan (ms) of a work queue handed out in `order` to `teams` workers, a fit costing nfev * t_trial"""
    h = [0.0] * teams
    heapq.heapify(h)
    end = 0.0
    for i in order:
        s = heapq.heappop(h)
        e = s + nf[i] * t_trial
        end = max(end, e)
        heapq.heappush(h, e)
    return 1e3 * end


def wave_then_teams(nf, cap, t_team, nteams=592, ncta=148, slots=32, passes_per_eval=1.6):
    """wave kernel (CTA-wide passes: 35 k cycles + 15.5 k per chunk of 8 evaluations; a fit gets an evaluation every
    1.6 passes) with a per-fit evaluation cap, the capped fits resumed afterwards by `nteams` teams at t_team per trial.
    Returns (total ms, ms at which the wave kernel ends, resumed fits)."""
    q = list(nf[::-1])
    resume, done = [], 0.0
    h = [(0.0, i) for i in range(ncta)]
    heapq.heapify(h)
    live = [[] for _ in range(ncta)]
    free = {}
    while h:
        t, i = heapq.heappop(h)
        sl = live[i]
        while len(sl) < slots and q:
            sl.append([float(q.pop()), 0.0])
        if not sl:
            free[i] = t
            continue
        tp = (35e3 + 15.5e3 * np.ceil(len(sl) / passes_per_eval / 8)) / GHZ
        keep = []
        for s in sl:
            s[0] -= 1 / passes_per_eval
            s[1] += 1 / passes_per_eval
            if s[0] <= 1e-9:
                done = max(done, t + tp)
            elif s[1] >= cap:
                resume.append(s[0])
            else:
                keep.append(s)
        live[i] = keep
        heapq.heappush(h, (t + tp, i))
    wave_end = max(free.values())
    th = [wave_end + 5e-6] * nteams
    heapq.heapify(th)
    for r in resume:
        s = heapq.heappop(th)
        e = s + (r + 1) * t_team
        done = max(done, e)
        heapq.heappush(th, e)
    return float(round(1e3 * done, 2)), float(round(1e3 * wave_end, 2)), len(resume)


def main():
    if "--regen" in sys.argv or not os.path.exists(DATA):
        regen()
    nf = np.load(DATA)["nfev"].astype(int)
    n = len(nf)
    print("copies %d: evaluations mean %.1f median %d 90%% %d 99%% %d max %d" % (
        n, nf.mean(), np.median(nf), np.percentile(nf, 90), np.percentile(nf, 99), nf.max()))
    t4 = 22.8e-6                       # trial point of a four-warp team on a loaded SM (13.2 us on an idle one)
    print("four-warp teams (592 in flight), %.1f us per trial:" % (1e6 * t4))
    print("  lower bounds: work / teams %.2f ms, longest fit alone %.2f ms" % (1e3 * nf.sum() * t4 / 592, 1e3 * nf.max() * t4))
    print("  input order %.2f ms, longest first %.2f ms, shortest first %.2f ms" % (
        list_schedule(nf, range(n), 592, t4), list_schedule(nf, np.argsort(-nf), 592, t4), list_schedule(nf, np.argsort(nf), 592, t4)))
    print("wave kernel with an evaluation cap, capped fits resumed by four-warp teams (ms total, ms wave, resumed fits):")
    for cap in (24, 40, 64, 10 ** 9):
        print("  cap %s:" % ("none" if cap > 10 ** 6 else cap),
              "  ".join("%4.1f us/trial -> %s" % (1e6 * tt, wave_then_teams(nf, cap, tt)) for tt in (13.5e-6, 16e-6, 20e-6)))


if __name__ == "__main__":
    main()
