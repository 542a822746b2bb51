#!/usr/bin/env python
"""GPU check of the team kernel (2 / 4 warps per fit) against the one-warp kernel: same fits on the
same inputs (results must agree to rounding), plus timings at several batch sizes.

    python tools/team_check.py [--quick] > gpurun_out/team_check.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import lsqfit_b200 as lb  # noqa: E402
from lsqfit_b200 import configs  # noqa: E402


def problem(K, ny=64, kind="dense", svdcut=1e-12, seed=5):
    cfg = configs.correlator(K, ny=ny, dt=8.0 / ny)
    npar = cfg["np"]
    N = ny + npar
    ycov = cfg["ycov"].copy()
    if kind == "diag":
        ycov = np.diag(np.diag(ycov))
    elif kind == "mixed":                       # two correlated blocks + uncorrelated rows in between
        m = np.zeros_like(ycov)
        a, b = ny // 3, 2 * ny // 3
        m[:a, :a] = ycov[:a, :a]
        m[b:, b:] = ycov[b:, b:]
        m[np.arange(a, b), np.arange(a, b)] = np.diag(ycov)[a:b]
        ycov = m
    full = np.zeros((N, N))
    full[:ny, :ny] = ycov
    full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
    mean0 = np.concatenate([cfg["f"], cfg["prior_mean"]])
    pdf = lb.PDF(mean0, full, svdcut=svdcut)
    cfg["ycov"] = ycov
    return cfg, pdf


def run(plan, means, p0, team, tol, maxit, reps=0, **kw):
    os.environ["B200LM_TEAM"] = str(team)
    out = plan.fit_batch(means, p0, tol=tol, maxit=maxit, **kw)
    torch.cuda.synchronize()
    ms = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(reps):
            e0.record()
            plan.fit_batch(means, p0, tol=tol, maxit=maxit, out=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
    stats = plan.last_stats()
    return out.numpy(), ms, stats


def compare(a, b):
    ok = (a["status"] > 0) & (b["status"] > 0)
    sd = np.sqrt(np.einsum("bii->bi", a["cov"][ok]))
    r = dict(n=int(ok.sum()), status_equal=float((a["status"] == b["status"]).mean()),
             nit_equal=float((a["nit"] == b["nit"]).mean()),
             dp_sd=float(np.max(np.abs(a["x"][ok] - b["x"][ok]) / sd)) if ok.any() else None,
             dp_sd_q99=float(np.quantile(np.max(np.abs(a["x"][ok] - b["x"][ok]) / sd, axis=1), 0.99)) if ok.any() else None,
             dchi2=float(np.max(np.abs(a["chi2"][ok] - b["chi2"][ok]) / a["chi2"][ok])) if ok.any() else None,
             dcov=float(np.max(np.abs(a["cov"][ok] - b["cov"][ok]) / (sd[:, :, None] * sd[:, None, :]))) if ok.any() else None)
    if "f" in a and a["f"] is not None and "f" in b:
        r["df"] = float(np.max(np.abs(a["f"][ok] - b["f"][ok])))
        r["dJ_rel"] = float(np.max(np.abs(a["J"][ok] - b["J"][ok])) / np.max(np.abs(a["J"][ok])))
    return r


def main():
    quick = "--quick" in sys.argv
    rep = dict(cases=[], timing=[])
    # ---- correctness: same fits through the three kernels ----
    cases = [(8, 64, "dense", 1e-12), (3, 64, "dense", 1e-12), (5, 64, "dense", 1e-12), (2, 64, "dense", 1e-12),
             (4, 100, "dense", 1e-12), (4, 100, "dense", -1e-3), (3, 40, "diag", 1e-12), (4, 150, "diag", 1e-12),
             (3, 90, "mixed", 1e-12), (8, 200, "mixed", 1e-12), (6, 30, "dense", 1e-12)]
    for K, ny, kind, cut in cases:
        cfg, pdf = problem(K, ny, kind, cut)
        npar = cfg["np"]
        B = 600
        means = configs.bootstrap_means(cfg, B, seed=17, cov=pdf.cov[:ny, :ny])
        plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts)
        tol = (1e-10, 1e-12, 1e-12)
        base, _, st1 = run(plan, means, cfg["prior_mean"], 1, tol, 600, want_fJ=True)
        for team in (2, 4):
            t0 = time.time()
            o, _, st = run(plan, means, cfg["prior_mean"], team, tol, 600, want_fJ=True)
            r = compare(base, o)
            r.update(K=K, ny=ny, kind=kind, svdcut=cut, team=team, nchiv=int(plan.nchiv), stats1=st1, stats=st,
                     conv=float((o["status"] > 0).mean()), sec=time.time() - t0)
            rep["cases"].append(r)
            print(json.dumps(r), file=sys.stderr, flush=True)
        # polish + ill-conditioned covariance path (mode 1): tight tolerance
        for team in (2, 4):
            a, _, _ = run(plan, means[:100], cfg["prior_mean"], 1, (1e-15, 0, 0), 2000, polish=100)
            o, _, _ = run(plan, means[:100], cfg["prior_mean"], team, (1e-15, 0, 0), 2000, polish=100)
            r = compare(a, o)
            r.update(K=K, ny=ny, kind=kind, svdcut=cut, team=team, polish=True)
            rep["cases"].append(r)
            print(json.dumps(r), file=sys.stderr, flush=True)
        plan.close()
    # ---- timing: C3 and C4 shapes ----
    for K, sizes in ((8, (1000, 10000, 40000, 160000)), (3, (10000, 200000))):
        cfg = configs.c3() if K == 8 else configs.c4()
        ny, npar = cfg["ny"], cfg["np"]
        N = ny + npar
        full = np.zeros((N, N))
        full[:ny, :ny] = cfg["ycov"]
        full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
        pdf = lb.PDF(np.concatenate([cfg["f"], cfg["prior_mean"]]), full, svdcut=cfg["svdcut"])
        plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts)
        for B in sizes:
            if quick and B > 40000:
                continue
            means = torch.as_tensor(configs.bootstrap_means(cfg, B, cfg["seed"], cov=pdf.cov[:ny, :ny],
                                                            vary_prior=(K == 8))).cuda()
            p0 = torch.as_tensor(cfg["p0"]).cuda()
            for team in (1, 2, 4):
                for teams_env in ([None] if team == 1 else [None, 4, 5] if team == 4 else [None, 8]):
                    if teams_env is None:
                        os.environ.pop("B200LM_TEAMS", None)
                    else:
                        os.environ["B200LM_TEAMS"] = str(teams_env)
                    o, ms, st = run(plan, means, p0, team, cfg["tol"], cfg["maxit"], reps=5)
                    r = dict(K=K, B=B, team=team, teams_per_cta=teams_env, ms=ms, fits_per_s=B / ms * 1e3,
                             nfev=st[0] / B, nfac=st[2] / B, conv=float((o["status"] > 0).mean()),
                             max_nit=int(o["nit"].max()))
                    rep["timing"].append(r)
                    print(json.dumps(r), file=sys.stderr, flush=True)
            os.environ.pop("B200LM_TEAMS", None)
        plan.close()
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
