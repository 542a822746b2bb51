"""Parity of the CUDA path with the CPU oracle AT THE BASELINE CONFIGS' sizes and settings:
C3 and C4 on 2000 copies at the settings bench.py runs, C2 on 200 perturbed starts per NIST problem, the
dense-block chiv / Jacobian element by element, the C4 bias, and config 5 at its full size against a committed
oracle fixture.  Needs a GPU (pytest -m gpu); the oracle runs in a process pool on the host cores.
"""
import os

import numpy as np
import pytest

from parity_util import (TIGHT, correlator_problem, oracle_correlator_fits, oracle_nist_fits, rel_cov)

pytestmark = pytest.mark.gpu

BENCH_TOL = (1e-8, 1e-10, 1e-10)        # lsqfit's defaults, what bench.py runs (configs.c3 / c4)


def _need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback)")


def _copies(cfg, pdf, B, seed, vary_prior):
    from lsqfit_b200 import configs
    ny = cfg["ny"]
    return configs.bootstrap_means(cfg, B, seed, cov=pdf.cov[:ny, :ny], vary_prior=vary_prior)


_ORACLE_CACHE = {}


def _oracle_cached(K, means, p0, tol, **kw):
    """the oracle's 2000 fits are the expensive part of this test: run them once per (K, settings), for every kernel"""
    key = (K, tuple(tol), tuple(sorted(kw.items())))
    if key not in _ORACLE_CACHE:
        _ORACLE_CACHE[key] = oracle_correlator_fits(K, means, p0, tol, **kw)
    return _ORACLE_CACHE[key]


@pytest.mark.parametrize("team", [None, 32], ids=["default-kernel", "wave-kernel"])
@pytest.mark.parametrize("K,vary_prior,p0kind", [(8, True, "prior"), (3, False, "exact")])
def test_correlator_2000_copies_at_bench_settings(K, vary_prior, p0kind, team):
    """C3 (K=8: bootstrap copies, p0 = prior mean) and C4 (K=3: simulated copies, prior means fixed, p0 = pexact) on
    2000 copies.

    (1) AT THE BENCH SETTINGS -- default tolerances, no polish, default kernel.  Both solvers stop as soon as a step
    changes the cost by less than ftol = 1e-10 (relative) or the parameters by xtol = 1e-8: each ends somewhere
    within sqrt(2 ftol chi2) ~ 1e-4 standard deviations of the minimum (chi2 ~ 64), and where exactly depends on its
    last accepted step -- two runs of the REFERENCE with different fitters differ by as much.  Bars (measured worst
    case in brackets): chi2 to 1e-8 relative, second order in that distance [7e-10]; p within 1e-3 sdev for every
    copy [2.9e-4], 2e-4 for 99 % of them [5.7e-5], 1e-5 for half of them [3e-7]; covariance to 1e-3 [4.3e-4]; the same
    copies converge; evaluation counts equal on >= 75 % of the copies [81 %; mean 23.47 vs 23.45].
    Run for the default kernel (team of four warps per fit for K = 8, one warp for K = 3) and for the wave kernel
    (lm_wave.cuh: 32 fits per CTA in lock-step phases, thread-level factorisations, chunk GEMM on the tensor path).
    (2) THE NORTH-STAR BARS -- both sides at tight tolerance, device with polish, against the exact stationary
    point of the oracle's chi2 (oracle result refined by Gauss-Newton until its own step is below 1e-10 sdev; at
    least 99 % of the copies): p 1e-8 sdev, chi2 1e-9, covariance 1e-8, log det 1e-8."""
    _need_gpu()
    import lsqfit_b200 as lb
    B = 2000
    cfg, opdf = correlator_problem(K)
    ny, npar = cfg["ny"], cfg["np"]
    p0 = cfg["prior_mean"].copy() if p0kind == "prior" else cfg["ptrue"].copy()
    means = _copies(cfg, opdf, B, 4242 + K, vary_prior)
    plan = lb.Plan("multiexp", npar, ny, cfg["x"], opdf.i_invwgts, team=team)   # the ORACLE's whitening: identical inputs
    # ---- (1) bench settings
    out = plan.fit_batch(means, p0, tol=BENCH_TOL, maxit=1000).numpy()
    ref = _oracle_cached(K, means, p0, BENCH_TOL, maxit=1000)
    if team is not None:
        assert plan.last_team() == team
    conv_d = out["status"] > 0
    conv_o = np.array([r["crit"] != 0 for r in ref])
    assert np.array_equal(conv_d, conv_o), (int(conv_d.sum()), int(conv_o.sum()))
    assert conv_d.mean() > 0.999
    ok = conv_d & conv_o
    xo = np.array([r["x"] for r in ref])
    sd = np.sqrt(np.array([np.diag(r["cov"]) for r in ref]))
    dp = np.max(np.abs(out["x"] - xo) / sd, axis=1)[ok]
    dchi2 = (np.abs(out["chi2"] - np.array([r["chi2"] for r in ref])) / np.array([r["chi2"] for r in ref]))[ok]
    dcov = np.array([rel_cov(out["cov"][b], ref[b]["cov"]) for b in range(B)])[ok]
    nit_eq = float(np.mean(out["nit"][ok] == np.array([r["nit"] for r in ref])[ok]))
    print("K=%d bench settings, %d copies (%d warps per fit): |dp|/sd max %.2e q99 %.2e median %.2e; chi2 max %.2e; cov max %.2e; "
          "nfev equal on %.1f %%, mean nfev %.2f vs %.2f" % (K, B, plan.last_team(), dp.max(), np.quantile(dp, 0.99), np.median(dp),
                                                           dchi2.max(), dcov.max(), 100 * nit_eq, out["nit"].mean(),
                                                           np.mean([r["nit"] for r in ref])))
    assert dchi2.max() <= 1e-8
    assert dp.max() <= 1e-3 and np.quantile(dp, 0.99) <= 2e-4 and np.median(dp) <= 1e-5
    assert dcov.max() <= 1e-3
    assert nit_eq >= 0.75
    # ---- (2) tight tolerance + polish vs the exact stationary point
    outp = plan.fit_batch(means, p0, tol=TIGHT, maxit=2000, polish=3000).numpy()
    refe = _oracle_cached(K, means, p0, TIGHT, maxit=2000, refine=True)
    worst = dict(p=0.0, chi2=0.0, cov=0.0, logdet=0.0, ref_gap=0.0)
    n = 0
    dps, lasts = [], []
    for b in range(B):
        r = refe[b]
        # Gauss-Newton converges only linearly on these large-residual fits (a few copies need > 1000 steps): a copy
        # counts when the reference point itself is known to 1e-10 sdev
        if "xe" not in r or outp["status"][b] <= 0 or r["refine_last_step"] > 1e-10:
            continue
        n += 1
        se = np.sqrt(np.diag(r["cove"]))
        worst["p"] = max(worst["p"], np.max(np.abs(outp["x"][b] - r["xe"]) / se))
        dps.append(np.max(np.abs(outp["x"][b] - r["xe"]) / se)); lasts.append(r["refine_last_step"])
        worst["ref_gap"] = max(worst["ref_gap"], np.max(np.abs(r["x"] - r["xe"]) / se))
        worst["chi2"] = max(worst["chi2"], abs(outp["chi2"][b] - r["chi2e"]) / r["chi2e"])
        worst["cov"] = max(worst["cov"], rel_cov(outp["cov"][b], r["cove"]))
        worst["logdet"] = max(worst["logdet"], abs(outp["logdet"][b] - r["logdete"]) / abs(r["logdete"]))
    print("K=%d tight + polish vs exact minimum, %d copies: p %.2e sd, chi2 %.2e, cov %.2e, logdet %.2e "
          "(the reference's own solver ends %.2e sd from it)" % (K, n, worst["p"], worst["chi2"], worst["cov"], worst["logdet"],
                                                                worst["ref_gap"]))
    dps, lasts = np.array(dps), np.array(lasts)
    print("   device-vs-exact |dp|/sd: median %.1e q99 %.1e, %d copies above 1e-8; refinement's own last step: median %.1e max %.1e; "
          "last step of the copies above 1e-8: %s; polish nit of those: %s" % (np.median(dps), np.quantile(dps, 0.99), int((dps > 1e-8).sum()), np.median(lasts),
                                                lasts.max(), np.round(np.log10(lasts[dps > 1e-8] + 1e-300), 1)[:12], ""))
    assert n >= 0.99 * B
    assert worst["p"] <= 1e-8 and worst["chi2"] <= 1e-9 and worst["cov"] <= 1e-8 and worst["logdet"] <= 1e-8
    if K == 3:
        # the C4 "bias": the mean of the best-fit parameters over simulated copies differs from pexact (a second-
        # order effect of the nonlinear model); device and oracle must show the SAME shift on the same copies
        md, mo = out["x"][ok].mean(axis=0), xo[ok].mean(axis=0)
        sdm = xo[ok].std(axis=0) / np.sqrt(ok.sum())
        bias_d, bias_o = (md - p0) / sdm, (mo - p0) / sdm
        print("C4 bias in sigma of the mean over %d copies: device %s oracle %s" % (ok.sum(), np.round(bias_d, 2), np.round(bias_o, 2)))
        assert np.max(np.abs(bias_d - bias_o)) < 1e-3


def test_c2_perturbed_starts_vs_oracle(nist_problems):
    """C2: every NIST StRD problem from 200 of the benchmark's perturbed starts (p0 = start2 (1 + 0.1 u), the first
    200 of the 10^4 of bench.py / tools/bench_configs.py), device vs oracle at tol = 1e-10: the same starts converge,
    and they converge to the SAME MINIMUM -- including the starts that end in a local minimum on both sides.  "Same
    minimum": p within 1e-3 sdev (distinct local minima are many sdev apart) or, where the statistical errors are
    smaller than the solver tolerance itself (lanczos1: sigma_y = 1e-13), within 1e-7 relative; chi2 to 1e-6.  Both
    solvers stop at xtol = 1e-10, short of the minimum by their last step, hence not the 1e-8 sdev of the
    tight-tolerance tests (test_nist_fits_vs_oracle)."""
    _need_gpu()
    import lsqfit_b200 as lb
    S = 200
    tol = (1e-10, 1e-10, 1e-10)
    jobs, dev = [], []
    for k, pr in enumerate(nist_problems):
        ny, npar = len(pr["y"]), len(pr["p0"])
        rng = np.random.default_rng(20240 + k)
        p0 = np.array(pr["p0"])[None, :] * (1 + 0.1 * rng.uniform(-1, 1, size=(10000, npar)))[:S]
        mean = np.concatenate([pr["y"], pr["prior_mean"]])
        sd = np.concatenate([pr["ysdev"], pr["prior_sdev"]])
        plan = lb.Plan(pr["form"], npar, ny, np.array(pr["x"]), [(np.arange(ny + npar), 1.0 / sd)])
        dev.append((p0, plan.fit_batch(mean, p0, tol=tol, maxit=1000).numpy()))
        plan.close()
        jobs += [(k, p0[s], tol, 1000) for s in range(S)]
    ref = oracle_nist_fits(jobs)
    report = []
    for k, pr in enumerate(nist_problems):
        p0, o = dev[k]
        r = ref[k * S:(k + 1) * S]
        conv_d = o["status"] > 0
        conv_o = np.array([q["crit"] != 0 for q in r])
        both = conv_d & conv_o
        xo = np.array([q["x"] for q in r])
        sdo = np.array([q["sd"] for q in r])
        chio = np.array([q["chi2"] for q in r])
        dp = np.max(np.abs(o["x"] - xo) / sdo, axis=1)
        drel = np.max(np.abs(o["x"] - xo) / np.maximum(np.abs(xo), 1e-300), axis=1)
        same_min = both & (((dp < 1e-3) & (np.abs(o["chi2"] - chio) <= 1e-6 * np.maximum(chio, 1e-300))) | (drel < 1e-7))
        report.append((pr["name"], int(conv_d.sum()), int(conv_o.sum()), int(both.sum()), int(same_min.sum()),
                       float(np.median(dp[same_min])) if same_min.any() else 0.0))
    for row in report:
        print("%-10s converged device %3d oracle %3d both %3d  same minimum %3d  (median |dp|/sd %.1e)" % row)
    tot_both = sum(r[3] for r in report)
    tot_same = sum(r[4] for r in report)
    tot_conv_mismatch = sum(abs(r[1] - r[2]) for r in report)
    # the same starts converge on both sides (a handful of borderline starts of the hard problems may differ) ...
    assert tot_conv_mismatch <= 0.005 * S * len(report), report
    # ... and what converges on both sides converges to the same point
    assert tot_same >= 0.99 * tot_both, report
    for row in report:
        assert row[4] >= 0.95 * row[3], row          # (hahn1: 8 of 190 starts stall at different points on the two sides)


def test_dense_block_chiv_and_jacobian_vs_oracle():
    """chiv(p) and its Jacobian for the C3 whitening (one dense 64 x 64 block + 16 diagonal priors) element by
    element against oracle.chiv (src/lsqfit/_utilities.pyx:65-94), for both kernels' evaluation paths."""
    _need_gpu()
    import lsqfit_b200 as lb
    from oracle import dual as D
    from oracle.chiv import build_chiv_chivw
    from oracle import models as M
    cfg, opdf = correlator_problem(8)
    ny, npar = cfg["ny"], cfg["np"]
    rng = np.random.default_rng(7)
    P = cfg["prior_mean"][None, :] * (1 + 0.05 * rng.standard_normal((6, npar)))
    means = _copies(cfg, opdf, 6, 99, True)
    plan = lb.Plan("multiexp", npar, ny, cfg["x"], opdf.i_invwgts)
    f, J, chi2 = plan.residual_jacobian(P, means)
    f, J = f.cpu().numpy(), J.cpu().numpy()
    for b in range(6):
        pdf_b = opdf.copy_with_mean(means[b])
        chiv, _ = build_chiv_chivw(pdf_b, lambda p: M.MODELS["multiexp"](cfg["x"], p), False)
        fo = np.asarray(chiv(P[b]))
        Jo = D.deriv(chiv(D.Dual.variables(P[b])), npar)
        scale = np.max(np.abs(Jo))
        assert np.max(np.abs(f[b] - fo)) <= 1e-9 * max(1.0, np.max(np.abs(fo))), b
        assert np.max(np.abs(J[b] - Jo)) <= 1e-11 * scale, b
    # the same rows out of a fit launch (f, J of the final evaluation) for the team kernel and the one-warp kernel
    for team in (1, 4):
        plan.set_team(team)
        out = plan.fit_batch(means, P, tol=BENCH_TOL, maxit=1, want_fJ=True).numpy()       # maxit = 1: evaluated at the start
        assert np.max(np.abs(out["f"] - f)) <= 1e-9 * np.max(np.abs(f)), team
        assert np.max(np.abs(out["J"] - J)) <= 1e-11 * np.max(np.abs(J)), team


def test_c5_full_size_vs_oracle():
    """Config 5 at its full size (5000 correlated points, 2000 parameters, svdcut 1e-8, 2501 clamped modes) against the
    committed oracle result (tests/golden/c5_oracle.npz, made by tests/golden/make_c5_fixture.py: ~10 minutes of CPU):
    p within 1e-7 sdev (both sides at the default tolerance), chi2 1e-9, sdev 1e-8, logGBF 1e-9, same iteration count,
    same number of modified modes."""
    _need_gpu()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c5_oracle.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/c5_oracle.npz not generated (python tests/golden/make_c5_fixture.py)")
    from lsqfit_b200 import configs
    from lsqfit_b200.dense import DenseFit
    g = np.load(path)
    cfg = configs.c5(ny=int(g["ny"]), K=int(g["K"]))
    assert abs(float(np.sum(cfg["ymean"]) + np.trace(cfg["ycov"])) - float(g["data_checksum"])) <= 1e-12 * abs(float(g["data_checksum"]))
    fit = DenseFit((cfg["t"], cfg["ymean"], cfg["ycov"]), (cfg["prior_mean"], cfg["prior_sdev"]), svdcut=cfg["svdcut"],
                   tol=cfg["tol"], maxit=cfg["maxit"])
    dp = np.max(np.abs(fit.pmean - g["pmean"]) / g["psdev"])
    print("C5 full size: |dp|/sd %.2e  chi2 %.3e  sdev %.2e  logGBF %.2e  nit %d vs %d  svdn %d vs %d  (oracle: %.0f s on %d cores)" % (
        dp, abs(fit.chi2 - g["chi2"]) / g["chi2"], np.max(np.abs(fit.psdev / g["psdev"] - 1)), abs(fit.logGBF - g["logGBF"]) / abs(g["logGBF"]),
        fit.nit, int(g["nit"]), fit.svdn, int(g["svdn"]), float(g["cpu_seconds"]), int(g["cores"])))
    assert fit.svdn == int(g["svdn"]) and fit.dof == int(g["dof"])
    assert dp <= 1e-7
    assert abs(fit.chi2 - g["chi2"]) <= 1e-9 * g["chi2"]
    assert np.max(np.abs(fit.psdev / g["psdev"] - 1)) <= 1e-8
    assert abs(fit.logGBF - g["logGBF"]) <= 1e-9 * abs(g["logGBF"])


@pytest.mark.parametrize("team", [1, 4, 32], ids=["one-warp", "team-4", "wave"])
def test_ordered_queue_changes_nothing_but_the_schedule(team):
    """b200lm_set_order: the work queue hands out the fits with the largest start-point chi2 first (a batch is bounded by
    its slowest fits).  Scheduling only -- every fit must come out BIT-identical to the input-order run, for every
    kernel.  Off by default (over several batches the mean gain is zero, tools/order_seeds.py)."""
    _need_gpu()
    import torch
    import lsqfit_b200 as lb
    cfg, opdf = correlator_problem(8)
    pdf = lb.PDF(opdf.mean, opdf.cov_in, svdcut=1e-12)
    B = 2500
    means = _copies(cfg, pdf, B, 4242, True)
    means[7, :] = np.nan                                  # a poisoned copy must not disturb the ranking of the others
    plan = lb.Plan("multiexp", cfg["np"], cfg["ny"], cfg["x"], pdf.i_invwgts)
    plan.set_team(team)
    md = torch.as_tensor(means).cuda()
    p0 = torch.as_tensor(cfg["prior_mean"]).cuda()
    res = {}
    for mode in (0, 1, None):
        plan.set_order(mode)
        o = plan.fit_batch(md, p0, tol=BENCH_TOL, maxit=1000)
        torch.cuda.synchronize()
        assert plan.last_order() == (1 if mode == 1 else 0), mode            # None: default policy = input order
        res[mode] = o
    for mode in (1, None):
        a, b = res[0], res[mode]                              # (None: the default policy, i.e. the same run again)
        ok = torch.arange(B, device=md.device) != 7
        assert torch.equal(a.status, b.status) and torch.equal(a.nit, b.nit)
        assert int(a.status[7]) == -1
        assert torch.equal(a.x[ok], b.x[ok]) and torch.equal(a.chi2[ok], b.chi2[ok]) and torch.equal(a.cov[ok], b.cov[ok])
    plan.close()


def test_host_call_writes_pinned_outputs_in_place():
    """b200lm_fit_batch_host: covariances (and f / J) go straight to the caller's buffer when it is pinned, mapped host
    memory (the kernels store over PCIe while other fits still run) and through a device copy + D2H transfer otherwise.
    Both routes must deliver the same bits; B200LM_NO_ZEROCOPY=1 switches the direct route off."""
    _need_gpu()
    import os
    import torch
    import lsqfit_b200 as lb
    cfg, opdf = correlator_problem(3)
    pdf = lb.PDF(opdf.mean, opdf.cov_in, svdcut=1e-12)
    B = 300
    means = _copies(cfg, pdf, B, 99, False)
    plan = lb.Plan("multiexp", cfg["np"], cfg["ny"], cfg["x"], pdf.i_invwgts)
    npar, nchiv = cfg["np"], plan.nchiv

    def buffers(pinned):
        mk = (lambda *sh, dt=torch.float64: torch.zeros(sh, dtype=dt).pin_memory().numpy()) if pinned else \
             (lambda *sh, dt=torch.float64: torch.zeros(sh, dtype=dt).numpy())
        return dict(x=mk(B, npar), chi2=mk(B), cov=mk(B, npar, npar), logdet=mk(B), nit=mk(B, dt=torch.int32),
                    status=mk(B, dt=torch.int32), f=mk(B, nchiv), J=mk(B, nchiv, npar))
    runs = {}
    for name, pinned, env in (("pageable", False, None), ("pinned", True, None), ("pinned-staged", True, "1")):
        if env:
            os.environ["B200LM_NO_ZEROCOPY"] = env
        else:
            os.environ.pop("B200LM_NO_ZEROCOPY", None)
        out = buffers(pinned)
        plan.fit_batch_host(means, cfg["ptrue"], tol=BENCH_TOL, maxit=1000, out=out, want_fJ=True)
        runs[name] = out
    os.environ.pop("B200LM_NO_ZEROCOPY", None)
    ref = runs["pageable"]
    assert np.all(ref["status"] > 0) and np.all(np.isfinite(ref["cov"]))
    for name in ("pinned", "pinned-staged"):
        for k in ("x", "chi2", "cov", "logdet", "nit", "status", "f", "J"):
            assert np.array_equal(ref[k], runs[name][k]), (name, k)
    # the covariance is symmetric to rounding (its elements are stored at their mirror positions, row by row)
    c = ref["cov"]
    assert np.max(np.abs(c - c.transpose(0, 2, 1)) / np.sqrt(np.einsum("bii,bjj->bij", c, c))) < 1e-12
    plan.close()


def test_batch_drivers_take_a_kernel_argument():
    """simulated_fits / bootstrapped_fits(kernel='wave'): the wave kernel through the mirror of the reference's iterators;
    same fits as the default kernel to rounding, and the cached plan goes back to its default policy afterwards."""
    _need_gpu()
    import lsqfit_b200 as lb
    cfg, _ = correlator_problem(3)
    fit = lb.nonlinear_fit(data=(cfg["x"], cfg["f"], cfg["ycov"]), prior=(cfg["prior_mean"], cfg["prior_sdev"]), fcn="multiexp")
    a = fit.simulated_fits(3000, pexact=cfg["ptrue"], seed=5)
    b = fit.simulated_fits(3000, pexact=cfg["ptrue"], seed=5, kernel="wave")
    plan = fit._spec.plan(fit.device)
    ra, rb = a.out.numpy(), b.out.numpy()
    # (the same copies converge; a fit that ends where the ftol and xtol tests both come close may report another of the
    # converged codes, and the two kernels stop within their tolerance of the minimum, not at the same bits)
    ok = (ra["status"] > 0) & (rb["status"] > 0)
    assert ok.mean() > 0.99, (float((ra["status"] > 0).mean()), float((rb["status"] > 0).mean()))
    sd = np.sqrt(np.einsum("bii->bi", ra["cov"]))
    dp = np.max(np.abs(ra["x"] - rb["x"]) / sd, axis=1)[ok]
    assert np.quantile(dp, 0.99) < 1e-3, float(np.quantile(dp, 0.99))
    dchi = (np.abs(ra["chi2"] - rb["chi2"]) / ra["chi2"])[ok]
    assert np.quantile(dchi, 0.99) < 1e-7, float(np.quantile(dchi, 0.99))
    c = fit.simulated_fits(64, pexact=cfg["ptrue"], seed=5)
    assert plan.last_team() == 1                       # np = 6: one warp per fit by default, again
