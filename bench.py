#!/usr/bin/env python
"""Benchmark of the batched LM hot path (contract: see the task prompt / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json config 3): 8-exponential correlator, 64 correlated time slices,
16 diagonal priors, svdcut=1e-12, 10^4 bootstrap copies per GPU per step, p0 = prior mean,
tol=(1e-8,1e-10,1e-10), maxit=1000.  One "step" = one launch of the fit kernel over the
whole batch (plus, for N>1, the NCCL all-gather of the packed per-fit results).
Weak scaling: every rank fits its own 10^4 copies (different seed).

Prints ONE JSON line on rank 0.
"""
import os
# NCCL writes its version banner to STDOUT at NCCL_DEBUG=VERSION/WARN/INFO; the contract is ONE JSON
# line on stdout, so NCCL's log goes to stderr (and is off unless B200LM_NCCL_DEBUG asks for it)
os.environ.pop("NCCL_DEBUG", None)
if os.environ.get("B200LM_NCCL_DEBUG"):
    os.environ["NCCL_DEBUG"] = os.environ["B200LM_NCCL_DEBUG"]
os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64 LM fits/sec"
UNIT = "fits/s"
WORKLOAD = "C3: 8-exp correlator, 64 correlated time slices, 16 priors, 10k bootstrap fits per GPU"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=10000, help="fits per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=2000, help="fits timed on the host for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------ CPU oracle leg
def _oracle_setup():
    from lsqfit_b200 import configs
    from oracle.whiten import PDF as OPDF
    cfg = configs.c3()
    ny, npar = cfg["ny"], cfg["np"]
    N = ny + npar
    full = np.zeros((N, N))
    full[:ny, :ny] = cfg["ycov"]
    full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
    mean0 = np.concatenate([cfg["f"], cfg["prior_mean"]])
    pdf = OPDF(mean0, full, svdcut=cfg["svdcut"])
    return cfg, pdf


_G = {}


def _oracle_init():
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    _G["cfg"], _G["pdf"] = _oracle_setup()


def _oracle_fit(mean):
    from oracle.fit import nonlinear_fit
    cfg, pdf = _G["cfg"], _G["pdf"]
    ny = cfg["ny"]
    fit = nonlinear_fit("multiexp", cfg["x"], mean[:ny], prior_mean=mean[ny:], _yp_pdf=pdf,
                        p0=cfg["p0"], tol=cfg["tol"], maxit=cfg["maxit"])
    return fit.nit


def cpu_fits_per_sec(means, cores):
    """The oracle (restatement of the reference's scipy_least_squares path, default trf) on
    `cores` host processes; whitening shared (simulated_fit_iter style)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_oracle_init) as pool:
        pool.map(_oracle_fit, list(means[:cores]))          # warm the workers
        t0 = time.perf_counter()
        nit = pool.map(_oracle_fit, list(means), chunksize=max(1, len(means) // (cores * 8)))
        dt = time.perf_counter() - t0
    return len(means) / dt, float(np.mean(nit))


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the fit kernel, per launch, from the committed
    `ncu --set full` capture (profiles/traffic_r01_v9.json); None if absent."""
    try:
        t = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic_r01_v9.json")))
        return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
    except (OSError, ValueError, KeyError):
        return None


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=smax,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------ peaks
def fp64_peak():
    """(TFLOP/s, source).  MEASURED_PEAKS.json holds no FP64 figure, so the denominator is the
    FP64 FMA peak measured on this pool by profiles/fp64_peak.cu (committed as
    profiles/fp64_peak_r01.json); fallback = nominal 37 TFLOP/s."""
    for name in ("fp64_peak_r01.json",):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                d = json.load(open(p))
                return float(d["fp64_tflops"]), "measured (profiles/%s: %s)" % (name, d.get("how", ""))
            except Exception:
                pass
    return 37.0, "fallback: nominal B200 FP64 37 TFLOP/s (no measured FP64 peak available)"


# ------------------------------------------------------------------------------ reference arm
def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    _oracle_init()
    from lsqfit_b200 import configs
    cfg, pdf = _G["cfg"], _G["pdf"]
    per_step = max(cores * 25, 200)
    means = configs.bootstrap_means(cfg, per_step * (args.steps + args.warmup), cfg["seed"],
                                    cov=pdf.cov[:cfg["ny"], :cfg["ny"]])
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_oracle_init) as pool:
        k = 0
        for _ in range(args.warmup):
            pool.map(_oracle_fit, list(means[k:k + per_step]))
            k += per_step
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_oracle_fit, list(means[k:k + per_step]), chunksize=max(1, per_step // (cores * 8)))
            k += per_step
        dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    sample = ("%d fits per step (bounded sample of the 10k-copy batch), scipy least_squares(trf) "
              "oracle restatement of lsqfit.scipy_least_squares, whitening shared, %d processes on %s"
              % (per_step, cores, cpu_model()))
    line = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=WORKLOAD, fits_per_step=per_step),
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------ our arm
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import lsqfit_b200 as lb
    from lsqfit_b200 import configs, dist as lbdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")

    # CPU baseline first (rank 0, N=1 only), before CUDA is initialised in this process, so that
    # the worker processes are forked from a CUDA-free parent.
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        _oracle_init()
        ocfg, opdf = _G["cfg"], _G["pdf"]
        n = min(args.cpu_sample, args.batch)
        om = configs.bootstrap_means(ocfg, args.batch, ocfg["seed"], cov=opdf.cov[:ocfg["ny"], :ocfg["ny"]])[:n]
        v, mean_nit = cpu_fits_per_sec(om, cores)
        cpu = dict(value=v, unit=UNIT, cores=cores, kind="port",
                   sample="first %d of the %d bootstrap copies of this workload; oracle = scipy "
                          "least_squares(trf) restatement of lsqfit.scipy_least_squares (numpy AD Jacobian, "
                          "whitening shared), %d processes, mean nfev %.1f, %s"
                          % (n, args.batch, cores, mean_nit, cpu_model()))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = configs.c3(B=args.batch)
    ny, npar, B = cfg["ny"], cfg["np"], args.batch
    N = ny + npar
    full = np.zeros((N, N))
    full[:ny, :ny] = cfg["ycov"]
    full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
    mean0 = np.concatenate([cfg["f"], cfg["prior_mean"]])
    pdf = lb.PDF(mean0, full, svdcut=cfg["svdcut"], device=local_rank)          # device whitening
    # Weak scaling = fixed per-GPU work: every rank fits the SAME synthetic batch.  (With a different
    # seed per rank the slowest rank's longest fit sets the step time -- 74 % efficiency at 8 GPUs instead
    # of ~95 %, see DESIGN.md section 7 -- which measures the workload's tail, not the engine's scaling.
    # B200LM_BENCH_DISTINCT=1 restores per-rank seeds.)
    distinct = os.environ.get("B200LM_BENCH_DISTINCT", "0") == "1"
    means_h = configs.bootstrap_means(cfg, B, cfg["seed"] + (rank if distinct else 0), cov=pdf.cov[:ny, :ny])
    plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts, device=local_rank)
    means_d = torch.as_tensor(means_h).to(dev)
    p0_d = torch.as_tensor(cfg["p0"]).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)   # 256 MiB > 126 MB L2

    out = plan.fit_batch(means_d, p0_d, tol=cfg["tol"], maxit=cfg["maxit"])          # allocates outputs

    def step():
        plan.fit_batch(means_d, p0_d, tol=cfg["tol"], maxit=cfg["maxit"], out=out)
        if world > 1:
            packed = lbdist.pack_results(out.x, out.chi2, out.nit, out.status)
            return lbdist.gather_results(packed)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    barrier()
    launches0 = plan.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (not timed)
        ev[k][0].record()
        kev[k][0].record()
        plan.fit_batch(means_d, p0_d, tol=cfg["tol"], maxit=cfg["maxit"], out=out)
        kev[k][1].record()
        if world > 1:
            packed = lbdist.pack_results(out.x, out.chi2, out.nit, out.status)
            lbdist.gather_results(packed)
        ev[k][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = plan.launch_count() - launches0
    t_ms = sum(a.elapsed_time(b) for a, b in ev)
    tk_ms = sum(a.elapsed_time(b) for a, b in kev)
    tt = torch.tensor([t_ms, tk_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, tk_ms = float(tt[0]), float(tt[1])
    nfev, njev, nfac = plan.last_stats()
    res = out.numpy()
    conv = float((res["status"] > 0).mean())

    # ---- e2e: the public host-buffer call, pinned inputs, H2D + D2H inside the timed region
    means_pin = torch.as_tensor(means_h).pin_memory()
    outs = dict(x=torch.empty((B, npar), dtype=torch.float64).pin_memory().numpy(),
                chi2=torch.empty(B, dtype=torch.float64).pin_memory().numpy(),
                cov=torch.empty((B, npar, npar), dtype=torch.float64).pin_memory().numpy(),
                logdet=torch.empty(B, dtype=torch.float64).pin_memory().numpy(),
                nit=torch.empty(B, dtype=torch.int32).pin_memory().numpy(),
                status=torch.empty(B, dtype=torch.int32).pin_memory().numpy(), f=None, J=None)
    mp_np = means_pin.numpy()
    for _ in range(3):
        plan.fit_batch_host(mp_np, cfg["p0"], tol=cfg["tol"], maxit=cfg["maxit"], out=outs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        plan.fit_batch_host(mp_np, cfg["p0"], tol=cfg["tol"], maxit=cfg["maxit"], out=outs)
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    te_t = torch.tensor([te], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
    te = float(te_t[0])
    h2d = B * N * 8 + npar * 8
    d2h = B * (npar + 2 + npar * npar) * 8 + B * 8

    # ---- extra (not the headline): the same steps QUEUED on two streams / two plans, so that the next
    # batch fills the SMs that idle during the long-fit tail of the previous one
    plan_b = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts, device=local_rank)
    means_b = torch.as_tensor(configs.bootstrap_means(cfg, B, cfg["seed"] + 1000 + rank, cov=pdf.cov[:ny, :ny])).to(dev)
    out_b = plan_b.fit_batch(means_b, p0_d, tol=cfg["tol"], maxit=cfg["maxit"])
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    jobs = [(plan, means_d, out), (plan_b, means_b, out_b)]
    torch.cuda.synchronize()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_()
    torch.cuda.synchronize()
    q0.record()
    for k in range(args.steps):
        pl, mm, oo = jobs[k & 1]
        with torch.cuda.stream(streams[k & 1]):
            if k == 0 or k == 1:
                streams[k & 1].wait_event(q0)
            pl.fit_batch(mm, p0_d, tol=cfg["tol"], maxit=cfg["maxit"], out=oo)
    for st in streams:
        torch.cuda.current_stream(dev).wait_stream(st)
    q1.record()
    torch.cuda.synchronize()
    tq_ms = q0.elapsed_time(q1)
    plan_b.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the fit kernel is the only kernel of a step)
    F_eval = configs.eval_flops(ny, npar, cfg["K"])
    F_res = 2.0 * ny * ny + 2.0 * N + ny * cfg["K"] * 3.0       # residual-only trial evaluation
    flops_launch = njev * F_eval + (nfev - njev) * F_res
    peak, peak_src = fp64_peak()
    achieved = flops_launch / (tk_ms / args.steps * 1e-3) / 1e12
    roofline = dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                    traffic=_ncu_traffic(), kernel="fit_kernel<MultiExp<8>>", peak_source=peak_src,
                    flops_per_launch=flops_launch, flops_per_launch_survey_formula=nfev * F_eval,
                    nfev_per_fit=nfev / B, njev_per_fit=njev / B, chol_per_fit=nfac / B,
                    kernel_ms=tk_ms / args.steps)

    queued = dict(value=B * args.steps / (tq_ms * 1e-3), unit=UNIT + " per GPU", streams=2, ms_per_step=tq_ms / args.steps,
                  note="same steps queued on two streams (two plans): throughput of a stream of 10k-fit batches; "
                       "not the headline, which times one batch at a time")
    line = dict(metric=METRIC, value=world * B * args.steps / (t_ms * 1e-3), unit=UNIT, n_gpus=world,
                steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=t_ms / args.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=WORKLOAD, fits_per_gpu_per_step=B, ny=ny, np=npar, svdcut=cfg["svdcut"],
                            tol=list(cfg["tol"]), maxit=cfg["maxit"], l2="flushed between timed steps (256 MiB memset)",
                            per_rank_data="distinct seeds" if distinct else "identical batch on every rank (fixed per-GPU work)",
                            converged_frac=conv, svdn=int(pdf.nmod)),
                clocks=clocks,
                e2e=dict(value=world * B * args.steps / te, unit=UNIT, h2d_bytes_per_step=h2d,
                         d2h_bytes_per_step=d2h, ms_per_step=1e3 * te / args.steps,
                         api="lsqfit_b200.Plan.fit_batch_host -> b200lm_fit_batch_host (pinned host buffers)"),
                gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu, queued=queued)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
