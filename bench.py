#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 ensemble ODE engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c3]

A "step" is one pass of the hot path over one batch of synthetic draws: the whole ensemble solved
from t=0 to 365 d with Tsit5 + PID controller and written out daily (diffeqsolve + SaveAt).
Workload at N=1 = BASELINE.json configs[3]: multi-strain age-stratified SEIRS, 100k draws per GPU
(weak scaling: every rank solves its own 100k draws, no data-path collective).

  value      trajectories/s, inputs resident in HBM, device-timed with CUDA events (max over ranks)
  e2e        same metric through the public API (simulate_ensemble) with HOST buffers:
             H2D of the draws and D2H of every saved trajectory inside the timed region
  roofline   dominant kernel (lane_solver_kernel): algorithmic bytes / measured duration vs the
             measured HBM peak; the FP64-FMA view of the same kernel is in roofline_fp64
  cpu_baseline   the CPU oracle (restatement of the reference's diffrax path, OpenMP) on a bounded
             sample of the same workload, timed on this box's host cores

--impl reference times that CPU oracle as the reference arm (the reference's own JAX/diffrax stack
cannot be installed in this image: no jax/diffrax wheels, no network -- see DESIGN.md).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (case name in dynode_b200/synthetic.py, draws per GPU, description)
    "c4": ("seirs_multi_a2s3", 100_000,
           "C4 multi-strain age-stratified SEIRS+C (A=2,S=3,n=26), 365 d, daily SaveAt T=366"),
    "c3": ("seirs_seasonal", 1_000_000, "C3 seasonally forced SEIRS (n=4), 365 d, daily SaveAt T=366"),
    # BASELINE.json configs[4]: a NUTS run, not an ensemble solve -- see run_c5
    "c5": ("seirs_multi_g6s3", 1024,
           "C5 age(3) x risk(2) x strain(3) SEIRS+C (n=78) NUTS inference of 3 r0 + 3 infectious periods, Poisson "
           "likelihood on 120 d of daily incidence, 1024 chains sharded over the GPUs, NCCL gather of "
           "posterior-predictive trajectories"),
}
F_RHS = {"seirs_multi_a2s3": 130, "seirs_seasonal": 18}  # flops per RHS evaluation


def gradient_work(n, m, T, f_rhs, n_att_total, n_acc_total, B, P=0, adjoint=False):
    """Algorithmic flops of one fused log-likelihood + gradient launch (SURVEY.md 8d "Gradient work"; DESIGN.md
    section 4).  Per attempted step the primal costs 6 F_rhs + 70 n + 50 (stage sums 49 n, error estimate and norm
    21 n, controller 50); nothing is saved, but the observed compartment (m elements) is evaluated by dense output at
    the T save times (14 m + 45) and enters the Poisson term (25 m: difference, clamp, log, multiply-add).
      forward mode, P directions: every RHS also pushes P tangents (a JVP is 2 F_rhs), the stage sums and the dense
        output run for the tangents too; the error estimate does not:  (6 N_att + 3) F_rhs (1 + 2 P)
        + N_att (49 n (1 + P) + 21 n + 50) + T (14 m + 45) (1 + P) + T m (25 + 2 P)
      discrete adjoint: the primal forward sweep, then per ACCEPTED step the 6 stages are recomputed (6 F_rhs + 49 n),
        7 vector-Jacobian products are pulled back (2 F_rhs each, + their parameter contractions ~ F_rhs each) and the
        stage cotangents summed (49 n):  primal + N_acc (27 F_rhs + 98 n) + T m 25"""
    primal = (6 * n_att_total + 3 * B) * f_rhs + n_att_total * (70 * n + 50) + B * T * (14 * m + 45) + B * T * m * 25
    if adjoint:
        return primal + n_acc_total * (27 * f_rhs + 98 * n)
    return ((6 * n_att_total + 3 * B) * f_rhs * (1 + 2 * P) + n_att_total * (49 * n * (1 + P) + 21 * n + 50)
            + B * T * (14 * m + 45) * (1 + P) + B * T * m * (25 + 2 * P))


def algorithmic_work(n, p_in, T, n_saved, f_rhs, n_att_total, B):
    """flops = (6*N_att+3)*F_rhs + N_att*(70n+50) + T*(14*n_s+45);
    bytes = 8*(n+P_in) + 8*T*n_s + 16 per trajectory."""
    flops = (6 * n_att_total + 3 * B) * f_rhs + n_att_total * (70 * n + 50) + B * T * (14 * n_saved + 45)
    byts = B * (8 * (n + p_in) + 8 * T * n_saved + 16)
    return flops, byts


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_ms: int = 50):
        self.gpu, self.rows, self.proc, self.thr = gpu_index, [], None, None
        self.period_ms = int(os.environ.get("DYNODE_BENCH_CLOCK_MS", period_ms))

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms),
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        self.thr = threading.Thread(target=lambda: [self.rows.append(l) for l in self.proc.stdout], daemon=True)
        self.thr.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs(workload: str, B: int, seed: int):
    from dynode_b200.synthetic import make_case
    return make_case(WORKLOADS[workload][0], B, seed=seed)


def host_threads() -> int:
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the
    CPU arm: the thread count is passed to the oracle explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_leg(workload: str, min_seconds: float, chunk: int, seed: int):
    """Time the CPU oracle (all host threads) on chunk-sized pieces of the workload until
    `min_seconds` of CPU wall time have been spent (the bounded sample of the cpu_baseline leg)."""
    from oracle import oracle as orc
    case = make_inputs(workload, chunk, seed)
    fam, dims, theta, shared = case["oracle"]
    nthr = host_threads()
    orc.solve(fam, dims, case["y0"][:256] if np.ndim(case["y0"]) == 2 else case["y0"], theta[:256], shared,
              t1=case["t1"], nthreads=nthr)  # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    sample_chunks = 0
    while True:
        orc.solve(fam, dims, case["y0"], theta, shared, t1=case["t1"], nthreads=nthr)
        sample_chunks += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return sample_chunks * chunk / dt, nthr, dt, sample_chunks


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path = the oracle port (kind 'port')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, B, desc = WORKLOADS[args.workload]
    if args.workload == "c5":
        return run_reference_c5(args)
    chunk = 8192
    from oracle import oracle as orc
    case = make_inputs(args.workload, chunk, 20260101)
    fam, dims, theta, shared = case["oracle"]
    cores = host_threads()
    for _ in range(max(1, min(args.warmup, 2))):
        orc.solve(fam, dims, case["y0"], theta, shared, t1=case["t1"], nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.solve(fam, dims, case["y0"], theta, shared, t1=case["t1"], nthreads=cores)
    dt = time.perf_counter() - t0
    v = args.steps * chunk / dt
    sample = f"{chunk} draws of the same workload per step, {args.steps} steps, OpenMP over {cores} threads"
    line = {
        "impl": "reference", "metric": "solved trajectories/s", "value": v, "unit": "trajectories/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "draws_per_step": chunk, "rtol": 1e-5, "atol": 1e-6, "solver": "Tsit5"},
        "cpu_baseline": {"value": v, "unit": "trajectories/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference JAX/diffrax stack not installable here; CPU oracle (C++ restatement, OpenMP) timed instead",
    }
    print(json.dumps(line))


def run_reference_c5(args):
    """Reference arm of --workload c5: what one NUTS leapfrog costs on the host -- the CPU oracle's solve with the six
    forward tangents (r0 and infectious period of three strains -> beta, gamma) plus the Poisson log-likelihood and
    its gradient, on all host threads, for `chunk` chains per step.  (numpyro itself is not installable here.)"""
    from oracle import oracle as orc
    from dynode_b200.synthetic import make_case
    chunk = 1024
    case = make_case("seirs_multi_g6s3", chunk, seed=20260105)
    fam, dims, theta, shared = case["oracle"]
    cores = host_threads()
    wrt = [0, 1, 2, 3, 4, 5]  # beta_s, gamma_s
    sizes = case["model"].compartment_sizes()
    idx = list(range(sum(sizes[:4]), sum(sizes)))
    truth, _, _ = orc.solve(fam, dims, case["y0"][:1], theta[:1], shared, t1=120, save_idx=idx)
    obs = np.abs(np.diff(truth[0], axis=0)) + 0.05

    def step():
        ys, dys, st = orc.solve(fam, dims, case["y0"], theta, shared, t1=120, save_idx=idx, wrt=wrt, nthreads=cores)
        return orc.poisson_incidence(ys, dys, obs)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = args.steps * chunk / dt
    sample = f"{chunk} chains' (log-density, gradient) per step, {args.steps} steps, OpenMP over {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "NUTS grad-evals/s", "value": v, "unit": "grad-evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS["c5"][2], "chains_per_step": chunk},
        "cpu_baseline": {"value": v, "unit": "grad-evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "grad-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "numpyro / JAX not installable here; CPU oracle (forward tangents + Poisson gradient) timed instead"}))


def nuts_leg(args, dev, world, barrier, fp64_peak_tf=None):
    """NUTS grad-evals/s (BASELINE metric 2) on config 2, per GPU and whole job.

    kernel: dynode_poisson_loglik_grad_f64 alone on B random unconstrained draws z ~ N(0,1)^2 mapped
            through the priors' bijectors (device-resident, CUDA events);
    sampler: the many-chain NUTS of dynode_b200.infer on the fused model, counting the leapfrogs that
            belong to a tree (wall clock, warm-up + sampling)."""
    import torch
    import torch.distributed as dist

    from dynode_b200.examples import sir_infer_parameters as m
    from dynode_b200.infer import MCMC, NUTS, ModelDensity, PRNGKey

    rank = int(os.environ.get("RANK", "0"))
    cfg = m.get_config()
    obs = m.synthetic_incidence(100).to(dev)
    md = ModelDensity(m.model_fused, (), dict(config=cfg, tf=100, obs_data=obs))
    out = {"config": "C2 age-stratified SIR (n=6), 100 d, Poisson likelihood on diff(R), P=2 (beta, gamma)",
           "unit": "grad-evals/s"}
    # kernel-level: one launch = B (log-density, gradient) evaluations
    Bk = 1 << 20
    g = torch.Generator(device=dev).manual_seed(20260102 + rank)
    Z = torch.randn(Bk, 2, dtype=torch.float64, device=dev, generator=g)
    cons = md.constrain(Z)
    r0, inf = cons["strains_0_r0"], cons["strains_0_infectious_period"]
    from dynode_b200 import _lib, engine
    from dynode_b200.synthetic import make_case
    case = make_case("sir_age2", 1)
    prm = {"beta": (r0 / inf).reshape(-1, 1).contiguous(), "gamma": (1.0 / inf).reshape(-1, 1).contiguous()}
    y0 = torch.as_tensor(case["y0"], dtype=torch.float64, device=dev)
    contact = torch.as_tensor(case["contact"], dtype=torch.float64, device=dev)
    ts = np.linspace(0.0, 100.0, 101)
    opts = engine.SolverOptions(t1=100.0)
    wrt = [_lib.wrt_id(_lib.P_BETA, 0), _lib.wrt_id(_lib.P_GAMMA, 0)]

    def kstep():
        return engine.poisson_loglik_grad(case["model"], y0, prm, contact, opts, ts, 2, obs, 0.0, wrt=wrt, B=Bk)

    for _ in range(3):
        lp, grad, st = kstep()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lp, grad, st = kstep()
    e1.record()
    barrier()
    kms = e0.elapsed_time(e1) / 5
    if world > 1:
        t = torch.tensor([kms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kms = float(t.item())
    n_att = float(st[:, 3].double().mean())
    def fp64_roofline(flops, ms):
        if not fp64_peak_tf:
            return None
        ach = flops / (ms * 1e-3) / 1e12
        return {"bound": "fp64_fma", "achieved": ach, "peak": fp64_peak_tf, "unit": "TFLOP/s",
                "frac": ach / fp64_peak_tf, "algorithmic_flops_per_launch": flops,
                "peak_source": "dynode_probe_dfma measured in this run"}

    fl2 = gradient_work(6, 2, 101, 26, int(st[:, 3].sum()), int(st[:, 1].sum()), Bk, P=2)
    out["kernel"] = {"value": world * Bk / (kms * 1e-3), "draws_per_gpu": Bk, "ms_per_launch": kms,
                     "mean_attempted_steps": n_att,
                     "finite_fraction": float(torch.isfinite(lp).double().mean()),
                     "roofline_fp64": fp64_roofline(fl2, kms)}
    # config 5 (age x risk x strain SEIRS + C, n = 78, six differentiated rates): kernel-level only
    try:
        from dynode_b200.examples import seirs_age_risk_strain as m5
        B5 = 32768
        case5 = make_case("seirs_multi_g6s3", B5, seed=20260105 + rank)
        obs5 = m5.synthetic_incidence(120).to(dev).reshape(120, -1).contiguous()
        prm5 = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in case5["params"].items()}
        y05 = torch.as_tensor(case5["y0"], dtype=torch.float64, device=dev)
        c5 = torch.as_tensor(case5["contact"], dtype=torch.float64, device=dev)
        wrt5 = [_lib.wrt_id(_lib.P_BETA, s_) for s_ in range(3)] + [_lib.wrt_id(_lib.P_GAMMA, s_) for s_ in range(3)]
        ts5 = np.linspace(0.0, 120.0, 121)
        o5 = engine.SolverOptions(t1=120.0)

        def k5():
            return engine.poisson_loglik_grad(case5["model"], y05, prm5, c5, o5, ts5, 4, obs5, 0.0, wrt=wrt5, B=B5)

        for _ in range(2):
            k5()
        barrier()
        e0.record()
        for _ in range(3):
            lp5, g5, st5 = k5()
        e1.record()
        barrier()
        ms5 = e0.elapsed_time(e1) / 3
        def k5a():
            return engine.poisson_loglik_adjoint(case5["model"], y05, prm5, c5, o5, ts5, 4, obs5, 0.0, B=B5, cap=256)

        for _ in range(2):
            k5a()
        barrier()
        e0.record()
        for _ in range(3):
            lpa, ga, _, sta = k5a()
        e1.record()
        barrier()
        ms5a = e0.elapsed_time(e1) / 3
        att5, acc5 = int(st5[:, 3].sum()), int(st5[:, 1].sum())
        out["kernel_config5_adjoint"] = {
            "value": world * B5 / (ms5a * 1e-3), "draws_per_gpu": B5, "ms_per_launch": ms5a,
            "roofline_fp64": fp64_roofline(gradient_work(78, 18, 121, 560, att5, acc5, B5, adjoint=True), ms5a),
            "directions": "all 12 rates (+ y0 on request) from one reverse sweep",
            "max_abs_diff_lp_vs_forward": float((lpa - lp5).abs().max()),
            "max_rel_diff_grad_vs_forward": float(((ga[:, :6] - g5).abs() / (g5.abs() + 1e-300)).max())}
        out["kernel_config5"] = {"value": world * B5 / (ms5 * 1e-3), "draws_per_gpu": B5, "ms_per_launch": ms5,
                                 "directions": 6, "tangent_groups_in_one_launch": 6,
                                 # six single-direction work items per draw, each repeating the primal
                                 "roofline_fp64": fp64_roofline(6 * gradient_work(78, 18, 121, 560, att5, acc5, B5, P=1), ms5),
                                 "config": "C5 age(3) x risk(2) x strain(3) SEIRS + C (n=78), 120 d, Poisson on diff(C)"}
    except Exception as exc:  # reported, not hidden: the headline NUTS numbers above do not depend on it
        out["kernel_config5"] = {"error": f"{type(exc).__name__}: {exc}"}
    # sampler-level: the whole NUTS run (warm-up + sampling) at `--nuts-chains` chains per GPU, and at 4096
    def sampler(C):
        mc = MCMC(NUTS(m.model_fused, max_tree_depth=10), num_warmup=100, num_samples=50, num_chains=C,
                  progress_bar=False)
        barrier()
        t0 = time.perf_counter()
        mc.run(PRNGKey(8675314 + rank), config=cfg, tf=100, obs_data=obs)
        barrier()
        dt = time.perf_counter() - t0
        evals = torch.tensor([float(mc.engine.grad_evals), dt], dtype=torch.float64, device=dev)
        if world > 1:
            tot = evals.clone()
            dist.all_reduce(tot[:1], op=dist.ReduceOp.SUM)
            dist.all_reduce(tot[1:], op=dist.ReduceOp.MAX)
            evals = tot
        s = mc.get_samples()
        return {"value": float(evals[0]) / float(evals[1]), "chains_per_gpu": C, "num_warmup": 100,
                "num_samples": 50, "wall_s": float(evals[1]), "rounds": mc.engine.rounds,
                "cuda_graph": mc.engine.graph_used,
                "schedule": "per-chain, no barriers; finished chains masked out of the ODE launch",
                "posterior_mean_r0": float(s["strains_0_r0"].mean()),
                "posterior_mean_infectious_period": float(s["strains_0_infectious_period"].mean())}

    # untimed warm-up run: lazy CUDA module loading, cuSOLVER / cuBLAS handles and the vmap traces of the model
    # are one-time costs of the process (1-3 s, more than a whole 4096-chain run), not of a NUTS run
    MCMC(NUTS(m.model_fused, max_tree_depth=10), num_warmup=20, num_samples=5, num_chains=512,
         progress_bar=False).run(PRNGKey(1 + rank), config=cfg, tf=100, obs_data=obs)
    out["sampler"] = sampler(args.nuts_chains)
    if args.nuts_chains != 4096:
        out["sampler_4096_chains"] = sampler(4096)
    # BASELINE.json configs[1] literally: 4 chains, 500 warm-up + 100 samples, max_tree_depth 10
    # (examples/sir_infer_parameters.py:92-98 with num_chains=4): pure latency -- four trajectories per launch
    mc4 = MCMC(NUTS(m.model_fused, max_tree_depth=10), num_warmup=500, num_samples=100, num_chains=4,
               progress_bar=False)
    barrier()
    t0 = time.perf_counter()
    mc4.run(PRNGKey(8675314 + rank), config=cfg, tf=100, obs_data=obs)
    torch.cuda.synchronize()
    dt4 = time.perf_counter() - t0
    s4 = mc4.get_samples()
    out["sampler_config2_4_chains"] = {
        "value": float(mc4.engine.grad_evals) / dt4, "chains_per_gpu": 4, "num_warmup": 500, "num_samples": 100,
        "wall_s": dt4, "rounds": mc4.engine.rounds, "us_per_round": 1e6 * dt4 / max(1, mc4.engine.rounds),
        "posterior_mean_r0": float(s4["strains_0_r0"].mean()),
        "posterior_mean_infectious_period": float(s4["strains_0_infectious_period"].mean())}
    out["value"] = out["sampler"]["value"]
    return out


def run_c5(args):
    """BASELINE.json configs[4]: age x risk x multi-strain SEIRS NUTS inference, `--c5-chains` chains STRONG-scaled over
    the ranks, then posterior-predictive trajectories from every rank's draws gathered with NCCL.

    A step = one complete NUTS run (warm-up + sampling) of all chains through the public `MCMC(NUTS(model))` API
    (reference src/dynode/infer/inference.py:149-163); no collective inside it.  value = gradient evaluations of the
    whole job per second of the slowest rank.  e2e adds what a user does next: posterior samples to the host,
    `Predictive` on a thinned subset (inference.py:225-235), `gather_draws` over NVLink, gathered draws to the host."""
    import torch
    import torch.distributed as dist

    from dynode_b200.distributed import gather_draws
    from dynode_b200.examples import seirs_age_risk_strain as m5
    from dynode_b200.infer import MCMC, NUTS, Predictive, PRNGKey

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def all_ranks(xs):  # [world][len(xs)]
        t = torch.tensor([float(x) for x in xs], dtype=torch.float64, device=dev)
        if world == 1:
            return [t.tolist()]
        out = torch.empty((world, len(xs)), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, t)
        return out.tolist()

    tf = 120
    total = args.c5_chains
    lo, hi = rank * total // world, (rank + 1) * total // world
    chains = hi - lo
    obs = m5.synthetic_incidence(tf).to(dev)
    cfg = m5.get_config(infer=True)

    def nuts_run(num_warmup, num_samples, seed):
        mc = MCMC(NUTS(m5.model_fused, max_tree_depth=args.c5_tree_depth), num_warmup=num_warmup,
                  num_samples=num_samples, num_chains=chains, progress_bar=False,
                  sync_every=int(os.environ.get("DYNODE_BENCH_SYNC_EVERY", "4")))
        mc.run(PRNGKey(seed + rank), config=cfg, tf=tf, obs_data=obs)
        return mc

    def predictive_and_gather(mc, seed):
        post = mc.get_samples()
        n = next(iter(post.values())).shape[0]
        per_rank = max(1, args.c5_predictive // world)
        pick = torch.linspace(0, n - 1, min(n, per_rank), device=dev).long()
        thin = {k: v[pick] for k, v in post.items()}
        t0 = time.perf_counter()
        pp = Predictive(m5.model, posterior_samples=thin)(PRNGKey(seed + rank), config=cfg, tf=tf, obs_data=None)
        torch.cuda.synchronize()
        t_pred = time.perf_counter() - t0
        barrier()
        t0 = time.perf_counter()
        allpp = gather_draws({"incidence": pp["incidence"], **thin})
        torch.cuda.synchronize()
        t_gather = time.perf_counter() - t0
        return post, allpp, t_pred, t_gather

    # ---- untimed warm-up steps: short runs of the same program (module loading, vmap traces, graph capture, NCCL
    # channels for the gather) -- one-time costs of the process, not of a NUTS run
    for w in range(max(args.warmup, 1)):
        mcw = nuts_run(40, 4, 1000 + 17 * w)  # long enough to go through a mass-matrix adaptation window
        predictive_and_gather(mcw, 5 + w)
    barrier()
    sampler_clk = ClockSampler(local)
    sampler_clk.start()
    walls, evals, rounds, e2e_walls, per_rank, timings = [], [], [], [], [], []
    d2h_bytes = 0
    for k in range(args.steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        mc = nuts_run(args.c5_warmup, args.c5_samples, 8675314 + 1000 * k)
        ev1.record()
        torch.cuda.synchronize()
        t_dev = ev0.elapsed_time(ev1) * 1e-3
        post, allpp, t_pred, t_gather = predictive_and_gather(mc, 31 + k)
        host = {name: v.cpu() for name, v in post.items()}
        host_pp = {name: v.cpu() for name, v in allpp.items()} if rank == 0 else {}
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        d2h_bytes = sum(v.numel() * v.element_size() for v in host.values()) + \
            sum(v.numel() * v.element_size() for v in host_pp.values())
        walls.append(reduce_max(t_dev))
        e2e_walls.append(reduce_max(t_e2e))
        evals.append(reduce_sum(float(mc.engine.grad_evals)))
        rounds.append(mc.engine.rounds)
        per_rank.append(all_ranks([t_dev, float(mc.engine.rounds), float(mc.engine.grad_evals)]))
        timings.append(dict(mc.timing))
    clocks = sampler_clk.stop()
    barrier()
    mean_wall = sum(walls) / len(walls)
    value = sum(evals) / sum(walls)
    leap = float(mc.get_extra_fields()["num_steps"].double().mean()) if "num_steps" in mc.get_extra_fields() else None
    if rank == 0:
        names_r0 = [f"strains_{k}_r0" for k in range(3)]
        names_inf = [f"strains_{k}_infectious_period" for k in range(3)]
        line = {
            "metric": "NUTS grad-evals/s", "value": value, "unit": "grad-evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * mean_wall,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS["c5"][2], "chains_total": total, "chains_per_gpu": chains,
                       "num_warmup": args.c5_warmup, "num_samples": args.c5_samples,
                       "max_tree_depth": args.c5_tree_depth, "tf": tf, "rtol": 1e-5, "atol": 1e-6,
                       "l2": "not applicable: a step is a whole NUTS run (thousands of dependent launches), "
                             "inputs are 6 numbers per chain"},
            "clocks": clocks,
            "gpu_launches": int(sum(rounds) * 3),
            "e2e": {"value": sum(evals) / sum(e2e_walls), "unit": "grad-evals/s", "ms_per_step": 1e3 * sum(e2e_walls) / len(e2e_walls),
                    "h2d_bytes_per_step": int(obs.numel() * 8), "d2h_bytes_per_step": int(d2h_bytes),
                    "api": "MCMC(NUTS(model)).run + get_samples().cpu() + Predictive + gather_draws + .cpu()"},
            "nuts": {
                "grad_evals_per_s": value, "mcmc_wall_s": mean_wall, "rounds": rounds,
                # every rank's own run: a rank is done when its slowest chain is, the job when the slowest rank is
                "per_rank": [{"wall_s": [r[0] for r in step], "rounds": [int(r[1]) for r in step],
                              "us_per_round": [1e6 * r[0] / max(r[1], 1.0) for r in step],
                              "grad_evals": [int(r[2]) for r in step]} for step in per_rank],
                "grad_evals_per_run": evals, "mean_leapfrogs_per_transition": leap,
                "cuda_graph": mc.engine.graph_used, "cuda_round_kernels": mc.engine.kernels_used,
                "graph_recaptures": mc.engine.recaptures,
                "run_anatomy_rank0": [{k: round(v, 4) for k, v in t.items()} for t in timings],
                "posterior_mean_r0": [float(allpp[k].mean()) for k in names_r0],
                "posterior_mean_infectious_period": [float(allpp[k].mean()) for k in names_inf],
                "truth_r0": list(m5.TRUE_R0), "truth_infectious_period": list(m5.TRUE_INF),
                "predictive_s": t_pred, "gather_s": t_gather,
                "posterior_predictive_gathered_shape": list(allpp["incidence"].shape),
                "gather_bytes": int(allpp["incidence"].numel() * 8),
            },
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_ours(args):
    import torch
    import torch.distributed as dist

    from dynode_b200 import _lib, engine
    from dynode_b200.config import SolverParams
    from dynode_b200.examples import rhs as ex
    from dynode_b200.simulation import simulate_ensemble

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    name, B, desc = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    case = make_inputs(args.workload, B, 20260101 + rank)
    model = case["model"]
    t1 = case["t1"]
    opts = engine.SolverOptions(t1=t1)
    save_ts_h = np.linspace(0.0, t1, int(t1 // 1) + 1)
    T, n = len(save_ts_h), model.state_size
    ns = n
    # ---- device-resident inputs
    params_d = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in case["params"].items()}
    y0_d = torch.as_tensor(case["y0"], dtype=torch.float64, device=dev)
    contact_d = None if case["contact"] is None else torch.as_tensor(case["contact"], dtype=torch.float64, device=dev)
    ts_d = torch.as_tensor(save_ts_h, device=dev)
    ys = torch.empty((B, T, ns), dtype=torch.float64, device=dev)
    stats = torch.empty((B, 4), dtype=torch.int32, device=dev)

    def step():
        engine.solve_ensemble(model, y0_d, params_d, contact_d, opts, save_ts_h, out=ys, stats_out=stats, B=B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- roofline probes (measured FP64 FMA and HBM write peaks on this GPU)
    L = _lib.load()
    import ctypes
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.dynode_probe_dfma(sink.data_ptr(), 2000, stream)
    torch.cuda.synchronize()
    e0.record(); flops = L.dynode_probe_dfma(sink.data_ptr(), 40000, stream); e1.record()
    torch.cuda.synchronize()
    fp64_peak_tf = flops / (e0.elapsed_time(e1) * 1e-3) / 1e12
    nw = ys.numel()
    L.dynode_probe_hbm_write(ys.data_ptr(), nw, stream)
    torch.cuda.synchronize()
    e0.record(); L.dynode_probe_hbm_write(ys.data_ptr(), nw, stream); e1.record()
    torch.cuda.synchronize()
    hbm_write_gbs = nw * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9

    # ---- warm-up, then K timed steps
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for a, b in evs:
        a.record(); step(); b.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    per_step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = evs[0][0].elapsed_time(evs[-1][1])
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    st = stats.cpu().numpy()
    assert (st[:, 0] == 0).all(), "some trajectories hit max_steps"
    n_att = int(st[:, 3].sum())
    p_in = sum(int(np.size(v)) // B for v in case["params"].values() if np.size(v) >= B)
    flops_alg, bytes_alg = algorithmic_work(n, p_in, T, ns, F_RHS[name], n_att, B)
    kern_ms = statistics.mean(per_step_ms)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # DRAM traffic per launch from the committed `ncu --set full` capture (per trajectory x B)
    traffic = traffic_src = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        if tr:
            traffic = tr["bytes_per_trajectory"] * B
            traffic_src = tr["source"]
    except Exception:
        pass
    ach_gbs = bytes_alg / (kern_ms * 1e-3) / 1e9
    ach_tf = flops_alg / (kern_ms * 1e-3) / 1e12

    # ---- end-to-end through the public API with HOST buffers (pinned): H2D + solve + D2H
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).pin_memory()
        hp = {k: pin(v) for k, v in case["params"].items()}
        G, S = model.n_groups, model.n_strains
        y0h = pin(np.broadcast_to(case["y0"], (B, n)))
        if name == "seirs_multi_a2s3":
            ode = ex.seirs_multi_strain_ode
            p = ex.SEIRS_MultiStrain_ODEParams(beta=hp["beta"], gamma=hp["gamma"], sigma=hp["sigma"], omega=hp["omega"],
                                               contact_matrix=torch.as_tensor(case["contact"]))
            sizes = [G, G * S, G * S, G * S, G * S]
            shp = [(B, G)] + [(B, G, S)] * 4
        else:
            ode = ex.seirs_ode_seasonal
            p = ex.SeasonalSEIRS_ODEParams(beta=hp["beta"], gamma=hp["gamma"], sigma=hp["sigma"], omega=hp["omega"],
                                           seasonality_params=ex.SeasonalityParams(
                                               forcing_amp=hp["season_amp"], forcing_phase=hp["season_phase"],
                                               forcing_period=hp["season_period"]))
            sizes = [1, 1, 1, 1]
            shp = [(B, 1)] * 4
        offs = np.concatenate([[0], np.cumsum(sizes)])
        # one page-locked array per compartment, as a caller holding `initial_state` tuples per draw would have
        state = tuple(y0h[:, offs[i]:offs[i + 1]].reshape(shp[i]).contiguous().pin_memory() for i in range(len(sizes)))
        from dynode_b200 import hostmem
        if args.host_alloc == "thp":  # page-locked + 2 MiB huge pages (dynode_b200/hostmem.py): the product default
            out_h = hostmem.pinned_empty((B, T, ns))
            huge = out_h._dynode_host_buffer.huge_bytes()
        else:
            out_h = torch.empty((B, T, ns), dtype=torch.float64).pin_memory()
            huge = None
        sp = SolverParams()
        k_e2e = max(1, min(args.steps, args.e2e_steps))

        def e2e_step():
            return simulate_ensemble(ode, t1, state, p, sp, batch_size=B, state_batched=True, out=out_h,
                                     host_chunk=args.host_chunk)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            sol = e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # the result the user reads: check it equals the device-resident run
        chk = float((sol.ys[4 if name == "seirs_multi_a2s3" else 3][:64].to(dev).reshape(64, T, -1)
                     - ys[:64, :, -(sizes[-1]):]).abs().max())
        h2d = int(sum(v.numel() for v in hp.values()) * 8 + y0h.numel() * 8)
        d2h = int(out_h.numel() * 8 + B * 16)
        # the host link's own ceiling, measured in this run on the same buffers: every rank copies its resident
        # `ys` into its `out_h` at once (no solve), same chunking -- what a perfect pipeline could reach on this box
        def raw_d2h():
            for lo in range(0, B, args.host_chunk):
                out_h[lo:lo + args.host_chunk].copy_(ys[lo:lo + args.host_chunk], non_blocking=True)
        raw_d2h()
        barrier()
        t0 = time.perf_counter()
        raw_d2h(); raw_d2h()
        barrier()
        dt_raw = (time.perf_counter() - t0) / 2
        if world > 1:
            t = torch.tensor([dt_raw], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_raw = float(t.item())
        ceiling = world * out_h.numel() * 8 / dt_raw / 1e9
        e2e_gbs = world * d2h * k_e2e / dt / 1e9
        e2e = {"value": world * B * k_e2e / dt, "unit": "trajectories/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": k_e2e, "ms_per_step": 1e3 * dt / k_e2e,
               "api": "dynode_b200.simulation.simulate_ensemble (page-locked host in/out, chunked solve/D2H overlap)",
               "d2h_gbs": e2e_gbs, "pcie_ceiling_gbs": ceiling, "link_fraction": e2e_gbs / ceiling,
               "pcie_ceiling_how": f"all {world} rank(s) copying their resident ys to the same host buffers at once, "
                                   "no solve, measured in this run (aggregate GB/s)",
               "host_buffer": "mmap + MADV_HUGEPAGE + cudaHostRegister" if args.host_alloc == "thp" else "cudaHostAlloc",
               "host_buffer_huge_bytes": huge,
               "max_abs_diff_vs_device_run": chk}

    # ---- all-gather of the saved trajectories (N > 1): the only collective of the path, reported beside
    # `value`, never inside it.  Rows are gathered in place into a [world*Bg, T, n] buffer over NVLink.
    gather = None
    if world > 1 and not args.no_gather:
        from dynode_b200.distributed import GatherBuffer
        Bg = min(B, args.gather_draws)
        gb = GatherBuffer(world * Bg, (T, ns), device=dev)
        gb.local.copy_(ys[:Bg])
        for _ in range(2):
            gb.all_gather()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(5):
            gb.all_gather()
        g1.record()
        barrier()
        t = torch.tensor([g0.elapsed_time(g1) / 5], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gms = float(t.item())
        gbytes = Bg * T * ns * 8
        gather = {"collective": "all_gather_into_tensor (NCCL, in place)", "draws_per_gpu": Bg,
                  "bytes_per_rank": gbytes, "ms": gms,
                  "busbw_gbs": gbytes * (world - 1) / (gms * 1e-3) / 1e9,
                  "trajectories_per_s_with_gather": world * Bg / ((ms_per_step * Bg / B + gms) * 1e-3)}
        del gb

    # ---- second BASELINE metric: NUTS grad-evals/s on config 2 (age-stratified SIR, Poisson incidence
    # likelihood, gradient w.r.t. r0 and the infectious period), fused kernel + many-chain sampler
    nuts = None
    if not args.no_nuts:
        nuts = nuts_leg(args, dev, world, barrier, fp64_peak_tf)

    # ---- CPU baseline on rank 0, N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, dt, chunks = cpu_oracle_leg(args.workload, args.cpu_seconds, 8192, 20260101)
        cpu = {"value": v, "unit": "trajectories/s", "cores": cores, "kind": "port",
               "sample": f"{chunks}x8192 draws of the same workload ({dt:.1f} s wall), C++ oracle + OpenMP"}

    if rank == 0:
        line = {
            "metric": "solved trajectories/s", "value": value, "unit": "trajectories/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "draws_per_gpu": B, "rtol": 1e-5, "atol": 1e-6, "solver": "Tsit5+PID",
                       "mean_accepted": float(st[:, 1].mean()), "mean_rejected": float(st[:, 2].mean()),
                       "l2": f"each step streams {bytes_alg / 1e9:.2f} GB of output through the 126 MB L2, "
                             "evicting the inputs between steps (inputs+outputs >> L2)"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": ach_gbs / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "kernel": "dynode::lane_solver_kernel", "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_launch": bytes_alg, "hbm_write_probe_gbs": hbm_write_gbs},
            "roofline_fp64": {"bound": "fp64_fma", "achieved": ach_tf, "peak": fp64_peak_tf, "unit": "TFLOP/s",
                              "frac": ach_tf / fp64_peak_tf, "peak_source": "dynode_probe_dfma measured in this run",
                              "algorithmic_flops_per_launch": flops_alg, "attempted_steps_total": n_att},
            "cpu_baseline": cpu,
            "gather": gather,
            "nuts": nuts,
            "wall_s_timed_region": t_wall,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 100; 2 for --workload c5)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override draws per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--host-chunk", type=int, default=2048)
    ap.add_argument("--host-alloc", default="thp", choices=["thp", "cuda"], help="page-locked output buffer kind")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU wall time of the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-nuts", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--gather-draws", type=int, default=20000, help="draws per GPU in the all-gather leg")
    ap.add_argument("--nuts-chains", type=int, default=65536, help="chains per GPU in the NUTS leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--c5-chains", type=int, default=1024, help="total NUTS chains of --workload c5 (over all GPUs)")
    ap.add_argument("--c5-warmup", type=int, default=150)
    ap.add_argument("--c5-samples", type=int, default=50)
    ap.add_argument("--c5-tree-depth", type=int, default=7)
    ap.add_argument("--c5-predictive", type=int, default=2048, help="posterior-predictive draws over all GPUs")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 2 if args.workload == "c5" else 100
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
