#!/usr/bin/env python
"""BASELINE config 5 end to end: age x risk x strain SEIRS NUTS inference, 1024 chains sharded over the ranks of
one B200 box, NCCL all-gather of the posterior-predictive trajectories.

    python scripts/run_config5.py                                  # 1 GPU, all chains on it
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_config5.py

Each rank runs its own many-chain NUTS (no collective during sampling), draws posterior-predictive incidence
for a thinned subset of its samples with ONE vmapped model pass, then all ranks all-gather those draws.
Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from dynode_b200.distributed import gather_draws  # noqa: E402
from dynode_b200.examples import seirs_age_risk_strain as m5  # noqa: E402
from dynode_b200.infer import MCMC, NUTS, Predictive, PRNGKey  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=1024, help="total chains over all ranks")
    ap.add_argument("--warmup", type=int, default=150)
    ap.add_argument("--samples", type=int, default=50)
    ap.add_argument("--tf", type=int, default=120)
    ap.add_argument("--max-tree-depth", type=int, default=7)
    ap.add_argument("--predictive-per-rank", type=int, default=256)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.zeros(1, device=dev))  # build the communicator before anything is timed
    chains = args.chains // world
    obs = m5.synthetic_incidence(args.tf).to(dev)
    cfg = m5.get_config(infer=True)
    mc = MCMC(NUTS(m5.model_fused, max_tree_depth=args.max_tree_depth), num_warmup=args.warmup,
              num_samples=args.samples, num_chains=chains, progress_bar=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mc.run(PRNGKey(8675314 + rank), config=cfg, tf=args.tf, obs_data=obs)
    torch.cuda.synchronize()
    t_mcmc = time.perf_counter() - t0
    post = mc.get_samples()
    n = next(iter(post.values())).shape[0]
    pick = torch.linspace(0, n - 1, min(n, args.predictive_per_rank), device=dev).long()
    thin = {k: v[pick] for k, v in post.items()}
    t0 = time.perf_counter()
    pp = Predictive(m5.model, posterior_samples=thin)(PRNGKey(rank), config=cfg, tf=args.tf, obs_data=None)
    torch.cuda.synchronize()
    t_pred = time.perf_counter() - t0
    t0 = time.perf_counter()
    allpp = gather_draws({"incidence": pp["incidence"], **thin})
    torch.cuda.synchronize()
    t_gather = time.perf_counter() - t0
    stats = torch.tensor([float(mc.engine.grad_evals), t_mcmc], dtype=torch.float64, device=dev)
    if world > 1:
        tot = stats.clone()
        dist.all_reduce(tot[:1], op=dist.ReduceOp.SUM)
        dist.all_reduce(tot[1:], op=dist.ReduceOp.MAX)
        stats = tot
    if rank == 0:
        line = {
            "config": "C5 age(3) x risk(2) x strain(3) SEIRS + C NUTS, Poisson on daily incidence",
            "n_gpus": world, "chains_total": chains * world, "chains_per_gpu": chains,
            "num_warmup": args.warmup, "num_samples": args.samples,
            "grad_evals_per_s": float(stats[0]) / float(stats[1]), "mcmc_wall_s": float(stats[1]),
            "rounds": mc.engine.rounds, "cuda_graph": mc.engine.graph_used, "cuda_round_kernels": mc.engine.kernels_used,
            "posterior_mean_r0": [float(allpp[f"strains_{k}_r0"].mean()) for k in range(3)],
            "posterior_mean_infectious_period": [float(allpp[f"strains_{k}_infectious_period"].mean()) for k in range(3)],
            "truth_r0": list(m5.TRUE_R0), "truth_infectious_period": list(m5.TRUE_INF),
            "posterior_predictive_gathered_shape": list(allpp["incidence"].shape),
            "predictive_s": t_pred, "gather_s": t_gather,
            "gather_bytes": int(allpp["incidence"].numel() * 8),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
