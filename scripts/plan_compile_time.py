"""Dev: what compiling the potential plan costs (it runs inside the first evaluation of every ModelDensity)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynode_b200.infer import ModelDensity, MCMC, NUTS, PRNGKey
from dynode_b200.infer.potential_plan import compile_plan
dev = torch.device("cuda", 0)
for which in ("c2", "c5", "c2", "c5"):
    if which == "c2":
        from dynode_b200.examples import sir_infer_parameters as m
        kw = dict(config=m.get_config(), tf=100, obs_data=m.synthetic_incidence(100).to(dev))
    else:
        from dynode_b200.examples import seirs_age_risk_strain as m
        kw = dict(config=m.get_config(infer=True), tf=120, obs_data=m.synthetic_incidence(120).to(dev))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    md = ModelDensity(m.model_fused, (), kw, device=dev)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    plan, why = compile_plan(md)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{which}: ModelDensity {1e3*(t1-t0):.1f} ms, compile_plan {1e3*(t2-t1):.1f} ms, plan={'ok' if plan else why}", flush=True)
# r_hat of the config-2 fit with and without the compiled evaluation (64 chains x 100 draws, as the GPU test)
from dynode_b200.examples import sir_infer_parameters as m
kw = dict(config=m.get_config(), tf=100, obs_data=m.synthetic_incidence(100).to(dev))
for plan_on in ("1", "0", "1", "0"):
    os.environ["DYNODE_B200_PLAN"] = plan_on
    for seed in (0, 1, 2):
        mc = MCMC(NUTS(m.model_fused, max_tree_depth=6), num_warmup=150, num_samples=100, num_chains=64, progress_bar=False)
        t0 = time.perf_counter()
        mc.run(PRNGKey(seed), **kw)
        dt = time.perf_counter() - t0
        s = mc.summary()
        print(f"plan={plan_on} seed={seed}: r_hat r0 {s['strains_0_r0']['r_hat']:.4f} inf {s['strains_0_infectious_period']['r_hat']:.4f} "
              f"mean r0 {float(mc.get_samples()['strains_0_r0'].mean()):.4f}  wall {dt:.2f} s rounds {mc.engine.rounds}", flush=True)
