"""Dev timing of the NUTS path on one GPU (not the bench): potential_and_grad per call, graph vs eager MCMC."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynode_b200.examples import sir_infer_parameters as m
from dynode_b200.infer import ModelDensity, MCMC, NUTS, PRNGKey
dev = torch.device("cuda", 0)
cfg = m.get_config(); obs = m.synthetic_incidence(100).to(dev)
for name, mod in (("general", m.model), ("fused", m.model_fused)):
    md = ModelDensity(mod, (), dict(config=cfg, tf=100, obs_data=obs))
    for C in (1024, 65536):
        Z = torch.randn(C, 2, dtype=torch.float64, device=dev)
        for _ in range(3): md.potential_and_grad(Z)
        torch.cuda.synchronize(); t = time.perf_counter(); n = 10
        for _ in range(n): md.potential_and_grad(Z)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / n
        print(f"{name} C={C}: {dt*1e3:.2f} ms per call, {C/dt:.3g} grad-evals/s", flush=True)
for graph in (True, False):
    for C in (1024, 16384):
        if not graph and C > 1024: continue
        mc = MCMC(NUTS(m.model_fused, max_tree_depth=10), num_warmup=200, num_samples=100, num_chains=C,
                  progress_bar=False, cuda_graph=graph)
        torch.cuda.synchronize(); t = time.perf_counter()
        mc.run(PRNGKey(5), config=cfg, tf=100, obs_data=obs)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        e = mc.engine
        print(f"mcmc graph={e.graph_used} C={C}: {dt:.2f} s, rounds {e.rounds}, tree grad-evals {e.grad_evals} "
              f"({e.grad_evals/dt:.3g}/s), launched {e.launched_evals} ({e.launched_evals/dt:.3g}/s), "
              f"{dt/e.rounds*1e6:.0f} us/round", flush=True)
        mc.print_summary()
