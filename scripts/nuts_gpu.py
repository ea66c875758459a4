import time, torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynode_b200.examples import sir_infer_parameters as m
from dynode_b200.infer import ModelDensity, MCMCProcess
dev = torch.device("cuda",0)
cfg = m.get_config(); obs = m.synthetic_incidence(100).to(dev)
for name, mod in (("general", m.model), ("fused", m.model_fused)):
    md = ModelDensity(mod, (), dict(config=cfg, tf=100, obs_data=obs))
    for C in (64, 4096, 65536):
        Z = torch.randn(C, 2, dtype=torch.float64, device=dev)
        for _ in range(3): md.potential_and_grad(Z)
        torch.cuda.synchronize(); t=time.perf_counter(); n=10
        for _ in range(n): md.potential_and_grad(Z)
        torch.cuda.synchronize(); dt=(time.perf_counter()-t)/n
        print(f"{name} C={C}: {dt*1e3:.2f} ms per call, {C/dt:.3g} grad-evals/s")
proc = MCMCProcess(numpyro_model=m.model_fused, num_warmup=200, num_samples=100, num_chains=1024, nuts_max_tree_depth=10, progress_bar=True)
t=time.perf_counter(); proc.infer(config=cfg, tf=100, obs_data=obs); torch.cuda.synchronize(); dt=time.perf_counter()-t
print("mcmc time", dt, "grad evals", proc._inferer.engine.grad_evals, proc._inferer.engine.grad_evals/dt, "per s")
proc._inferer.print_summary()
