"""Dev check: does the NUTS round capture into a CUDA graph, and do graph / eager runs agree statistically?"""
import os, sys, time, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynode_b200.examples import sir_infer_parameters as m
from dynode_b200.infer import MCMC, NUTS, PRNGKey
dev = torch.device("cuda", 0)
cfg = m.get_config(); obs = m.synthetic_incidence(100).to(dev)
for mod in (m.model_fused, m.model):
    for C in (256, 4096):
        mc = MCMC(NUTS(mod, max_tree_depth=8), num_warmup=150, num_samples=100, num_chains=C, progress_bar=False, cuda_graph=True)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            torch.cuda.synchronize(); t = time.perf_counter()
            mc.run(PRNGKey(5), config=cfg, tf=100, obs_data=obs)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
        for x in w: print("WARNING:", x.message)
        e = mc.engine
        print(f"{mod.__name__} graph={e.graph_used} kernels={e.kernels_used} C={C}: {dt:.2f} s, rounds {e.rounds}, tree grad-evals {e.grad_evals} "
              f"({e.grad_evals/dt:.3g}/s), {dt/e.rounds*1e6:.0f} us/round", flush=True)
        mc.print_summary()
