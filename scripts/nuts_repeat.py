"""Dev: does a second MCMC run in the same process run slower?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynode_b200.examples import sir_infer_parameters as m
from dynode_b200.infer import MCMC, NUTS, PRNGKey
dev = torch.device("cuda", 0)
cfg = m.get_config(); obs = m.synthetic_incidence(100).to(dev)
for C in [int(x) for x in sys.argv[1:]]:
    mc = MCMC(NUTS(m.model_fused, max_tree_depth=10), num_warmup=100, num_samples=50, num_chains=C, progress_bar=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    mc.run(PRNGKey(8675314), config=cfg, tf=100, obs_data=obs)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(C, f"{dt:.3f}s rounds={mc.engine.rounds} us/round={dt/mc.engine.rounds*1e6:.0f} evals/s={mc.engine.grad_evals/dt/1e6:.2f}M", flush=True)
