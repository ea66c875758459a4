"""Dev: which chains set the number of rounds, and where do their leapfrogs go?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynode_b200.examples import sir_infer_parameters as m
from dynode_b200.infer import MCMC, NUTS, PRNGKey
dev = torch.device("cuda", 0)
cfg = m.get_config(); obs = m.synthetic_incidence(100).to(dev)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mc = MCMC(NUTS(m.model_fused, max_tree_depth=10), num_warmup=100, num_samples=50, num_chains=C, progress_bar=False)
mc.run(PRNGKey(8675314), config=cfg, tf=100, obs_data=obs)
b = mc.engine.b
nl = b.n_leap.double()
print("rounds", mc.engine.rounds, "n_leap mean", float(nl.mean()), "median", float(nl.median()), "p99", float(nl.quantile(0.99)), "p999", float(nl.quantile(0.999)), "max", float(nl.max()))
samp = b.out_stats["num_steps"].sum(1)  # leapfrogs in the 50 sampling transitions
print("sampling-phase leapfrogs: mean", float(samp.mean()), "max", float(samp.max()))
top = torch.topk(nl, 8).indices
for c in top.tolist():
    print(f"chain {c}: total {int(nl[c])}, sampling {int(samp[c])}, eps {float(b.eps[c]):.4g}, mean accept (sampling) {float(b.out_stats['accept_prob'][c].mean()):.3f}, "
          f"imm diag {[round(float(x), 4) for x in torch.diagonal(b.imm[c])]}, z {[round(float(x), 3) for x in b.z[c]]}")
print("typical chain eps", float(b.eps.median()), "imm diag median", [round(float(x), 4) for x in torch.diagonal(b.imm, dim1=1, dim2=2).median(0).values])
