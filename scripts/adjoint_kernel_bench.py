"""Dev: kernel-level time of the discrete-adjoint log-likelihood on config 5 (n = 78), 32768 draws."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynode_b200 import engine
from dynode_b200.examples import seirs_age_risk_strain as m5
from tests.cases import make_case
dev = torch.device("cuda", 0)
B5 = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
case5 = make_case("seirs_multi_g6s3", B5, seed=20260105)
obs5 = m5.synthetic_incidence(120).to(dev).reshape(120, -1).contiguous()
prm5 = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in case5["params"].items()}
y05 = torch.as_tensor(case5["y0"], dtype=torch.float64, device=dev)
c5 = torch.as_tensor(case5["contact"], dtype=torch.float64, device=dev)
ts5 = np.linspace(0.0, 120.0, 121); o5 = engine.SolverOptions(t1=120.0)
f = lambda: engine.poisson_loglik_adjoint(case5["model"], y05, prm5, c5, o5, ts5, 4, obs5, 0.0, B=B5, cap=256)
for _ in range(2): lp, g, _, st = f()
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): lp, g, _, st = f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{ms:.3f} ms per launch, {B5 / ms / 1e3:.3f} M grad-evals/s, lp checksum {float(lp.sum()):.12e} grad checksum {float(g.sum()):.12e}")
