"""Dev: kernel-level time of the fused Poisson log-likelihood + gradient on the C2 model (1M draws)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynode_b200 import _lib, engine
from dynode_b200.examples import sir_infer_parameters as m
from tests.cases import make_case
dev = torch.device("cuda", 0)
B = 1 << 20
obs = m.synthetic_incidence(100).to(dev)
g = torch.Generator(device=dev).manual_seed(20260102)
r0 = 1.5 + torch.rand(B, dtype=torch.float64, device=dev, generator=g)
inf = 2.0 + 13.0 * torch.rand(B, dtype=torch.float64, device=dev, generator=g)
case = make_case("sir_age2", 1)
prm = {"beta": (r0 / inf).reshape(-1, 1).contiguous(), "gamma": (1.0 / inf).reshape(-1, 1).contiguous()}
y0 = torch.as_tensor(case["y0"], dtype=torch.float64, device=dev)
contact = torch.as_tensor(case["contact"], dtype=torch.float64, device=dev)
ts = np.linspace(0.0, 100.0, 101); opts = engine.SolverOptions(t1=100.0)
wrt = [_lib.wrt_id(_lib.P_BETA, 0), _lib.wrt_id(_lib.P_GAMMA, 0)]
f = lambda: engine.poisson_loglik_grad(case["model"], y0, prm, contact, opts, ts, 2, obs, 0.0, wrt=wrt, B=B)
for _ in range(3): lp, gr, st = f()
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): lp, gr, st = f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{ms:.3f} ms per launch, {B / ms / 1e3:.1f} M grad-evals/s, lp checksum {float(lp.sum()):.12e} grad checksum {float(gr.sum()):.12e}")
