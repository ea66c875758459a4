"""Dev: the fused log-likelihood launch at a handful of chains (one warp): where does a lone warp's time go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynode_b200 import _lib, engine
from dynode_b200.examples import sir_infer_parameters as m
from dynode_b200.synthetic import make_case
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
obs = m.synthetic_incidence(100).to(dev)
case = make_case("sir_age2", B)
prm = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in case["params"].items()}
y0 = torch.as_tensor(case["y0"], dtype=torch.float64, device=dev)
contact = torch.as_tensor(case["contact"], dtype=torch.float64, device=dev)
ts = np.linspace(0.0, 100.0, 101)
opts = engine.SolverOptions(t1=100.0)
wrt = [_lib.wrt_id(_lib.P_BETA, 0), _lib.wrt_id(_lib.P_GAMMA, 0)]
f = lambda w: engine.poisson_loglik_grad(case["model"], y0, prm, contact, opts, ts, 2, obs, 0.0, wrt=w, B=B)
for w in (wrt, []):
    for _ in range(3): lp, g, st = f(w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): lp, g, st = f(w)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} P={len(w)}: {e0.elapsed_time(e1)/20*1e3:.1f} us per launch, steps {st[:,3].tolist()[:8]}")
