"""Dev: fused log-likelihood gradient, forward sensitivities vs discrete adjoint, kernel time per launch."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynode_b200 import _lib, engine
from tests.cases import make_case
dev = torch.device("cuda", 0)
def timeit(fn, n=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("| case | B | directions | forward ms | adjoint ms | primal-only ms |")
print("|---|---|---|---|---|---|")
for name, comp, B, t1 in (("sir_age2", 2, 1 << 18, 100), ("seirs_seasonal", 3, 1 << 18, 120), ("seirs_multi_a2s3", 4, 1 << 16, 120), ("seirs_multi_g6s3", 4, 1 << 14, 120), ("seirs_multi_g6s3", 4, 128, 120)):
    case = make_case(name, B)
    model = case["model"]; S = model.n_strains
    prm = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in case["params"].items()}
    y0 = torch.as_tensor(np.broadcast_to(case["y0"], (B, model.state_size)).copy(), dtype=torch.float64, device=dev)
    K = None if case["contact"] is None else torch.as_tensor(case["contact"], dtype=torch.float64, device=dev)
    ts = np.linspace(0.0, t1, t1 + 1)
    m = model.compartment_sizes()[comp]
    obs = torch.rand(t1, m, dtype=torch.float64, device=dev) + 0.5
    o = engine.SolverOptions(t1=float(t1))
    kinds = [_lib.P_BETA, _lib.P_GAMMA] + ([_lib.P_SIGMA, _lib.P_OMEGA] if model.flow != _lib.FLOW_SIR else [])
    for nk in (1, 2, len(kinds)):
        wrt = [_lib.wrt_id(k, s) for k in kinds[:nk] for s in range(S)]
        tf = timeit(lambda: engine.poisson_loglik_grad(model, y0, prm, K, o, ts, comp, obs, 0.0, wrt=wrt, B=B))
        ta = timeit(lambda: engine.poisson_loglik_adjoint(model, y0, prm, K, o, ts, comp, obs, 0.0, B=B, cap=256))
        t0 = timeit(lambda: engine.poisson_loglik_grad(model, y0, prm, K, o, ts, comp, obs, 0.0, wrt=[], B=B))
        print(f"| {name} | {B} | {len(wrt)} | {tf:.3f} | {ta:.3f} | {t0:.3f} |", flush=True)
