"""Dev: where a config-5 NUTS round goes at few chains per GPU -- kernel latency of the adjoint and forward-mode
log-likelihood, the whole model evaluation (eager and graph-replayed), per chain count."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynode_b200 import _lib, engine
from dynode_b200.examples import seirs_age_risk_strain as m5
from dynode_b200.infer import ModelDensity
from tests.cases import make_case
dev = torch.device("cuda", 0)
obs = m5.synthetic_incidence(120).to(dev)
cfg = m5.get_config(infer=True)
md = ModelDensity(m5.model_fused, (), dict(config=cfg, tf=120, obs_data=obs))


def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


for B in (128, 256, 512, 1024, 4096):
    case = make_case("seirs_multi_g6s3", B, seed=20260105)
    obs5 = obs.reshape(120, -1).contiguous()
    prm = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in case["params"].items()}
    y0 = torch.as_tensor(case["y0"], dtype=torch.float64, device=dev)
    c5 = torch.as_tensor(case["contact"], dtype=torch.float64, device=dev)
    ts = np.linspace(0.0, 120.0, 121); o = engine.SolverOptions(t1=120.0)
    wrt = [_lib.wrt_id(_lib.P_BETA, s) for s in range(3)] + [_lib.wrt_id(_lib.P_GAMMA, s) for s in range(3)]
    t_adj = timeit(lambda: engine.poisson_loglik_adjoint(case["model"], y0, prm, c5, o, ts, 4, obs5, 0.0, B=B, cap=256))
    t_fwd = timeit(lambda: engine.poisson_loglik_grad(case["model"], y0, prm, c5, o, ts, 4, obs5, 0.0, wrt=wrt, B=B))
    t_prim = timeit(lambda: engine.poisson_loglik_grad(case["model"], y0, prm, c5, o, ts, 4, obs5, 0.0, B=B))
    Z = md.init_to_median(B) + 0.1 * torch.randn(B, md.dim, dtype=torch.float64, device=dev)
    res = {}
    for mode in ("0", "1"):
        os.environ["DYNODE_B200_ADJOINT"] = mode
        t_pg = timeit(lambda: md.potential_and_grad(Z), n=10)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3): md.potential_and_grad(Z)
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            md.potential_and_grad(Z)
        t_gr = timeit(g.replay, n=20)
        res[mode] = (t_pg, t_gr)
    os.environ.pop("DYNODE_B200_ADJOINT")
    print(f"B={B:5d}: kernel primal {t_prim:7.1f} us  forward(6 dirs) {t_fwd:7.1f} us  adjoint {t_adj:7.1f} us | "
          f"potential_and_grad eager/graph: forward {res['0'][0]:7.1f}/{res['0'][1]:7.1f} us  adjoint {res['1'][0]:7.1f}/{res['1'][1]:7.1f} us",
          flush=True)
