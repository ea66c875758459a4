"""Tiny invocation of every kernel family, meant to run under `compute-sanitizer --tool memcheck`."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynode_b200 import _lib, engine, seip
from dynode_b200.engine import SolverOptions
from tests.cases import make_case, make_seip_case
dev = torch.device("cuda", 0)
for name in ("seirs_multi_a2s3", "seirs_seasonal", "sir_age2", "seirs_multi_g6s3"):
    case = make_case(name, 37)
    t1 = 60
    ts = np.linspace(0.0, t1, t1 + 1)
    ys, _, st = engine.solve_ensemble(case["model"], case["y0"], case["params"], case["contact"], SolverOptions(t1=t1), ts)
    ys, dys, st = engine.solve_ensemble(case["model"], case["y0"], case["params"], case["contact"], SolverOptions(t1=t1), ts, wrt=[0, 16], save_mask=0b101)
    ys, _, st = engine.solve_ensemble(case["model"], case["y0"], case["params"], case["contact"], SolverOptions(t1=t1, jump_ts=(10.0, 33.3)), ts)
    m = case["model"].compartment_sizes()[2]
    obs = np.random.default_rng(0).uniform(0.5, 2.0, (t1, m))
    engine.poisson_loglik_grad(case["model"], case["y0"], case["params"], case["contact"], SolverOptions(t1=t1), ts, 2, obs, 0.0, wrt=[0, 16])
    engine.poisson_loglik_adjoint(case["model"], case["y0"], case["params"], case["contact"], SolverOptions(t1=t1), ts, 2, obs, 0.0, with_y0_grad=True, cap=64)
    torch.cuda.synchronize()
    print("ok", name, flush=True)
c = make_seip_case(9, A=3, K=2, W=3, t1=40)
seip.solve_ensemble(c["model"], c["y0"], c["params"], c["contact"], c["pop"], c["immunity"], SolverOptions(t1=40.0), np.linspace(0, 40, 41))
c = make_seip_case(5, A=2, K=4, W=5, t1=40)
seip.solve_ensemble(c["model"], c["y0"], c["params"], c["contact"], c["pop"], c["immunity"], SolverOptions(t1=40.0), np.linspace(0, 40, 41))
torch.cuda.synchronize(); print("ok seip", flush=True)
from dynode_b200.infer.nuts import BatchedNUTS
def pg(z): return 0.5 * (z * z).sum(1), z
eng = BatchedNUTS(pg, max_tree_depth=5, cuda_graph=False, cuda_kernels=True, generator=torch.Generator(device=dev).manual_seed(1))
eng.run(torch.zeros(33, 3, dtype=torch.float64, device=dev), 20, 10)
torch.cuda.synchronize(); print("ok nuts", flush=True)
