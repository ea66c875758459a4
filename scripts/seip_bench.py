"""Dev: throughput of the CTA-per-trajectory SEIP kernel (A=4 ages, K=3 strains, W=4 waning stages, n=416)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynode_b200 import seip
from dynode_b200.engine import SolverOptions
from tests.cases import make_seip_case
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
case = make_seip_case(B, A=4, K=3, W=4, t1=365)
m = case["model"]
prm = {k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in case["params"].items()}
y0 = torch.as_tensor(case["y0"], dtype=torch.float64, device=dev)
C, pop, imm = (torch.as_tensor(case[k], dtype=torch.float64, device=dev) for k in ("contact", "pop", "immunity"))
ts = np.linspace(0.0, 365.0, 366)
out = torch.empty((B, 366, m.state_size), dtype=torch.float64, device=dev)
o = SolverOptions(t1=365.0)
step = lambda: seip.solve_ensemble(m, y0, prm, C, pop, imm, o, ts, out=out, B=B)
for _ in range(3): ys, st = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ys, st = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
n = m.state_size
natt = float(st[:, 3].double().sum())
bytes_alg = B * (8 * (n + 13) + 8 * 366 * n + 16)
print(json.dumps({"kernel": "seip_solver_kernel", "B": B, "n": n, "ms_per_launch": ms, "trajectories_per_s": B / ms * 1e3,
                  "mean_attempted_steps": natt / B, "hbm_gbs": bytes_alg / ms / 1e6, "hbm_frac_of_6539.9": bytes_alg / ms / 1e6 / 6539.9}))
