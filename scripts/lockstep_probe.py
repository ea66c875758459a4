"""Dev probe: how much of a launch is lost to lock-step (slots of one warp waiting for the slowest)?
Times the kernel on random draws and on the same draws replicated `rep` times consecutively (every slot of
a warp then integrates the same trajectory), and reports ms per million attempted steps for both."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynode_b200 import engine
from tests.cases import make_case

def run(name, B, rep, t1=365.0, iters=10):
    dev = torch.device("cuda", 0)
    case = make_case(name, B, seed=20260104)
    model = case["model"]
    idx = np.arange(B) if rep == 1 else np.repeat(np.arange(B // rep), rep)[:B]
    params = {k: torch.as_tensor(np.asarray(v)[idx] if np.ndim(v) >= 1 and np.shape(v)[0] == B else v, dtype=torch.float64, device=dev)
              for k, v in case["params"].items()}
    y0 = case["y0"]
    y0 = torch.as_tensor(y0[idx] if np.ndim(y0) == 2 else y0, dtype=torch.float64, device=dev)
    contact = None if case["contact"] is None else torch.as_tensor(case["contact"], dtype=torch.float64, device=dev)
    ts = np.linspace(0.0, t1, int(t1) + 1)
    ys = torch.empty((B, len(ts), model.state_size), dtype=torch.float64, device=dev)
    stats = torch.empty((B, 4), dtype=torch.int32, device=dev)
    opts = engine.SolverOptions(t1=t1)
    f = lambda: engine.solve_ensemble(model, y0, params, contact, opts, ts, out=ys, stats_out=stats, B=B)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    att = int(stats[:, 3].sum().item())
    return {"case": name, "B": B, "rep": rep, "ms": ms, "attempted_steps": att, "ms_per_Mstep": ms / att * 1e6,
            "mean_att": att / B, "max_att": int(stats[:, 3].max().item())}

if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--B", type=int, default=50000)
    ap.add_argument("--rep", type=int, default=1)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    if a.case:
        print(json.dumps(run(a.case, a.B, a.rep, iters=a.iters)), flush=True)
        sys.exit(0)
    for name, B, reps in (("seirs_multi_a2s3", 100000, (1, 5, 160)), ("seirs_seasonal", 1000000, (1, 32, 160))):
        for rep in reps:
            print(json.dumps(run(name, B, rep)), flush=True)
