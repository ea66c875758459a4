"""Dev: compiled vs composed model evaluation (config 2), eager and graph-replayed, and the plan's three launches."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynode_b200 import _lib, engine
from dynode_b200.examples import sir_infer_parameters as m
from dynode_b200.infer import ModelDensity
dev = torch.device("cuda", 0)
obs = m.synthetic_incidence(100).to(dev)
md = ModelDensity(m.model_fused, (), dict(config=m.get_config(), tf=100, obs_data=obs))


def timeit(f, n=30):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def graphed(f):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): f()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        f()
    return g.replay


for C in (512, 4096, 65536):
    Z = md.init_to_median(C) + 0.3 * torch.randn(C, md.dim, dtype=torch.float64, device=dev)
    md.potential_and_grad(Z)
    assert md._plan is not None, md.plan_reason
    plan = md._plan
    t = {}
    t["plan eager"] = timeit(lambda: plan.potential_and_grad(Z))
    t["composed eager"] = timeit(lambda: md.potential_and_grad_composed(Z), n=10)
    t["plan graph"] = timeit(graphed(lambda: plan.potential_and_grad(Z)))
    t["composed graph"] = timeit(graphed(lambda: md.potential_and_grad_composed(Z)))
    # the three launches alone
    L = _lib.load()
    theta = torch.empty((C, plan.K), dtype=torch.float64, device=dev)
    aux = torch.empty((C, 3 * plan.D + 1), dtype=torch.float64, device=dev)
    st = ctypes.c_void_p(_lib.current_stream_ptr())
    t["pre"] = timeit(lambda: L.dynode_potential_pre_f64(ctypes.byref(plan.c_plan), C, Z.data_ptr(), Z.stride(0), theta.data_ptr(), aux.data_ptr(), None, st))
    cfg, pl = plan.cfg, plan.cfg.payload
    from dynode_b200.simulation import autograd as ag
    prm = ag._kernel_params(cfg, theta)
    f = lambda: engine.poisson_loglik_grad(cfg.model, plan.y0, prm, pl.contact, cfg.opts(), pl.save_ts, pl.obs_comp, pl.obs, pl.lp_const, wrt=cfg.wrt_ids(), B=C, zero_masked=False)
    t["loglik"] = timeit(f)
    t["loglik graph"] = timeit(graphed(f))
    print(f"C={C}: " + "  ".join(f"{k} {v:.1f} us" for k, v in t.items()), flush=True)
