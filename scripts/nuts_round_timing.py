"""Dev: where does a NUTS round go at a small chain count -- host launch or device?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynode_b200.examples import sir_infer_parameters as m
from dynode_b200.infer import ModelDensity
from dynode_b200.infer.nuts import BatchedNUTS, build_transition_schedule
dev = torch.device("cuda", 0)
cfg = m.get_config(); obs = m.synthetic_incidence(100).to(dev)
md = ModelDensity(m.model_fused, (), dict(config=cfg, tf=100, obs_data=obs))
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = BatchedNUTS(md.potential_and_grad, max_tree_depth=10)
z0 = md.init_to_median(C)
eng._allocate(z0, 50); eng._g = eng.gen
U, g = eng._eval(eng.b.z); eng.b.U.copy_(U); eng.b.g.copy_(g); eng.b.need_tree.fill_(True)
fl, wl = build_transition_schedule(100000, 50, True, True)
eng.set_schedule(fl, wl, 100000)
eng._prepare_round_fn()
print("graph:", eng.graph_used)
for n_sync in (1, 4, 64):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    R = 512
    for i in range(R):
        eng._round_fn()
        if (i + 1) % n_sync == 0: bool(eng.b.any_active)
    e1.record(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"sync every {n_sync:3d}: wall {dt / R * 1e6:.0f} us/round, device span {e0.elapsed_time(e1) / R * 1e3:.0f} us/round")
t0 = time.perf_counter()
for i in range(256): eng._round_fn()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue {1e6 * (t1 - t0) / 256:.0f} us/round, then drain {1e6 * (t2 - t1) / 256:.0f} us/round")
