"""Dev: kernel-time breakdown of NUTS rounds (graph replay) early and late in a run."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from dynode_b200.examples import sir_infer_parameters as m
from dynode_b200.infer import ModelDensity
from dynode_b200.infer.nuts import BatchedNUTS, build_transition_schedule
dev = torch.device("cuda", 0)
cfg = m.get_config(); obs = m.synthetic_incidence(100).to(dev)
md = ModelDensity(m.model_fused, (), dict(config=cfg, tf=100, obs_data=obs))
C = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
eng = BatchedNUTS(md.potential_and_grad, max_tree_depth=10)
eng._allocate(md.init_to_median(C), 50); eng._g = eng.gen
U, g = eng._eval(eng.b.z); eng.b.U.copy_(U); eng.b.g.copy_(g); eng.b.need_tree.fill_(True)
fl, wl = build_transition_schedule(100, 50, True, True)
eng.set_schedule(fl, wl, 100)
eng._prepare_round_fn()
done = 3
for target in (300, 3000, 6000):
    while done < target:
        eng._round_fn(); done += 1
    torch.cuda.synchronize()
    act = float(eng.b.active.double().mean())
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10): eng._round_fn()
        torch.cuda.synchronize()
    done += 10
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    tot = collections.Counter(); cnt = collections.Counter()
    for e in ev:
        k = e.name.split("<")[0].split("(")[0][-60:]
        tot[k] += e.device_time; cnt[k] += 1
    print(f"--- round {target}: active fraction {act:.3f}; kernel time per round {sum(tot.values())/10:.0f} us in {len(ev)/10:.0f} launches")
    for k, v in tot.most_common(6): print(f"   {v/10:8.1f} us  x{cnt[k]/10:.0f}  {k}")
