"""Dev: CUDA kernels launched by one model evaluation (potential + gradient) of the C2 NUTS model, and by one round."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from dynode_b200.examples import sir_infer_parameters as m
from dynode_b200.infer import ModelDensity
dev = torch.device("cuda", 0)
cfg = m.get_config(); obs = m.synthetic_incidence(100).to(dev)
md = ModelDensity(m.model_fused, (), dict(config=cfg, tf=100, obs_data=obs))
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Z = torch.randn(C, md.dim, dtype=torch.float64, device=dev)
for _ in range(3): md.potential_and_grad(Z)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    md.potential_and_grad(Z)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
cnt = collections.Counter(e.name[:90] for e in ev)
tot = sum(e.device_time for e in ev) if hasattr(ev[0], "device_time") else sum(e.cuda_time for e in ev)
print("kernels:", len(ev), "sum of kernel time (us):", tot)
for k, v in cnt.most_common(40): print(v, k)
