"""Dev: the aten ops (in order) of one model evaluation of the C2 NUTS model."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.utils._python_dispatch import TorchDispatchMode
from dynode_b200.examples import sir_infer_parameters as m
from dynode_b200.infer import ModelDensity
dev = torch.device("cuda", 0)
cfg = m.get_config(); obs = m.synthetic_incidence(100).to(dev)
md = ModelDensity(m.model_fused, (), dict(config=cfg, tf=100, obs_data=obs))
Z = torch.randn(4096, md.dim, dtype=torch.float64, device=dev)
md.potential_and_grad(Z)
import traceback
class Log(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        name = str(func)
        if not any(k in name for k in ("view", "reshape", "expand", "select", "slice", "unsqueeze", "squeeze", "detach", "alias", "permute", "transpose", "as_strided", "_unsafe_view", "t.default", "unbind", "split")):
            st = [f for f in traceback.extract_stack() if "/dynode_b200/" in f.filename]
            loc = f"{os.path.basename(st[-1].filename)}:{st[-1].lineno}" if st else "-"
            shp = tuple(out.shape) if isinstance(out, torch.Tensor) else "-"
            print(f"{name:40s} {str(shp):18s} {loc}")
        return out
with Log():
    md.potential_and_grad(Z)
