"""Dev: wall time of SVI on config 2 (reference examples/sir_infer_parameters.py through SVIProcess) with the compiled
potential against the composed (vmap + autograd) evaluation."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynode_b200.examples import sir_infer_parameters as m
from dynode_b200.infer import SVIProcess
from dynode_b200.infer.model_density import ModelDensity
dev = torch.device("cuda", 0)
obs = m.synthetic_incidence(100).to(dev)
composed = "--composed" in sys.argv
if composed:
    ModelDensity.potential_and_grad = ModelDensity.potential_and_grad_composed
for rep in range(3):
    proc = SVIProcess(numpyro_model=m.model_fused, num_iterations=500, num_samples=200, progress_bar=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    proc.infer(config=m.get_config(), tf=100, obs_data=obs)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    post = proc.get_samples()
    r0 = float(post["strains_0_r0"].mean()); ip = float(post["strains_0_infectious_period"].mean())
    print(f"{'composed' if composed else 'compiled'}: {t:.3f} s for 500 iterations ({1e3 * t / 500:.2f} ms each)  r0 {r0:.3f}  infectious period {ip:.3f}")
