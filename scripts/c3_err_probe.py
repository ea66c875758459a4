import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynode_b200 import engine
from oracle import oracle as orc
from tests.cases import make_case
from tests.test_gpu_fullsize import _solve
B = 1_000_000
case = make_case("seirs_seasonal", B, seed=20260102)
ys, st = _solve(torch, engine, case, B=B)
rng = np.random.Generator(np.random.PCG64(7))
idx = np.sort(rng.choice(B, size=512, replace=False))
fam, dims, theta, shared = case["oracle"]
ref, _, rst = orc.solve(fam, dims, case["y0"], theta[idx], shared, t1=case["t1"])
got = ys[torch.as_tensor(idx, device=ys.device)].cpu().numpy()
err = np.abs(got - ref); scale = np.abs(ref).max()
rel = err / (1e-12 * scale / 1e-9 + np.abs(ref))
w = np.unravel_index(np.argmax(rel), rel.shape)
print("max rel (with atol floor)", rel.max(), "at", w, "got", got[w], "ref", ref[w], "abs err", err[w])
print("traj params", theta[idx[w[0]]], "steps", rst[w[0]])
print("count > 1e-9:", int((rel > 1e-9).sum()), "trajectories:", np.unique(np.nonzero(rel > 1e-9)[0]).size)
print("per-traj max rel top5:", np.sort(rel.max(axis=(1, 2)))[-5:])
