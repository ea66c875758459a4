"""Dev: run the torch round and the CUDA round in lock-step from the same state with the same random numbers
and report the first state field that differs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynode_b200.infer.nuts import BatchedNUTS
dev = torch.device("cuda", 0)
cov = torch.tensor([[1.0, 0.6, 0.0], [0.6, 2.0, -0.4], [0.0, -0.4, 0.5]], dtype=torch.float64, device=dev)
mu = torch.tensor([0.5, -1.0, 2.0], dtype=torch.float64, device=dev)
prec = torch.linalg.inv(cov)
def pg(z):
    d = z - mu; g = d @ prec
    return 0.5 * (d * g).sum(1), g
C = 96
engs = []
for kernels in (False, True):
    e = BatchedNUTS(pg, max_tree_depth=7, generator=torch.Generator(device=dev).manual_seed(11), cuda_graph=False, cuda_kernels=kernels)
    e._allocate(torch.zeros(C, 3, dtype=torch.float64, device=dev), 30)
    e._g = e.gen
    U, g = e._eval(e.b.z); e.b.U.copy_(U); e.b.g.copy_(g); e.b.need_tree.fill_(True)
    e.set_schedule([3] * 40, [40.0] * 40, 40)
    e._prepare_round_fn()
    b = e.b
    engs.append(e)
fields = ["z","U","g","eps","k","active","need_tree","energy0","zL","rL","gL","zR","rR","gR","zP","gP","r_sum","UP","weight","sum_acc","depth","nprop","turning","diverging","s_n","s_right","s_turn","s_div","s_z","s_r","s_g","s_zP","s_gP","s_rsum","s_UP","s_w","s_acc","da_x","da_xavg","da_gavg","da_t","wf_n","wf_mean","wf_m2"]
bad_chains = set()
for rnd in range(400):
    for e in engs: e._round_fn()
    a, b = engs[0].b, engs[1].b
    for f in fields:
        x, y = getattr(a, f).double(), getattr(b, f).double()
        d = (x - y).abs().reshape(C, -1).amax(1)
        tol = 1e-9 * (1 + x.abs().reshape(C, -1).amax(1))
        nb = torch.nonzero((d > tol) | torch.isnan(d) & ~(torch.isnan(x.reshape(C,-1)).any(1) & torch.isnan(y.reshape(C,-1)).any(1))).flatten().tolist()
        new = [c for c in nb if c not in bad_chains]
        if new:
            c = new[0]
            print(f"round {rnd}: field {f} differs first for chain {c}: torch={getattr(a,f)[c].tolist()} cuda={getattr(b,f)[c].tolist()}")
            for g2 in ("s_n","depth","s_w","weight","s_turn","s_div","turning","diverging","energy0","s_acc","nprop","eps","need_tree","k"):
                print("   ", g2, getattr(a,g2)[c].tolist(), getattr(b,g2)[c].tolist())
            bad_chains.update(new)
    if len(bad_chains) > 5: break
print("rounds run", rnd + 1, "bad chains", sorted(bad_chains))
