"""BASELINE.json's full sizes (C4: 100k multi-strain age-stratified SEIRS draws, C3: 1M seasonal SEIRS draws,
365 days, daily SaveAt) through properties that do not need the oracle to integrate the whole ensemble:
conservation laws of the flow family, `ys[0] == y0` exactly, monotone cumulative incidence, the stats identity,
schedule independence (a permutation of the draws permutes the outputs bit for bit; a shard of the ensemble
solved alone equals its rows of the full launch -- what multi-GPU sharding relies on), plus the oracle itself on
a random sample of the ensemble at the parity bar of tests/test_gpu_parity.py."""
import numpy as np
import pytest

from tests.cases import make_case

pytestmark = pytest.mark.gpu


def _solve(torch, engine, case, idx=None, B=None):
    dev = torch.device("cuda", 0)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    B = case["y0"].shape[0] if (B is None and np.ndim(case["y0"]) == 2) else B
    sel = (lambda a: a) if idx is None else (lambda a: a[idx] if (np.ndim(a) >= 1 and np.shape(a)[0] == B) else a)
    prm = {k: t(sel(np.asarray(v))) for k, v in case["params"].items()}
    y0 = t(sel(case["y0"]))
    contact = None if case["contact"] is None else t(case["contact"])
    t1 = float(case["t1"])
    ts = np.linspace(0.0, t1, int(t1) + 1)
    n_rows = B if idx is None else len(idx)
    ys, _, st = engine.solve_ensemble(case["model"], y0, prm, contact, engine.SolverOptions(t1=t1), ts, B=n_rows)
    return ys, st


@pytest.mark.parametrize("name,B", [("seirs_multi_a2s3", 100_000), ("seirs_seasonal", 1_000_000)])
def test_full_size_ensemble_properties(name, B):
    import torch
    from dynode_b200 import engine
    from oracle import oracle as orc

    case = make_case(name, B, seed=20260103 if name == "seirs_multi_a2s3" else 20260102)
    model = case["model"]
    G, S = model.n_groups, model.n_strains
    ys, st = _solve(torch, engine, case, B=B)
    T, n = ys.shape[1], ys.shape[2]
    assert ys.shape == (B, 366, model.state_size)
    # ---- stats: every trajectory finished, accepted + rejected == steps
    assert int((st[:, 0] != 0).sum()) == 0
    assert torch.equal(st[:, 1] + st[:, 2], st[:, 3])
    assert bool(torch.isfinite(ys).all())
    # ---- ys[0] == y0 exactly (reference tests/test_simulation/test_odes.py:63-74)
    y0 = torch.as_tensor(np.broadcast_to(case["y0"], (B, n)).copy(), device=ys.device)
    assert torch.equal(ys[:, 0, :], y0)
    # ---- people are conserved per population group (cumulative incidence excluded), to rounding
    sizes = model.compartment_sizes()
    offs = np.concatenate([[0], np.cumsum(sizes)])
    ncons = 4 if model.n_compartments >= 4 else 3  # s, (e,) i, r
    pop = ys[:, :, offs[0]:offs[1]].clone()
    for c in range(1, ncons):
        pop += ys[:, :, offs[c]:offs[c + 1]].reshape(B, T, G, S).sum(-1)
    drift = ((pop - pop[:, :1]) / pop[:, :1]).abs().max()
    assert float(drift) < 1e-10, f"population drift {float(drift):.2e}"
    # ---- compartments stay non-negative up to the tolerance; cumulative incidence never decreases
    assert float(ys.min()) > -1e-5 * float(ys.max())
    if model.n_compartments == 5:
        c = ys[:, :, offs[4]:offs[5]]
        assert float((c[:, 1:] - c[:, :-1]).min()) > -1e-9 * float(c.max())
    # ---- the oracle on a random sample of the ensemble: 1e-9 relative and identical step counts
    rng = np.random.Generator(np.random.PCG64(7))
    idx = np.sort(rng.choice(B, size=512, replace=False))
    fam, dims, theta, shared = case["oracle"]
    y0s = case["y0"][idx] if np.ndim(case["y0"]) == 2 else case["y0"]
    ref, _, rst = orc.solve(fam, dims, y0s, theta[idx], shared, t1=case["t1"])
    got = ys[torch.as_tensor(idx, device=ys.device)].cpu().numpy()
    assert np.array_equal(st[torch.as_tensor(idx, device=st.device)].cpu().numpy(), rst)
    scale = np.abs(ref).max()
    rel = (np.abs(got - ref) / (1e-3 * scale + np.abs(ref))).max(axis=(1, 2))  # per trajectory
    # 1e-9 is the bar of tests/test_gpu_parity.py; on samples this large a few ill-conditioned trajectories
    # amplify rounding-level differences further -- the oracle compiled with and without FMA contraction differs
    # from ITSELF by 1e-9..1e-8 on exactly those (C3 draw 194 of this sample: 8.8e-9 kernel vs oracle, same order
    # oracle vs oracle) -- so: 99 % within 1e-9, all within 1e-7 (BASELINE's bar is 1e-6)
    assert np.quantile(rel, 0.99) <= 1e-9, f"99% quantile {np.quantile(rel, 0.99):.2e}"
    assert rel.max() <= 1e-7, f"max {rel.max():.2e}"
    # ---- a shard solved alone == its rows of the full launch (rank g of G owns a contiguous block)
    lo, hi = B // 2 - 1234, B // 2 + 4321
    sub = np.arange(lo, hi)
    ys_s, st_s = _solve(torch, engine, case, idx=sub, B=B)
    assert torch.equal(ys_s, ys[lo:hi]) and torch.equal(st_s, st[lo:hi])
    del ys_s
    # ---- schedule independence: permuting the draws permutes the outputs, bit for bit
    digest = ys.view(torch.int64).sum(dim=(1, 2))  # a checksum per trajectory (wrapping integer sum of the bits)
    del ys, pop
    torch.cuda.empty_cache()
    perm = rng.permutation(B)
    ys_p, st_p = _solve(torch, engine, case, idx=perm, B=B)
    p = torch.as_tensor(perm, device=ys_p.device)
    assert torch.equal(ys_p.view(torch.int64).sum(dim=(1, 2)), digest[p])
    assert torch.equal(st_p, st[p])
