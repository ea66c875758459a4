"""Shared by the NUTS oracle tests: a BatchedNUTS whose random numbers come from the oracle's tape
(oracle/nuts_np.py::Tape -- entry i of chain c depends on (seed, i, c) only), so that the many-chain sampler and the
one-chain numpy restatement of numpyro's algorithm consume identical numbers in identical roles."""
import numpy as np
import torch

from dynode_b200.infer.nuts import BatchedNUTS
from oracle.nuts_np import NutsChain, Tape


def taped_sampler(seed, C, D, device="cpu"):
    class TapedNUTS(BatchedNUTS):
        _round_i = 0

        def _randn(self, *shape):
            n, _ = Tape.block(seed, self._round_i, C, D)
            return torch.as_tensor(n, device=device)

        def _rand(self, *shape):
            _, u = Tape.block(seed, self._round_i, C, D)
            self._round_i += 1
            return torch.as_tensor(u, device=device)

    return TapedNUTS


def compare_with_oracle(pg_batched, pg_single, z0, num_warmup, num_samples, *, seed, max_tree_depth, device="cpu",
                        cuda_kernels=False, atol=1e-8, **adapt):
    """Runs both; returns the largest |z - z_oracle| over the sampling transitions after asserting that every
    transition has the same tree depth and number of leapfrogs and (to atol) the same acceptance statistic."""
    C, D = z0.shape
    cls = taped_sampler(seed, C, D, device)
    eng = cls(pg_batched, max_tree_depth=max_tree_depth, cuda_graph=False, cuda_kernels=cuda_kernels, sync_every=1,
              **adapt)
    zs, stats, b = eng.run(torch.as_tensor(z0, device=device), num_warmup, num_samples)
    zs = zs.cpu().numpy()
    worst = 0.0
    for c in range(C):
        chain = NutsChain(pg_single, D, Tape(seed, c, C, D), max_tree_depth=max_tree_depth, **adapt)
        rec = chain.run(z0[c], num_warmup, num_samples)[num_warmup:]
        assert [r["tree_depth"] for r in rec] == stats["tree_depth"][c].cpu().numpy().astype(int).tolist(), c
        assert [r["num_steps"] for r in rec] == stats["num_steps"][c].cpu().numpy().astype(int).tolist(), c
        assert [bool(r["diverging"]) for r in rec] == (stats["diverging"][c].cpu().numpy() > 0).tolist(), c
        acc = np.array([r["accept_prob"] for r in rec])
        assert np.abs(acc - stats["accept_prob"][c].cpu().numpy()).max() <= atol, c
        assert abs(rec[-1]["step_size"] - float(b.eps[c])) <= atol * max(1.0, float(b.eps[c])), c
        worst = max(worst, float(np.abs(np.array([r["z"] for r in rec]) - zs[c]).max()))
    assert worst <= atol, worst
    return worst


def random_nuts_case(seed):
    """A random sampler configuration for the oracle comparisons (tests/test_gpu_fuzz.py on the CUDA round,
    tests/test_nuts_oracle.py on the torch round): (pg_batched, pg_single, z0, num_warmup, num_samples, atol, kwargs)."""
    rng = np.random.default_rng(555_000 + seed)
    D = int(rng.choice([1, 2, 3, 5, 8, 16]))
    C = int(rng.choice([1, 2, 5, 9]))
    kind = str(rng.choice(["gauss", "gauss", "student", "banana"])) if D >= 2 else "gauss"
    Q = np.linalg.qr(rng.normal(size=(D, D)))[0]
    prec = (Q * rng.uniform(0.3, 30.0, D)) @ Q.T
    prec = 0.5 * (prec + prec.T)
    mu = rng.normal(size=D)

    if kind == "gauss":
        def pg_single(z):
            d = z - mu
            return 0.5 * d @ prec @ d, prec @ d

        def pg_batched(Z):
            P, m = torch.as_tensor(prec, device=Z.device), torch.as_tensor(mu, device=Z.device)
            d = Z - m
            g = d @ P
            return 0.5 * (d * g).sum(1), g
    elif kind == "student":  # U = sum 2 log(1 + z_i^2 / 3): heavy tails, deep trees
        def pg_single(z):
            return float(np.sum(2.0 * np.log1p(z * z / 3.0))), 4.0 * z / (3.0 + z * z)

        def pg_batched(Z):
            return (2.0 * torch.log1p(Z * Z / 3.0)).sum(1), 4.0 * Z / (3.0 + Z * Z)
    else:  # banana in the first two coordinates, unit Gaussian in the rest
        def pg_single(z):
            r = z[1] - 0.5 * z[0] * z[0]
            g = z.copy()
            g[0] = z[0] / 4.0 - 2.0 * r * z[0]
            g[1] = 2.0 * r
            return 0.125 * z[0] ** 2 + r * r + 0.5 * float(np.sum(z[2:] ** 2)), g

        def pg_batched(Z):
            r = Z[:, 1] - 0.5 * Z[:, 0] ** 2
            g = Z.clone()
            g[:, 0] = Z[:, 0] / 4.0 - 2.0 * r * Z[:, 0]
            g[:, 1] = 2.0 * r
            return 0.125 * Z[:, 0] ** 2 + r * r + 0.5 * (Z[:, 2:] ** 2).sum(1), g

    z0 = rng.normal(size=(C, D))
    depth = int(rng.choice([1, 2, 4, 7, 10]))
    adapt = bool(rng.random() < 0.5)
    kw = dict(max_tree_depth=depth, dense_mass=bool(rng.random() < 0.5),
              target_accept_prob=float(rng.choice([0.6, 0.8, 0.95])))
    if adapt:
        # adaptation (dual averaging, Welford windows, the Cholesky factor of the mass matrix) feeds rounding back
        # into the step size: the draws are compared to 1e-6 after short warm-ups, 1e-3 after one that has been
        # through a mass-matrix window; the discrete decisions exactly either way
        warm = int(rng.choice([5, 25, 60]))
        atol = 1e-6 if warm < 60 else 1e-3
        if kind != "gauss":
            # nonlinear Hamiltonian dynamics amplify the last bit ~10x every 4-5 transitions (measured on the banana:
            # 1e-15 after 5 transitions, 1e-9 after 20, a flipped tree after 60): keep the comparison inside the
            # window where the two runs are still the same run
            warm = min(warm, 10)
        kw.update(adapt_mass_matrix=bool(rng.random() < 0.7))
    else:
        warm, atol = int(rng.choice([0, 10])), 1e-9
        kw.update(adapt_step_size=False, adapt_mass_matrix=False, step_size=float(rng.choice([0.05, 0.2, 0.6])))
    n_draws = int(rng.choice([5, 25])) if kind == "gauss" or not adapt else 8
    return pg_batched, pg_single, z0, warm, n_draws, atol, kw
