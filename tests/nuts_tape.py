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
