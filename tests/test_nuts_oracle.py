"""The many-chain NUTS (dynode_b200/infer/nuts.py) against an independent one-chain restatement of numpyro's
algorithm (oracle/nuts_np.py: build_tree / _iterative_build_subtree / dual averaging / Welford windows /
find_reasonable_step_size, written in numpyro's sequential shape), transition by transition on a shared tape of
random numbers.  CPU: the sampler's torch round.  The CUDA round kernels meet the same oracle in
tests/test_gpu_infer.py."""
import numpy as np
import pytest
import torch

from oracle import nuts_np
from tests.nuts_tape import compare_with_oracle

A = np.array([[2.0, 0.6, 0.1], [0.6, 1.0, -0.3], [0.1, -0.3, 0.5]])
PREC, MU = np.linalg.inv(A), np.array([0.5, -1.0, 2.0])


def pg_single(z):
    d = z - MU
    return 0.5 * d @ PREC @ d, PREC @ d


def pg_batched(Z):
    P, mu = torch.as_tensor(PREC, device=Z.device), torch.as_tensor(MU, device=Z.device)
    d = Z - mu
    return 0.5 * torch.einsum("ci,ij,cj->c", d, P, d), d @ P.T


Z0 = np.random.default_rng(1).normal(size=(4, 3))


def test_adaptation_schedule_restatements_agree():
    from dynode_b200.infer import build_adaptation_schedule
    for n in (0, 5, 19, 20, 30, 100, 150, 151, 500, 1000):
        if n == 0:
            continue
        assert nuts_np.build_adaptation_schedule(n) == build_adaptation_schedule(n), n


def test_checkpoint_index_rule():
    # numpyro _leaf_idx_to_ckpt_idxs: leaf 0 -> (1, 0) (no test), leaf 1 -> (0, 0), leaf 3 -> (0, 1), leaf 7 -> (0, 2)
    assert [nuts_np._leaf_idx_to_ckpt_idxs(n) for n in (0, 1, 2, 3, 5, 7)] == \
        [(1, 0), (0, 0), (2, 1), (0, 1), (1, 1), (0, 2)]


def test_full_adaptation_short_run_matches_transition_by_transition():
    """30 warm-up transitions cover every adaptation event (initial step-size search, a slow window's Welford
    covariance -> regularised inverse mass matrix, the step-size search that follows it, dual averaging with its
    restart, the averaged final step size), then 20 draws.  Rounding differences between numpy and torch are
    amplified by the step-size feedback by ~15 % per transition, hence the short horizon and 1e-8."""
    compare_with_oracle(pg_batched, pg_single, Z0, 30, 20, seed=77, max_tree_depth=6, atol=1e-8)


def test_long_run_without_step_size_feedback_is_identical_to_rounding():
    """120 warm-up (three slow windows of mass-matrix adaptation) + 40 draws at a fixed step size: nothing amplifies
    rounding, and the two statements agree to 1e-12 in every draw."""
    worst = compare_with_oracle(pg_batched, pg_single, Z0, 120, 40, seed=5, max_tree_depth=6, atol=1e-12,
                                adapt_step_size=False, step_size=0.5)
    assert worst < 1e-12


def test_long_run_keeps_the_same_tree_shapes():
    """With every adaptation on, 150 + 30 transitions: draws may drift apart by amplified rounding, but as long as they
    have not, the discrete decisions (tree depth, leapfrogs, divergences) are the same."""
    compare_with_oracle(pg_batched, pg_single, Z0[:2], 150, 30, seed=77, max_tree_depth=6, atol=5e-2)


def test_oracle_chain_recovers_gaussian_moments():
    tape = nuts_np.Tape(3, 0, 1, 3)
    chain = nuts_np.NutsChain(pg_single, 3, tape, max_tree_depth=8)
    rec = chain.run(np.zeros(3), 300, 1500)[300:]
    z = np.array([r["z"] for r in rec])
    assert np.allclose(z.mean(0), MU, atol=0.15)
    assert np.allclose(np.cov(z.T), A, atol=0.35)
    assert 0.6 < np.mean([r["accept_prob"] for r in rec]) < 0.95


@pytest.mark.parametrize("seed", range(12))
def test_random_configurations_on_the_torch_round(seed):
    """tests/test_gpu_fuzz.py's random sampler configurations (targets, chain counts, depth limits, mass matrix kind,
    adaptation on / off) through the masked-tensor torch round on the CPU."""
    from tests.nuts_tape import compare_with_oracle, random_nuts_case
    pg_batched, pg_single, z0, warm, n_draws, atol, kw = random_nuts_case(seed)
    compare_with_oracle(pg_batched, pg_single, z0, warm, n_draws, seed=1000 + seed, atol=atol, **kw)
