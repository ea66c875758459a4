"""One-chain NUTS in numpy, written after numpyro 0.15 -- the sampler the reference runs through
`MCMC(NUTS(model, dense_mass=True, max_tree_depth, init_strategy=init_to_median))`
(reference src/dynode/infer/inference.py:149-163).  TEST INFRASTRUCTURE ONLY: the checker of the many-chain CUDA sampler
(dynode_b200/infer/nuts.py, csrc/nuts_round.cu), never imported by the product.

PARITY UNPINNED against numpyro itself (numpyro is third-party, pinned `0.15.*` in the reference's pyproject.toml:17,
absent from /root/reference and from this image).  What is restated, function by function, from its published
algorithm (numpyro/infer/hmc_util.py):

    velocity_verlet                 one leapfrog
    find_reasonable_step_size       doubling / halving search around exp(-dE) = 0.8 with fresh momenta
    dual_averaging                  t0 = 10, kappa = 0.75, gamma = 0.05, prox centre log(10 * step size)
    welford_covariance              Welford update; Stan regularisation (n/(n+5)) cov + 1e-3 (5/(n+5)) I at window ends
    build_adaptation_schedule       Stan's 75 / 25-doubling / 50 warm-up windows
    warmup_adapter                  update_fn / _update_at_window_end order of operations
    build_tree, _double_tree, _iterative_build_subtree, _build_basetree, _combine_tree,
    _leaf_idx_to_ckpt_idxs, _is_iterative_turning, _is_turning,
    _uniform_transition_kernel, _biased_transition_kernel          the iterative NUTS tree, max_delta_energy = 1000

It is written in numpyro's own shape -- a sequential chain that builds whole subtrees, with TreeInfo records -- not in
the sampler's (one leapfrog of every chain per round, state machines in two kernels), so the two are independent
statements of one algorithm.  Two conventions are the sampler's, because numpyro's cannot be reproduced without its
PRNG: random numbers come from a TAPE indexed by the chain's leapfrog count (`tape.normal(i)` for a momentum refresh
at leapfrog i, `tape.uniform(i)[0..2]` for the direction of a doubling that starts at leapfrog i, the in-subtree
transition of leaf i and the top-level transition of a subtree that ends at leaf i); and momenta are drawn as
L^-T xi with inverse mass matrix L L^T (numpyro uses cholesky_of_inverse: another square root of the same mass
matrix, i.e. the same law for the momentum).
"""
from __future__ import annotations

import math
from collections import namedtuple

import numpy as np

MAX_DELTA_ENERGY = 1000.0

TreeInfo = namedtuple("TreeInfo", "z_left r_left g_left z_right r_right g_right z_proposal pe_proposal g_proposal "
                                  "depth weight r_sum turning diverging sum_accept_probs num_proposals")


class Tape:
    """Counter-based random numbers: entry i depends on (seed, i, chain) only."""

    def __init__(self, seed: int, chain: int, n_chains: int, dim: int):
        self.seed, self.chain, self.C, self.D = seed, chain, n_chains, dim

    @staticmethod
    def block(seed: int, i: int, n_chains: int, dim: int):
        g = np.random.Generator(np.random.Philox(key=[seed, i]))
        return g.standard_normal((n_chains, dim)), g.random((n_chains, 3))

    def normal(self, i):
        return self.block(self.seed, i, self.C, self.D)[0][self.chain]

    def uniform(self, i):
        return self.block(self.seed, i, self.C, self.D)[1][self.chain]


def build_adaptation_schedule(num_steps):
    """numpyro.infer.hmc_util.build_adaptation_schedule: [(start, end)] inclusive."""
    adaptation_schedule = []
    if num_steps < 20:
        adaptation_schedule.append((0, num_steps - 1))
        return adaptation_schedule
    start_buffer_size, end_buffer_size, init_window_size = 75, 50, 25
    if (start_buffer_size + end_buffer_size + init_window_size) > num_steps:
        start_buffer_size = int(0.15 * num_steps)
        end_buffer_size = int(0.1 * num_steps)
        init_window_size = num_steps - start_buffer_size - end_buffer_size
    adaptation_schedule.append((0, start_buffer_size - 1))
    end_window_start = num_steps - end_buffer_size
    next_window_size = init_window_size
    next_window_start = start_buffer_size
    while next_window_start < end_window_start:
        cur_window_start, cur_window_size = next_window_start, next_window_size
        if 3 * cur_window_size <= end_window_start - cur_window_start:
            next_window_size = 2 * cur_window_size
        else:
            cur_window_size = end_window_start - cur_window_start
        next_window_start = cur_window_start + cur_window_size
        adaptation_schedule.append((cur_window_start, next_window_start - 1))
    adaptation_schedule.append((end_window_start, num_steps - 1))
    return adaptation_schedule


def _leaf_idx_to_ckpt_idxs(n):
    """numpyro: idx_max = number of set bits of n >> 1; idx_min = idx_max - (trailing set bits of n) + 1."""
    idx_max = bin(n >> 1).count("1")
    num_subtrees, m = 0, n
    while m & 1:
        num_subtrees += 1
        m >>= 1
    return idx_max - num_subtrees + 1, idx_max


class NutsChain:
    def __init__(self, potential_and_grad, dim, tape, max_tree_depth=10, target_accept_prob=0.8, dense_mass=True,
                 step_size=1.0, adapt_step_size=True, adapt_mass_matrix=True):
        self.pg, self.D, self.tape = potential_and_grad, dim, tape
        self.max_depth, self.target, self.dense = max_tree_depth, target_accept_prob, dense_mass
        self.adapt_ss, self.adapt_mm = adapt_step_size, adapt_mass_matrix
        self.step_size = float(step_size)
        self.imm = np.eye(dim)       # inverse mass matrix
        self.msqrt = np.eye(dim)     # square root of the mass matrix used for momentum draws
        self.i = 0                   # leapfrogs done so far = tape position

    # ------------------------------------------------------------------ dynamics
    def _eval(self, z):
        U, g = self.pg(z)
        U, g = float(U), np.asarray(g, dtype=np.float64)
        if not (np.isfinite(U) and np.all(np.isfinite(g))):
            return math.inf, np.zeros(self.D)
        return U, g

    def _kinetic(self, r):
        return 0.5 * float(r @ (self.imm @ r))

    def _leapfrog(self, z, r, g, eps):
        """velocity_verlet update with identity-free dense metric."""
        r_half = r - 0.5 * eps * g
        z_new = z + eps * (self.imm @ r_half)
        U_new, g_new = self._eval(z_new)
        r_new = r_half - 0.5 * eps * g_new
        self.i += 1
        return z_new, r_new, U_new, g_new

    def find_reasonable_step_size(self, step_size, z, U, g):
        direction = 0
        tiny, huge = np.finfo(np.float64).tiny, np.finfo(np.float64).max
        while True:
            step_size = (2.0 ** direction) * step_size
            r = self.msqrt @ self.tape.normal(self.i)
            energy_current = U + self._kinetic(r)
            _, r_new, U_new, _ = self._leapfrog(z, r, g, step_size)
            delta_energy = U_new + self._kinetic(r_new) - energy_current
            direction_new = 1 if math.log(0.8) < -delta_energy else -1  # NaN compares false: -1
            last_direction, direction = direction, direction_new
            not_extreme = (step_size > tiny or direction >= 0) and (step_size < huge or direction <= 0)
            if not (not_extreme and (last_direction == 0 or direction == last_direction)):
                return step_size

    # ------------------------------------------------------------------ the tree
    def _is_turning(self, r_left, r_right, r_sum):
        v_left, v_right = self.imm @ r_left, self.imm @ r_right
        r_sum = r_sum - (r_left + r_right) / 2
        return bool(v_left @ r_sum <= 0) or bool(v_right @ r_sum <= 0)

    def _build_basetree(self, z, r, g, eps_signed, energy_current):
        z_new, r_new, pe_new, g_new = self._leapfrog(z, r, g, eps_signed)
        energy_new = pe_new + self._kinetic(r_new)
        delta_energy = energy_new - energy_current
        if math.isnan(delta_energy):
            delta_energy = math.inf
        diverging = delta_energy > MAX_DELTA_ENERGY
        accept_prob = min(math.exp(-delta_energy), 1.0) if delta_energy > -700 else 1.0
        return TreeInfo(z_new, r_new, g_new, z_new, r_new, g_new, z_new, pe_new, g_new, 0, -delta_energy, r_new,
                        False, diverging, accept_prob, 1)

    def _combine_tree(self, current, new, going_right, u, biased):
        if going_right:
            z_left, r_left, g_left = current.z_left, current.r_left, current.g_left
            z_right, r_right, g_right = new.z_right, new.r_right, new.g_right
        else:
            z_left, r_left, g_left = new.z_left, new.r_left, new.g_left
            z_right, r_right, g_right = current.z_right, current.r_right, current.g_right
        r_sum = current.r_sum + new.r_sum
        if biased:
            p = 0.0 if (new.turning or new.diverging) else min(math.exp(min(new.weight - current.weight, 0.0)), 1.0) \
                if new.weight - current.weight < 0 else 1.0
            if new.turning or new.diverging:
                p = 0.0
            turning = new.turning or self._is_turning(r_left, r_right, r_sum)
        else:
            d = new.weight - current.weight
            p = 1.0 / (1.0 + math.exp(-d)) if d > -700 else 0.0  # expit
            turning = current.turning
        take = u < p
        prop = new if take else current
        return TreeInfo(z_left, r_left, g_left, z_right, r_right, g_right, prop.z_proposal, prop.pe_proposal,
                        prop.g_proposal, current.depth + 1, float(np.logaddexp(current.weight, new.weight)), r_sum,
                        turning, new.diverging, current.sum_accept_probs + new.sum_accept_probs,
                        current.num_proposals + new.num_proposals)

    def _iterative_build_subtree(self, prototype, going_right, eps, energy_current):
        max_num_proposals = 2 ** prototype.depth
        r_ckpts = np.zeros((self.max_depth, self.D))
        r_sum_ckpts = np.zeros((self.max_depth, self.D))
        tree, turning = None, False
        num = 0
        last_leaf_index = self.i
        while num < max_num_proposals and not turning and not (tree is not None and tree.diverging):
            if tree is None:  # start from the edge of the current (prototype) tree
                z, r, g = (prototype.z_right, prototype.r_right, prototype.g_right) if going_right else \
                    (prototype.z_left, prototype.r_left, prototype.g_left)
            else:
                z, r, g = (tree.z_right, tree.r_right, tree.g_right) if going_right else \
                    (tree.z_left, tree.r_left, tree.g_left)
            last_leaf_index = self.i
            u = self.tape.uniform(self.i)
            leaf = self._build_basetree(z, r, g, eps if going_right else -eps, energy_current)
            if tree is None:
                new_tree = leaf
            else:
                new_tree = self._combine_tree(tree, leaf, going_right, u[1], biased=False)
            leaf_idx = num
            idx_min, idx_max = _leaf_idx_to_ckpt_idxs(leaf_idx)
            r_leaf = leaf.r_right  # the leaf's momentum (left == right for a base tree)
            if leaf_idx % 2 == 0 and idx_max < self.max_depth:
                r_ckpts[idx_max] = r_leaf
                r_sum_ckpts[idx_max] = new_tree.r_sum
            turning = False
            i = idx_max
            while i >= idx_min and not turning:
                if 0 <= i < self.max_depth:
                    subtree_r_sum = new_tree.r_sum - r_sum_ckpts[i] + r_ckpts[i]
                    turning = self._is_turning(r_ckpts[i], r_leaf, subtree_r_sum)
                i -= 1
            tree = new_tree
            num = tree.num_proposals
        return tree._replace(depth=prototype.depth, turning=turning), last_leaf_index

    def build_tree(self, z, U, g, eps):
        r = self.msqrt @ self.tape.normal(self.i)
        energy_current = U + self._kinetic(r)
        tree = TreeInfo(z, r, g, z, r, g, z, U, g, 0, 0.0, r, False, False, 0.0, 0)
        while tree.depth < self.max_depth and not tree.turning and not tree.diverging:
            going_right = bool(self.tape.uniform(self.i)[0] < 0.5)
            new_tree, last_leaf = self._iterative_build_subtree(tree, going_right, eps, energy_current)
            tree = self._combine_tree(tree, new_tree, going_right, self.tape.uniform(last_leaf)[2], biased=True)
        return tree

    # ------------------------------------------------------------------ the chain
    def run(self, z0, num_warmup, num_samples):
        """Returns per-transition records: z, accept_prob, num_steps, tree_depth, diverging, step_size (the one the
        NEXT transition uses), potential_energy -- for all num_warmup + num_samples transitions."""
        z = np.asarray(z0, dtype=np.float64).copy()
        U, g = self._eval(z)
        eps = self.step_size
        sched = build_adaptation_schedule(num_warmup) if num_warmup > 0 else []
        num_windows = len(sched)
        window_idx = 0
        # warmup_adapter.init
        if self.adapt_ss and num_warmup > 0:
            eps = self.find_reasonable_step_size(eps, z, U, g)
        prox, x_t, x_avg, g_avg, t_da = math.log(10.0 * eps), 0.0, 0.0, 0.0, 0.0
        wf_n, wf_mean, wf_m2 = 0.0, np.zeros(self.D), np.zeros((self.D, self.D))
        out = []
        for t in range(num_warmup + num_samples):
            tree = self.build_tree(z, U, g, eps)
            accept = tree.sum_accept_probs / max(tree.num_proposals, 1)
            z, U, g = tree.z_proposal, tree.pe_proposal, tree.g_proposal
            if t < num_warmup:
                # ---- warmup_adapter.update_fn
                if self.adapt_ss:
                    t_da += 1.0
                    g_avg = (1.0 - 1.0 / (t_da + 10.0)) * g_avg + (self.target - accept) / (t_da + 10.0)
                    x_t = prox - math.sqrt(t_da) / 0.05 * g_avg
                    w = t_da ** (-0.75)
                    x_avg = (1.0 - w) * x_avg + w * x_t
                    eps = math.exp(x_avg) if t == num_warmup - 1 else math.exp(x_t)
                is_middle_window = 0 < window_idx < num_windows - 1
                if self.adapt_mm and is_middle_window:
                    wf_n += 1.0
                    delta_pre = z - wf_mean
                    wf_mean = wf_mean + delta_pre / wf_n
                    delta_post = z - wf_mean
                    wf_m2 = wf_m2 + np.outer(delta_post, delta_pre)
                t_at_window_end = t == sched[window_idx][1]
                if t_at_window_end:
                    window_idx += 1
                if t_at_window_end and is_middle_window:
                    # ---- _update_at_window_end
                    if self.adapt_mm and wf_n > 1:
                        cov = wf_m2 / (wf_n - 1.0)
                        scaled_cov = (wf_n / (wf_n + 5.0)) * cov
                        shrinkage = 1e-3 * (5.0 / (wf_n + 5.0))
                        cov = scaled_cov + shrinkage * np.eye(self.D)
                        if not self.dense:
                            cov = np.diag(np.diag(cov))
                        self.imm = cov
                        L = np.linalg.cholesky(cov)
                        self.msqrt = np.linalg.inv(L).T  # M = imm^-1 = L^-T L^-1
                    wf_n, wf_mean, wf_m2 = 0.0, np.zeros(self.D), np.zeros((self.D, self.D))
                    if self.adapt_ss:
                        eps = self.find_reasonable_step_size(eps, z, U, g)
                        prox, x_t, x_avg, g_avg, t_da = math.log(10.0 * eps), 0.0, 0.0, 0.0, 0.0
            out.append(dict(z=z.copy(), accept_prob=accept, num_steps=tree.num_proposals, tree_depth=tree.depth,
                            diverging=tree.diverging, step_size=eps, potential_energy=U))
        self.step_size = eps
        return out
