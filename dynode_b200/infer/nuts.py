"""Many-chain NUTS on the device.

The reference runs numpyro's `MCMC(NUTS(model, dense_mass=True, max_tree_depth, init_to_median))`
(reference src/dynode/infer/inference.py:149-163); numpyro is absent here and, more to the point, its
chains run one after another on a CPU.  This sampler keeps numpyro's algorithm -- iterative tree doubling
with checkpointed U-turn tests, multinomial proposal sampling with a biased top-level transition,
divergence at an energy error of 1000, dual-averaging step size (target 0.8, t0=10, kappa=0.75,
gamma=0.05), Stan's windowed dense mass-matrix adaptation with Welford covariance and shrinkage -- but
advances ALL chains in lock-step: every leapfrog round is ONE batched evaluation of
`potential_and_grad(z[C, D])`, i.e. one ensemble launch of the ODE kernel for C chains.  Chain state lives
in device tensors; per-chain control flow is masks, not Python branches.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional, Tuple

import torch

MAX_DELTA_ENERGY = 1000.0


def build_adaptation_schedule(num_steps: int):
    """Stan's warmup windows [(start, end)], inclusive ends: a fast initial buffer, doubling slow windows
    in which the mass matrix is estimated, a fast terminal buffer."""
    if num_steps < 20:
        return [(0, num_steps - 1)]
    start_buffer, end_buffer, window = 75, 50, 25
    if start_buffer + end_buffer + window > num_steps:
        start_buffer = int(0.15 * num_steps)
        end_buffer = int(0.1 * num_steps)
        window = num_steps - start_buffer - end_buffer
    sched = [(0, start_buffer - 1)]
    end_win_start = num_steps - end_buffer
    start = start_buffer
    while start < end_win_start:
        nxt = start + window
        if nxt + 2 * window > end_win_start:  # the last slow window absorbs the remainder
            nxt = end_win_start
        sched.append((start, nxt - 1))
        start = nxt
        window *= 2
    sched.append((end_win_start, num_steps - 1))
    return sched


def _tree_index_tables(max_depth: int, device):
    """For leaf index n inside a subtree: which checkpoint slots take part in the U-turn tests
    (numpyro `_leaf_idx_to_ckpt_idxs`): idx_max = popcount(n >> 1), idx_min = idx_max - trailing_ones(n) + 1."""
    size = 1 << max_depth
    idx_max = torch.zeros(size, dtype=torch.long)
    idx_min = torch.zeros(size, dtype=torch.long)
    for n in range(size):
        mx = bin(n >> 1).count("1")
        ones, m = 0, n
        while m & 1:
            ones += 1
            m >>= 1
        idx_max[n], idx_min[n] = mx, mx - ones + 1
    return idx_min.to(device), idx_max.to(device)


@dataclass
class NUTSState:
    z: torch.Tensor        # [C, D] unconstrained position
    U: torch.Tensor        # [C] potential energy
    g: torch.Tensor        # [C, D] gradient of U
    step_size: torch.Tensor  # [C]
    inv_mass: torch.Tensor   # [C, D, D]
    i: int = 0
    # dual averaging
    da_x: torch.Tensor = None
    da_xavg: torch.Tensor = None
    da_gavg: torch.Tensor = None
    da_t: torch.Tensor = None
    da_prox: torch.Tensor = None
    # Welford accumulators of the current slow window
    wf_n: int = 0
    wf_mean: torch.Tensor = None
    wf_m2: torch.Tensor = None
    stats: Dict[str, torch.Tensor] = field(default_factory=dict)


class BatchedNUTS:
    def __init__(self, potential_and_grad: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                 max_tree_depth: int = 10, target_accept_prob: float = 0.8, dense_mass: bool = True,
                 step_size: float = 1.0, adapt_step_size: bool = True, adapt_mass_matrix: bool = True,
                 generator: Optional[torch.Generator] = None):
        self.pg = potential_and_grad
        self.max_depth = int(max_tree_depth)
        self.target = float(target_accept_prob)
        self.dense = dense_mass
        self.init_step = float(step_size)
        self.adapt_ss, self.adapt_mm = adapt_step_size, adapt_mass_matrix
        self.gen = generator
        self.grad_evals = 0  # batched evaluations x chains
        self._tables = None

    # ------------------------------------------------------------------ helpers
    def _randn(self, *shape, like):
        return torch.randn(*shape, dtype=like.dtype, device=like.device, generator=self.gen)

    def _rand(self, *shape, like):
        return torch.rand(*shape, dtype=like.dtype, device=like.device, generator=self.gen)

    def _eval(self, z):
        U, g = self.pg(z)
        self.grad_evals += z.shape[0]
        bad = ~torch.isfinite(U) | ~torch.isfinite(g).all(dim=1)
        U = torch.where(bad, torch.full_like(U, math.inf), U)
        g = torch.where(bad[:, None], torch.zeros_like(g), g)
        return U, g

    @staticmethod
    def _vel(inv_mass, r):
        return torch.einsum("cij,cj->ci", inv_mass, r)

    def _kinetic(self, inv_mass, r):
        return 0.5 * (r * self._vel(inv_mass, r)).sum(1)

    def _is_turning(self, inv_mass, r_left, r_right, r_sum):
        v_l, v_r = self._vel(inv_mass, r_left), self._vel(inv_mass, r_right)
        rc = r_sum - 0.5 * (r_left + r_right)
        return ((v_l * rc).sum(-1) <= 0) | ((v_r * rc).sum(-1) <= 0)

    # ------------------------------------------------------------------ initialisation
    def init(self, z0: torch.Tensor) -> NUTSState:
        C, D = z0.shape
        U, g = self._eval(z0)
        if not torch.isfinite(U).all():
            raise RuntimeError("cannot find valid initial parameters: the potential is not finite at the "
                               "initial position of some chain")
        eye = torch.eye(D, dtype=z0.dtype, device=z0.device).expand(C, D, D).contiguous()
        ss = torch.full((C,), self.init_step, dtype=z0.dtype, device=z0.device)
        st = NUTSState(z=z0.clone(), U=U, g=g, step_size=ss, inv_mass=eye)
        self._da_reset(st)
        self._wf_reset(st)
        self._tables = _tree_index_tables(self.max_depth, z0.device)
        return st

    def _da_reset(self, st: NUTSState):
        zeros = torch.zeros_like(st.step_size)
        st.da_prox = torch.log(10.0 * st.step_size)
        st.da_x, st.da_xavg, st.da_gavg, st.da_t = zeros.clone(), zeros.clone(), zeros.clone(), zeros.clone()

    def _wf_reset(self, st: NUTSState):
        C, D = st.z.shape
        st.wf_n = 0
        st.wf_mean = torch.zeros((C, D), dtype=st.z.dtype, device=st.z.device)
        st.wf_m2 = torch.zeros((C, D, D), dtype=st.z.dtype, device=st.z.device)

    # ------------------------------------------------------------------ one transition for all chains
    def step(self, st: NUTSState) -> NUTSState:
        C, D = st.z.shape
        dev, dt = st.z.device, st.z.dtype
        md = self.max_depth
        imm, eps = st.inv_mass, st.step_size
        idx_min_tab, idx_max_tab = self._tables
        ar = torch.arange(C, device=dev)
        lvl = torch.arange(md, device=dev)

        # momentum r ~ N(0, M), M = inv_mass^-1: r = L^-T xi with inv_mass = L L^T
        L = torch.linalg.cholesky(imm)
        r0 = torch.linalg.solve_triangular(L.transpose(1, 2), self._randn(C, D, 1, like=st.z), upper=True)[..., 0]
        energy0 = st.U + self._kinetic(imm, r0)

        # whole-tree state
        zL, rL, gL = st.z.clone(), r0.clone(), st.g.clone()
        zR, rR, gR = st.z.clone(), r0.clone(), st.g.clone()
        zP, UP, gP = st.z.clone(), st.U.clone(), st.g.clone()
        depth = torch.zeros(C, dtype=torch.long, device=dev)
        weight = torch.zeros(C, dtype=dt, device=dev)
        r_sum = r0.clone()
        turning = torch.zeros(C, dtype=torch.bool, device=dev)
        diverging = torch.zeros(C, dtype=torch.bool, device=dev)
        sum_acc = torch.zeros(C, dtype=dt, device=dev)
        nprop = torch.zeros(C, dtype=torch.long, device=dev)
        active = torch.ones(C, dtype=torch.bool, device=dev)

        # subtree under construction
        s_n = torch.zeros(C, dtype=torch.long, device=dev)  # leaves so far
        s_right = torch.zeros(C, dtype=torch.bool, device=dev)
        s_z, s_r, s_g = st.z.clone(), r0.clone(), st.g.clone()           # moving edge
        s_z0, s_r0, s_g0 = st.z.clone(), r0.clone(), st.g.clone()        # first leaf (inner edge)
        s_zP, s_UP, s_gP = st.z.clone(), st.U.clone(), st.g.clone()
        s_w = torch.zeros(C, dtype=dt, device=dev)
        s_rsum = torch.zeros(C, D, dtype=dt, device=dev)
        s_turn = torch.zeros(C, dtype=torch.bool, device=dev)
        s_div = torch.zeros(C, dtype=torch.bool, device=dev)
        s_acc = torch.zeros(C, dtype=dt, device=dev)
        r_ck = torch.zeros(C, md, D, dtype=dt, device=dev)
        rs_ck = torch.zeros(C, md, D, dtype=dt, device=dev)

        def sel(m, a, b):
            return torch.where(m.reshape(-1, *([1] * (a.dim() - 1))), a, b)

        rounds = 0
        while True:
            if not bool(active.any()):
                break
            rounds += 1
            # ---- chains starting a new doubling: pick a direction, start from that edge of the tree
            start = active & (s_n == 0)
            go_right = self._rand(C, like=st.z) < 0.5
            s_right = torch.where(start, go_right, s_right)
            s_z = sel(start, sel(s_right, zR, zL), s_z)
            s_r = sel(start, sel(s_right, rR, rL), s_r)
            s_g = sel(start, sel(s_right, gR, gL), s_g)

            # ---- one leapfrog for every chain (inactive chains are evaluated too and discarded)
            h = torch.where(s_right, eps, -eps)[:, None]
            r_half = s_r - 0.5 * h * s_g
            z_new = s_z + h * self._vel(imm, r_half)
            U_new, g_new = self._eval(z_new)
            r_new = r_half - 0.5 * h * g_new
            delta = U_new + self._kinetic(imm, r_new) - energy0
            delta = torch.where(torch.isnan(delta), torch.full_like(delta, math.inf), delta)
            leaf_w = -delta
            leaf_div = delta > MAX_DELTA_ENERGY
            leaf_acc = torch.clamp(torch.exp(-delta), max=1.0)

            # ---- fold the leaf into the subtree (uniform/multinomial transition inside a subtree)
            first = s_n == 0
            new_w = torch.where(first, leaf_w, torch.logaddexp(s_w, leaf_w))
            p_take = torch.where(first, torch.ones_like(leaf_w), torch.sigmoid(leaf_w - s_w))
            take = active & (self._rand(C, like=st.z) < p_take)
            s_zP, s_UP, s_gP = sel(take, z_new, s_zP), torch.where(take, U_new, s_UP), sel(take, g_new, s_gP)
            new_rsum = torch.where(first[:, None], r_new, s_rsum + r_new)
            upd = active
            s_z0, s_r0, s_g0 = sel(upd & first, z_new, s_z0), sel(upd & first, r_new, s_r0), sel(upd & first, g_new, s_g0)
            s_z, s_r, s_g = sel(upd, z_new, s_z), sel(upd, r_new, s_r), sel(upd, g_new, s_g)
            s_w = torch.where(upd, new_w, s_w)
            s_rsum = sel(upd, new_rsum, s_rsum)
            s_acc = torch.where(upd, torch.where(first, leaf_acc, s_acc + leaf_acc), s_acc)
            s_div = torch.where(upd, leaf_div, s_div)

            # ---- checkpointed U-turn tests over the sub-subtrees that this leaf completes
            n = s_n
            i_min, i_max = idx_min_tab[n], idx_max_tab[n]
            even = (n % 2 == 0) & upd
            slot = torch.zeros(C, md, dtype=torch.bool, device=dev)
            slot[ar, i_max.clamp(max=md - 1)] = even
            r_ck = torch.where(slot[:, :, None], r_new[:, None, :], r_ck)
            rs_ck = torch.where(slot[:, :, None], s_rsum[:, None, :], rs_ck)
            sub_sum = s_rsum[:, None, :] - rs_ck + r_ck                     # [C, md, D]
            v_l = torch.einsum("cij,cmj->cmi", imm, r_ck)
            v_r = self._vel(imm, r_new)[:, None, :]
            rc = sub_sum - 0.5 * (r_ck + r_new[:, None, :])
            turn_lv = ((v_l * rc).sum(-1) <= 0) | ((v_r * rc).sum(-1) <= 0)  # [C, md]
            in_rng = (lvl[None, :] >= i_min[:, None]) & (lvl[None, :] <= i_max[:, None])
            it_turn = (turn_lv & in_rng).any(1)
            s_turn = torch.where(upd, torch.where(first, torch.zeros_like(it_turn), it_turn), s_turn)
            s_n = torch.where(upd, s_n + 1, s_n)

            # ---- subtree finished (full size, U-turn or divergence): double the tree
            target = torch.ones_like(depth) << depth
            done_sub = active & ((s_n >= target) | s_turn | s_div)
            p_bias = torch.clamp(torch.exp(s_w - weight), max=1.0)
            p_bias = torch.where(s_turn | s_div, torch.zeros_like(p_bias), p_bias)
            take2 = done_sub & (self._rand(C, like=st.z) < p_bias)
            zP, UP, gP = sel(take2, s_zP, zP), torch.where(take2, s_UP, UP), sel(take2, s_gP, gP)
            mR, mL = done_sub & s_right, done_sub & ~s_right
            zR, rR, gR = sel(mR, s_z, zR), sel(mR, s_r, rR), sel(mR, s_g, gR)
            zL, rL, gL = sel(mL, s_z, zL), sel(mL, s_r, rL), sel(mL, s_g, gL)
            weight = torch.where(done_sub, torch.logaddexp(weight, s_w), weight)
            r_sum = sel(done_sub, r_sum + s_rsum, r_sum)
            new_turn = s_turn | self._is_turning(imm, rL, rR, r_sum)
            turning = torch.where(done_sub, new_turn, turning)
            diverging = torch.where(done_sub, s_div, diverging)
            sum_acc = torch.where(done_sub, sum_acc + s_acc, sum_acc)
            nprop = torch.where(done_sub, nprop + s_n, nprop)
            depth = torch.where(done_sub, depth + 1, depth)
            s_n = torch.where(done_sub, torch.zeros_like(s_n), s_n)
            finished = done_sub & ((depth >= md) | turning | diverging)
            active = active & ~finished

        accept = sum_acc / nprop.clamp(min=1).to(dt)
        st.z, st.U, st.g = zP, UP, gP
        st.stats = {"accept_prob": accept, "num_steps": nprop, "tree_depth": depth, "diverging": diverging,
                    "potential_energy": UP, "rounds": rounds}
        return st

    # ------------------------------------------------------------------ warmup adaptation
    def _adapt(self, st: NUTSState, t: int, schedule, window_idx: int) -> int:
        num_windows = len(schedule)
        w_end = schedule[window_idx][1]
        if self.adapt_ss:
            gstat = self.target - st.stats["accept_prob"]
            st.da_t = st.da_t + 1
            tt = st.da_t
            st.da_gavg = (1 - 1 / (tt + 10.0)) * st.da_gavg + gstat / (tt + 10.0)
            st.da_x = st.da_prox - torch.sqrt(tt) / 0.05 * st.da_gavg
            wt = tt ** (-0.75)
            st.da_xavg = (1 - wt) * st.da_xavg + wt * st.da_x
            st.step_size = torch.exp(st.da_x.clamp(-700.0, 700.0))
        middle = 0 < window_idx < num_windows - 1
        if self.adapt_mm and middle:
            st.wf_n += 1
            d = st.z - st.wf_mean
            st.wf_mean = st.wf_mean + d / st.wf_n
            d2 = st.z - st.wf_mean
            st.wf_m2 = st.wf_m2 + d[:, :, None] * d2[:, None, :]
        if t == w_end:
            if self.adapt_mm and middle and st.wf_n > 1:
                n = st.wf_n
                cov = st.wf_m2 / (n - 1)
                D = cov.shape[-1]
                eye = torch.eye(D, dtype=cov.dtype, device=cov.device)
                cov = (n / (n + 5.0)) * cov + 1e-3 * (5.0 / (n + 5.0)) * eye
                if not self.dense:
                    cov = torch.diag_embed(torch.diagonal(cov, dim1=1, dim2=2))
                st.inv_mass = cov
                self._wf_reset(st)
            if self.adapt_ss:
                st.step_size = torch.exp(st.da_xavg.clamp(-700.0, 700.0))
                self._da_reset(st)
            window_idx += 1
        return window_idx

    def run(self, z0: torch.Tensor, num_warmup: int, num_samples: int, progress: Optional[Callable] = None):
        """Returns (samples z [C, num_samples, D], per-sample stats dict, final state)."""
        st = self.init(z0)
        schedule = build_adaptation_schedule(num_warmup) if num_warmup > 0 else []
        w = 0
        C, D = z0.shape
        out = torch.empty((C, num_samples, D), dtype=z0.dtype, device=z0.device)
        keep = {k: torch.empty((C, num_samples), dtype=torch.float64, device=z0.device)
                for k in ("accept_prob", "num_steps", "diverging", "potential_energy")}
        for t in range(num_warmup + num_samples):
            st = self.step(st)
            st.i = t + 1
            if t < num_warmup:
                w = self._adapt(st, t, schedule, w)
            else:
                k = t - num_warmup
                out[:, k] = st.z
                for name in keep:
                    keep[name][:, k] = st.stats[name].to(torch.float64)
            if progress is not None:
                progress(t, st)
        return out, keep, st


def effective_sample_size(x: torch.Tensor) -> torch.Tensor:
    """Bulk ESS of draws x [chains, samples] (Geyer initial-positive-sequence, as arviz/numpyro)."""
    x = x.to(torch.float64)
    C, N = x.shape
    xc = x - x.mean(1, keepdim=True)
    nfft = 1 << (2 * N - 1).bit_length()
    f = torch.fft.rfft(xc, n=nfft)
    acov = torch.fft.irfft(f * f.conj(), n=nfft)[:, :N] / N
    chain_var = acov[:, 0] * N / (N - 1.0)
    mean_var = chain_var.mean()
    var_plus = mean_var * (N - 1.0) / N
    if C > 1:
        var_plus = var_plus + x.mean(1).var(unbiased=True)
    rho = 1.0 - (mean_var - acov.mean(0)) / var_plus
    rho[0] = 1.0
    pairs = rho[: (N // 2) * 2].reshape(-1, 2).sum(1)
    pos = torch.cumprod((pairs > 0).to(torch.float64), 0)
    pairs = pairs * pos
    pairs = torch.cummin(pairs, 0).values.clamp(min=0.0)
    tau = -1.0 + 2.0 * pairs.sum()
    return torch.as_tensor(C * N, dtype=torch.float64) / tau.clamp(min=1.0 / math.log10(max(C * N, 10)))


def split_rhat(x: torch.Tensor) -> torch.Tensor:
    """Split-R-hat of draws x [chains, samples]."""
    x = x.to(torch.float64)
    C, N = x.shape
    h = N // 2
    y = torch.cat([x[:, :h], x[:, N - h:]], 0)
    w = y.var(1, unbiased=True).mean()
    b = y.mean(1).var(unbiased=True) * h
    var_plus = (h - 1.0) / h * w + b / h
    return torch.sqrt(var_plus / w)
