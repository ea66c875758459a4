"""Many-chain NUTS on the device.

The reference runs numpyro's `MCMC(NUTS(model, dense_mass=True, max_tree_depth, init_to_median))`
(reference src/dynode/infer/inference.py:149-163); numpyro is absent here and, more to the point, its
chains run one after another on a CPU.  This sampler keeps numpyro's algorithm -- iterative tree doubling
with checkpointed U-turn tests, multinomial proposal sampling with a biased top-level transition,
divergence at an energy error of 1000, dual-averaging step size (target 0.8, t0=10, kappa=0.75,
gamma=0.05), Stan's windowed dense mass-matrix adaptation with Welford covariance and shrinkage -- and
re-organises it for the GPU:

  * all chains advance together; every *round* is ONE batched `potential_and_grad(z[C, D])`, i.e. one
    ensemble launch of the ODE kernel, plus masked tensor updates of the per-chain tree state;
  * chains are NOT held in lock-step per transition: a chain whose tree is complete commits its
    transition (sample, step-size and covariance updates) and starts its next tree in the very next
    round, and -- the adaptation state being per chain, as in numpyro -- walks the whole warm-up + sampling
    schedule at its own pace, doing its own window-end updates (shrunk Welford covariance -> inverse mass
    matrix and its Cholesky factor, dual-averaging restart).  No chain ever waits for another one, so the
    number of rounds is the largest per-chain TOTAL of leapfrogs, not a sum of per-window (or per-transition)
    maxima: measured 2-4x fewer rounds than with window barriers (profiles/r1/nuts_async.md);
  * all state lives in persistent device buffers updated in place, so on CUDA the whole round -- model
    evaluation included -- is captured once into a CUDA graph and replayed (one graph launch per round,
    one host sync every `sync_every` rounds to see whether every chain is finished).
"""

from __future__ import annotations

import math
from types import SimpleNamespace
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


# per-transition schedule flags (DYNODE_NUTS_* of include/dynode_b200_nuts.h)
ADAPT, WELFORD, SAMPLING, END_SLOW, END_WARMUP = 1, 2, 4, 8, 16


def build_transition_schedule(num_warmup: int, num_samples: int, adapt_ss: bool, adapt_mm: bool):
    """(flags uint8 [n], window length float64 [n]) for a chain's n = num_warmup + num_samples transitions:
    what happens after each of them (see the flag definitions)."""
    flags, wlen = [], []
    windows = build_adaptation_schedule(num_warmup) if num_warmup > 0 else []
    for w, (a, e) in enumerate(windows):
        middle = 0 < w < len(windows) - 1
        fl = (ADAPT if adapt_ss else 0) | (WELFORD if (middle and adapt_mm) else 0)
        for t in range(a, e + 1):
            f = fl
            if t == e and middle:
                f |= END_SLOW
            if t == e and w == len(windows) - 1:
                f |= END_WARMUP
            flags.append(f)
            wlen.append(float(e - a + 1))
    flags += [SAMPLING] * num_samples
    wlen += [float(max(1, num_samples))] * num_samples
    return flags, wlen


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


def _put(dst: torch.Tensor, mask: torch.Tensor, src) -> None:
    """dst[mask] = src[mask], in place, static shapes (one fused select kernel)."""
    m = mask.reshape(mask.shape + (1,) * (dst.dim() - mask.dim()))
    if not isinstance(src, torch.Tensor):
        src = torch.full_like(dst, src)
    torch.where(m, src, dst, out=dst)


def _lib_limits():
    from .. import _lib
    return _lib.NUTS_MAX_DIM, _lib.NUTS_MAX_DEPTH


class BatchedNUTS:
    def __init__(self, potential_and_grad: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                 max_tree_depth: int = 10, target_accept_prob: float = 0.8, dense_mass: bool = True,
                 step_size: float = 1.0, adapt_step_size: bool = True, adapt_mass_matrix: bool = True,
                 generator: Optional[torch.Generator] = None, cuda_graph: Optional[bool] = None,
                 sync_every: int = 4, cuda_kernels: Optional[bool] = None,
                 launch_key: Optional[Callable[[int, int], object]] = None):
        self.pg = potential_and_grad
        self.max_depth = int(max_tree_depth)
        self.target = float(target_accept_prob)
        self.dense = dense_mass
        self.init_step = float(step_size)
        self.adapt_ss, self.adapt_mm = adapt_step_size, adapt_mass_matrix
        self.gen = generator
        self.cuda_graph = cuda_graph
        self.cuda_kernels = cuda_kernels  # None: use the hand-written round kernels whenever the chains are on CUDA
        self.sync_every = max(1, int(sync_every))
        self.grad_evals = 0          # leapfrogs that belong to a tree (what "NUTS grad-evals" counts)
        self.launched_evals = 0      # rounds x chains (includes chains idling at a window boundary)
        self.rounds = 0
        self._n_running = None  # chains still running, as told to the model's launches (None: all of them)
        self.launch_key = launch_key  # (chains, chains still running) -> what the model evaluation would launch
        self.recaptures = 0
        self.graph_used = False
        self.b: Optional[SimpleNamespace] = None

    # ------------------------------------------------------------------ helpers
    def _randn(self, *shape):
        b = self.b
        return torch.randn(*shape, dtype=b.dtype, device=b.dev, generator=self._g)

    def _rand(self, *shape):
        b = self.b
        return torch.rand(*shape, dtype=b.dtype, device=b.dev, generator=self._g)

    def _eval(self, z):
        U, g = self.pg(z)
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

    # ------------------------------------------------------------------ buffers
    def set_schedule(self, flags, wlen, num_warmup: int):
        """Install a per-transition schedule (flags / window lengths, see build_transition_schedule); all
        chains restart at transition 0."""
        b = self.b
        n = max(1, len(flags))
        b.sched = torch.tensor(list(flags) or [0], dtype=torch.uint8, device=b.dev)
        b.sched_n = torch.tensor(list(wlen) or [1.0], dtype=b.dtype, device=b.dev)
        b.n_warm = int(num_warmup)
        b.nwin.fill_(len(flags))
        b.k.zero_()
        b.active.fill_(len(flags) > 0)
        b.any_active.fill_(len(flags) > 0)

    def _allocate(self, z0: torch.Tensor, num_samples: int):
        C, D = z0.shape
        dev, dt, md = z0.device, z0.dtype, self.max_depth
        f = lambda *s: torch.zeros(*s, dtype=dt, device=dev)
        l = lambda *s: torch.zeros(*s, dtype=torch.long, device=dev)
        bl = lambda *s: torch.zeros(*s, dtype=torch.bool, device=dev)
        b = SimpleNamespace(C=C, D=D, dev=dev, dtype=dt)
        b.z, b.U, b.g = z0.clone(), f(C), f(C, D)
        b.eps = torch.full((C,), self.init_step, dtype=dt, device=dev)
        b.imm = torch.eye(D, dtype=dt, device=dev).expand(C, D, D).contiguous()
        b.msqrt = b.imm.clone()
        b.k, b.nwin = l(C), l(())
        b.active, b.need_tree = bl(C), bl(C)
        b.searching, b.fr_dir, b.fr_last = bl(C), l(C), l(C)  # find_reasonable_step_size probes
        b.energy0 = f(C)
        for name in ("zL", "rL", "gL", "zR", "rR", "gR", "zP", "gP", "r_sum", "s_z", "s_r", "s_g", "s_zP", "s_gP", "s_rsum"):
            setattr(b, name, f(C, D))
        for name in ("UP", "weight", "sum_acc", "s_UP", "s_w", "s_acc"):
            setattr(b, name, f(C))
        for name in ("depth", "nprop", "s_n"):
            setattr(b, name, l(C))
        for name in ("turning", "diverging", "s_right", "s_turn", "s_div"):
            setattr(b, name, bl(C))
        b.r_ck, b.rs_ck = f(C, md, D), f(C, md, D)
        b.da_x, b.da_xavg, b.da_gavg, b.da_t = f(C), f(C), f(C), f(C)
        b.da_prox = torch.log(10.0 * b.eps)
        b.wf_n, b.wf_mean, b.wf_m2 = f(C), f(C, D), f(C, D, D)
        b.sched = torch.zeros(1, dtype=torch.uint8, device=dev)
        b.sched_n = f(1) + 1.0
        b.n_warm = 0
        N = max(1, num_samples)
        b.out_z = f(C, N, D)
        b.out_stats = {k: f(C, N) for k in ("accept_prob", "num_steps", "diverging", "potential_energy", "tree_depth")}
        b.last_accept, b.last_steps = f(C), f(C)
        b.n_useful = l(())
        b.n_leap = l(C)
        b.z_new, b.r_half = f(C, D), f(C, D)
        b.any_active = bl(())
        b.idx_min_tab, b.idx_max_tab = _tree_index_tables(md, dev)
        b.ar = torch.arange(C, device=dev)
        b.lvl = torch.arange(md, device=dev)
        self.b = b

    # ------------------------------------------------------------------ one round for all chains
    def _round(self):
        b = self.b
        C, D, md = b.C, b.D, self.max_depth
        imm, eps = b.imm, b.eps
        act = b.active.clone()  # chains that take part in this round
        rnd_n, rnd_u = self._randn(C, D), self._rand(C, 3)  # same draws, same roles as the CUDA round

        # ---- step-size probes (numpyro find_reasonable_step_size): fresh momentum, one leapfrog forward from the
        # current state with eps * 2^dir
        probe = act & b.searching
        r0 = torch.einsum("cij,cj->ci", b.msqrt, rnd_n)
        fac = torch.where(b.fr_dir > 0, 2.0, 1.0) * torch.where(b.fr_dir < 0, 0.5, 1.0)
        _put(b.eps, probe, b.eps * fac.to(b.dtype))
        _put(b.energy0, probe, b.U + self._kinetic(imm, r0))
        _put(b.s_z, probe, b.z)
        _put(b.s_r, probe, r0)
        _put(b.s_g, probe, b.g)
        _put(b.s_right, probe, True)
        act = act & ~probe  # what follows is tree building

        # ---- chains beginning a transition: fresh momentum r ~ N(0, M), one-node tree at the current state
        nt = act & b.need_tree
        _put(b.energy0, nt, b.U + self._kinetic(imm, r0))
        for dst, src in ((b.zL, b.z), (b.zR, b.z), (b.zP, b.z), (b.gL, b.g), (b.gR, b.g), (b.gP, b.g),
                         (b.rL, r0), (b.rR, r0), (b.r_sum, r0)):
            _put(dst, nt, src)
        _put(b.UP, nt, b.U)
        for dst in (b.weight, b.sum_acc):
            _put(dst, nt, 0.0)
        for dst in (b.depth, b.nprop, b.s_n):
            _put(dst, nt, 0)
        for dst in (b.turning, b.diverging):
            _put(dst, nt, False)
        _put(b.need_tree, nt, False)

        # ---- chains starting a new doubling: pick a direction, start from that edge of the tree
        start = act & (b.s_n == 0)
        _put(b.s_right, start, rnd_u[:, 0] < 0.5)
        sr = b.s_right[:, None]
        _put(b.s_z, start, torch.where(sr, b.zR, b.zL))
        _put(b.s_r, start, torch.where(sr, b.rR, b.rL))
        _put(b.s_g, start, torch.where(sr, b.gR, b.gL))

        # ---- one leapfrog for every chain (idle chains are evaluated too and discarded)
        h = torch.where(b.s_right, eps, -eps)[:, None]
        r_half = b.s_r - 0.5 * h * b.s_g
        z_new = b.s_z + h * self._vel(imm, r_half)
        U_new, g_new = self._eval(z_new)
        r_new = r_half - 0.5 * h * g_new
        delta = U_new + self._kinetic(imm, r_new) - b.energy0
        # ---- outcome of the probes (_body_fn / _cond_fn of find_reasonable_step_size)
        dir_new = torch.where(math.log(0.8) < -delta, 1, -1)  # NaN compares false: -1
        tiny, big = 2.2250738585072014e-308, 1.7976931348623157e308
        not_extreme = ((eps > tiny) | (dir_new >= 0)) & ((eps < big) | (dir_new <= 0))
        go_on = probe & not_extreme & ((b.fr_dir == 0) | (dir_new == b.fr_dir))
        stop = probe & ~go_on
        _put(b.fr_last, go_on, b.fr_dir)
        _put(b.fr_dir, go_on, dir_new)
        _put(b.searching, stop, False)
        _put(b.fr_dir, stop, 0)
        _put(b.fr_last, stop, 0)
        _put(b.da_prox, stop, torch.log(10.0 * eps))
        for t_ in (b.da_x, b.da_xavg, b.da_gavg, b.da_t):
            _put(t_, stop, 0.0)
        _put(b.need_tree, stop, True)
        delta = torch.where(torch.isnan(delta), torch.full_like(delta, math.inf), delta)
        leaf_w = -delta
        leaf_div = delta > MAX_DELTA_ENERGY
        leaf_acc = torch.clamp(torch.exp(-delta), max=1.0)

        # ---- fold the leaf into the subtree (uniform/multinomial transition inside a subtree)
        first = b.s_n == 0
        new_w = torch.where(first, leaf_w, torch.logaddexp(b.s_w, leaf_w))
        p_take = torch.where(first, torch.ones_like(leaf_w), torch.sigmoid(leaf_w - b.s_w))
        take = act & (rnd_u[:, 1] < p_take)
        _put(b.s_zP, take, z_new)
        _put(b.s_UP, take, U_new)
        _put(b.s_gP, take, g_new)
        _put(b.s_rsum, act, torch.where(first[:, None], r_new, b.s_rsum + r_new))
        _put(b.s_z, act, z_new)
        _put(b.s_r, act, r_new)
        _put(b.s_g, act, g_new)
        _put(b.s_w, act, new_w)
        _put(b.s_acc, act, torch.where(first, leaf_acc, b.s_acc + leaf_acc))
        _put(b.s_div, act, leaf_div)

        # ---- checkpointed U-turn tests over the sub-subtrees that this leaf completes
        n = b.s_n
        i_min, i_max = b.idx_min_tab[n], b.idx_max_tab[n]
        even = (n % 2 == 0) & act
        slot = (b.lvl[None, :] == i_max[:, None]) & even[:, None]
        _put(b.r_ck, slot, r_new[:, None, :].expand(C, md, D))
        _put(b.rs_ck, slot, b.s_rsum[:, None, :].expand(C, md, D))
        sub_sum = b.s_rsum[:, None, :] - b.rs_ck + b.r_ck                # [C, md, D]
        v_l = torch.einsum("cij,cmj->cmi", imm, b.r_ck)
        v_r = self._vel(imm, r_new)[:, None, :]
        rc = sub_sum - 0.5 * (b.r_ck + r_new[:, None, :])
        turn_lv = ((v_l * rc).sum(-1) <= 0) | ((v_r * rc).sum(-1) <= 0)  # [C, md]
        in_rng = (b.lvl[None, :] >= i_min[:, None]) & (b.lvl[None, :] <= i_max[:, None])
        it_turn = (turn_lv & in_rng).any(1) & ~first
        _put(b.s_turn, act, it_turn)
        _put(b.s_n, act, b.s_n + 1)

        # ---- subtree finished (full size, U-turn or divergence): double the tree
        target = torch.ones_like(b.depth) << b.depth
        done_sub = act & ((b.s_n >= target) | b.s_turn | b.s_div)
        p_bias = torch.clamp(torch.exp(b.s_w - b.weight), max=1.0)
        p_bias = torch.where(b.s_turn | b.s_div, torch.zeros_like(p_bias), p_bias)
        take2 = done_sub & (rnd_u[:, 2] < p_bias)
        _put(b.zP, take2, b.s_zP)
        _put(b.UP, take2, b.s_UP)
        _put(b.gP, take2, b.s_gP)
        mR, mL = done_sub & b.s_right, done_sub & ~b.s_right
        for dst, src in ((b.zR, b.s_z), (b.rR, b.s_r), (b.gR, b.s_g)):
            _put(dst, mR, src)
        for dst, src in ((b.zL, b.s_z), (b.rL, b.s_r), (b.gL, b.s_g)):
            _put(dst, mL, src)
        _put(b.weight, done_sub, torch.logaddexp(b.weight, b.s_w))
        _put(b.r_sum, done_sub, b.r_sum + b.s_rsum)
        _put(b.turning, done_sub, b.s_turn | self._is_turning(imm, b.rL, b.rR, b.r_sum))
        _put(b.diverging, done_sub, b.s_div)
        _put(b.sum_acc, done_sub, b.sum_acc + b.s_acc)
        _put(b.nprop, done_sub, b.nprop + b.s_n)
        _put(b.depth, done_sub, b.depth + 1)
        _put(b.s_n, done_sub, 0)

        # ---- tree complete: commit the transition for those chains and let them start the next one
        fin = done_sub & ((b.depth >= md) | b.turning | b.diverging)
        accept = b.sum_acc / b.nprop.clamp(min=1).to(b.dtype)
        _put(b.z, fin, b.zP)
        _put(b.U, fin, b.UP)
        _put(b.g, fin, b.gP)
        _put(b.last_accept, fin, accept)
        _put(b.last_steps, fin, b.nprop.to(b.dtype))
        kidx = b.k.clamp(max=b.sched.numel() - 1)
        fl = b.sched[kidx].to(torch.int32)
        # dual averaging of log step size (warmup windows)
        ad = fin & ((fl & ADAPT) != 0)
        tt = b.da_t + 1.0
        gavg = (1.0 - 1.0 / (tt + 10.0)) * b.da_gavg + (self.target - accept) / (tt + 10.0)
        x = b.da_prox - torch.sqrt(tt) / 0.05 * gavg
        wt = tt ** (-0.75)
        xavg = (1.0 - wt) * b.da_xavg + wt * x
        _put(b.da_t, ad, tt)
        _put(b.da_gavg, ad, gavg)
        _put(b.da_x, ad, x)
        _put(b.da_xavg, ad, xavg)
        _put(b.eps, ad, torch.exp(x.clamp(-700.0, 700.0)))
        # Welford covariance of the positions (slow windows)
        wf = fin & ((fl & WELFORD) != 0)
        n1 = b.wf_n + 1.0
        d1 = b.z - b.wf_mean
        mean1 = b.wf_mean + d1 / n1[:, None]
        d2 = b.z - mean1
        _put(b.wf_m2, wf, b.wf_m2 + d1[:, :, None] * d2[:, None, :])
        _put(b.wf_mean, wf, mean1)
        _put(b.wf_n, wf, n1)
        # end of a slow window: the chain's own mass-matrix update and dual-averaging restart
        es = fin & ((fl & END_SLOW) != 0)
        # (on CUDA this masked-tensor round -- the cross-check, not the product -- stays free of host syncs so
        # that it can still be graph-captured: the update is then computed every round and masked)
        if b.dev.type == "cuda" or bool(es.any()):
            nn = b.sched_n[kidx]
            mm = es & ((fl & WELFORD) != 0)
            upd = mm & (nn > 1.0)
            eye = torch.eye(D, dtype=b.dtype, device=b.dev)
            cov = b.wf_m2 / (nn - 1.0).clamp(min=1.0)[:, None, None]
            cov = (nn / (nn + 5.0))[:, None, None] * cov + (1e-3 * (5.0 / (nn + 5.0)))[:, None, None] * eye
            if not self.dense:
                cov = torch.diag_embed(torch.diagonal(cov, dim1=1, dim2=2))
            _put(b.imm, upd, cov)
            Lc = torch.linalg.cholesky_ex(b.imm, check_errors=False).L
            msq = torch.linalg.solve_triangular(Lc.transpose(1, 2), eye.expand(C, D, D), upper=True)  # L^-T
            _put(b.msqrt, upd, msq)
            _put(b.wf_n, mm, 0.0)
            _put(b.wf_mean, mm, 0.0)
            _put(b.wf_m2, mm, 0.0)
            ads = es & ((fl & ADAPT) != 0)  # step-size search under the new metric, from the current step size
            _put(b.searching, ads, True)
            _put(b.fr_dir, ads, 0)
            _put(b.fr_last, ads, 0)
        ew = fin & ((fl & END_WARMUP) != 0) & ((fl & ADAPT) != 0)
        _put(b.eps, ew, torch.exp(b.da_xavg.clamp(-700.0, 700.0)))
        # store the draw (sampling phase)
        st = fin & ((fl & SAMPLING) != 0)
        kk = (b.k - b.n_warm).clamp(min=0, max=b.out_z.shape[1] - 1)
        cur = b.out_z[b.ar, kk]
        b.out_z[b.ar, kk] = torch.where(st[:, None], b.z, cur)
        for name, val in (("accept_prob", accept), ("num_steps", b.nprop.to(b.dtype)),
                          ("diverging", b.diverging.to(b.dtype)), ("potential_energy", b.U),
                          ("tree_depth", b.depth.to(b.dtype))):
            buf = b.out_stats[name]
            buf[b.ar, kk] = torch.where(st, val, buf[b.ar, kk])
        _put(b.k, fin, b.k + 1)
        _put(b.need_tree, fin, True)
        _put(b.active, fin & (b.k >= b.nwin), False)
        b.n_useful += act.sum() + probe.sum()
        b.any_active.copy_(b.active.any())

    # ------------------------------------------------------------------ the same round on the CUDA kernels
    def _bind_cuda_state(self):
        """DynodeNutsState over the persistent buffers (include/dynode_b200_nuts.h)."""
        from .. import _lib
        b = self.b
        st = _lib.NutsState()
        st.C, st.D, st.max_depth, st.N = b.C, b.D, self.max_depth, b.out_z.shape[1]
        st.n_warmup, st.dense = int(b.n_warm), int(bool(self.dense))
        st.target_accept = self.target
        alias = {"out_accept": b.out_stats["accept_prob"], "out_steps": b.out_stats["num_steps"],
                 "out_div": b.out_stats["diverging"], "out_energy": b.out_stats["potential_energy"],
                 "out_depth": b.out_stats["tree_depth"]}
        for name in _lib._NUTS_PTRS:
            t = alias.get(name, getattr(b, name, None))
            assert t is not None and t.is_cuda and t.is_contiguous(), name
            setattr(st, name, t.data_ptr())
        self._st = st
        self._lib = _lib

    def _round_cuda(self):
        import ctypes
        b, L = self.b, self._lib.load()
        rnd_n, rnd_u = self._randn(b.C, b.D), self._rand(b.C, 3)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        self._lib.check(L.dynode_nuts_round_pre(ctypes.byref(self._st), rnd_n.data_ptr(), rnd_u.data_ptr(), stream))
        # finished chains cost nothing in the model's ensemble launches (engine.only_rows -> DynodeSolverDesc.only)
        from .. import engine as _engine
        with _engine.only_rows(b.active.view(torch.uint8), n_rows=self._n_running):
            U_new, g_new = self.pg(b.z_new)
        U_new, g_new = U_new.contiguous(), g_new.contiguous()
        self._lib.check(L.dynode_nuts_round_post(ctypes.byref(self._st), U_new.data_ptr(), g_new.data_ptr(),
                                                 rnd_u.data_ptr(), stream))  # also maintains b.any_active

    # ------------------------------------------------------------------ graph capture
    def _prepare_round_fn(self):
        b = self.b
        want = self.cuda_graph if self.cuda_graph is not None else b.dev.type == "cuda"
        self._g = self.gen
        # the one-thread-per-chain kernels hold a chain's vectors in registers: up to 16 dimensions and 12 doublings
        # (include/dynode_b200_nuts.h); larger models run the same round as masked tensor operations on the device
        fits = b.D <= _lib_limits()[0] and self.max_depth <= _lib_limits()[1]
        use_kernels = self.cuda_kernels if self.cuda_kernels is not None else (b.dev.type == "cuda" and fits)
        if use_kernels and b.dev.type != "cuda":
            raise RuntimeError("the NUTS round kernels need the chains on a CUDA device")
        self.kernels_used = bool(use_kernels)
        if use_kernels:
            self._bind_cuda_state()
        round_impl = self._round_cuda if use_kernels else self._round
        self._round_fn = round_impl
        self._launch_key_now = self.launch_key(b.C, b.C) if self.launch_key is not None else None
        self.graph_used = False
        if not want or b.dev.type != "cuda":
            return
        try:
            self._g = None  # the default CUDA generator is graph-safe (philox offsets are patched per replay)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):  # warm-up rounds are real rounds of the algorithm
                    round_impl()
                    self.rounds += 1
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._round_impl = round_impl
            self._capture()
        except Exception as e:  # the SAME round implementation runs eagerly; say that replay is off and why
            import traceback
            import warnings
            self.capture_error = traceback.format_exc()
            torch.cuda.synchronize()
            self._g = self.gen
            self._round_fn = round_impl  # kernels_used stays what it says: only the replay is dropped
            self.graph_used = False
            where = "".join(self.capture_error.splitlines(keepends=True)[-14:])
            warnings.warn(f"BatchedNUTS: CUDA-graph capture of the round failed ({type(e).__name__}: {e}); "
                          f"running rounds eagerly\n{where}")

    def _capture(self):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._round_impl()
        self._graph = graph
        self._round_fn = graph.replay
        self.graph_used = True

    # ------------------------------------------------------------------ driver
    LOOK_EVERY = 64  # rounds between two looks at the number of chains still running

    def _look_at_running_chains(self):
        """Most rounds of a run belong to the few chains that are still building deep trees (config 5: the slowest of
        1024 chains takes 4-5x the mean number of leapfrogs).  How the model's gradient is best launched depends on
        how many rows really run (simulation.autograd.use_adjoint: below a GPU's worth of warps, independent forward
        directions beat the adjoint's longer dependent chain), and that choice is frozen into the captured graph -- so
        when the model says its launches would change (`launch_key`), the round is captured again."""
        n = int(self.b.active.sum())
        if n <= 0:
            return
        key = self.launch_key(self.b.C, n)
        if key == self._launch_key_now:
            return
        self._launch_key_now, self._n_running = key, n
        if self.graph_used:
            try:
                torch.cuda.synchronize()
                self._round_impl()  # one eager round first: a launch path that is new must not load inside a capture
                self.rounds += 1
                torch.cuda.synchronize()
                self._capture()
                self.recaptures += 1
            except Exception:  # keep replaying the graph that exists
                torch.cuda.synchronize()

    def _run_schedule(self, progress: Optional[Callable] = None):
        """Rounds until every chain has walked the whole schedule (one host sync every `sync_every` rounds)."""
        b = self.b
        step = max(1, int(b.nwin) // 10)
        next_mark = step
        next_look = self.LOOK_EVERY
        while True:
            for _ in range(self.sync_every):
                self._round_fn()
                self.rounds += 1
            if not bool(b.any_active):
                break
            if self.launch_key is not None and self.kernels_used and self.rounds >= next_look:
                self._look_at_running_chains()
                next_look = self.rounds + self.LOOK_EVERY
            if progress is not None:
                kmin = int(b.k.min())  # the slowest chain's position in the schedule
                if kmin >= next_mark:
                    progress(kmin - 1, self)
                    next_mark = (kmin // step + 1) * step

    def run(self, z0: torch.Tensor, num_warmup: int, num_samples: int, progress: Optional[Callable] = None):
        """Returns (samples z [C, num_samples, D], per-sample stats dict, final state namespace)."""
        import time
        t0 = time.perf_counter()
        self._allocate(z0, num_samples)
        b = self.b
        self._g = self.gen
        U, g = self._eval(b.z)
        if not bool(torch.isfinite(U).all()):
            raise RuntimeError("cannot find valid initial parameters: the potential is not finite at the "
                               "initial position of some chain")
        b.U.copy_(U)
        b.g.copy_(g)
        b.need_tree.fill_(True)
        flags, wlen = build_transition_schedule(num_warmup, num_samples, self.adapt_ss, self.adapt_mm)
        self.set_schedule(flags, wlen, num_warmup)
        if self.adapt_ss:
            b.searching.fill_(True)  # warmup_adapter.init: find_reasonable_step_size before the first transition
        t1 = time.perf_counter()  # the isfinite check above synchronised
        self._prepare_round_fn()
        t2 = time.perf_counter()
        if len(flags) > 0:
            self._run_schedule(progress)
            if progress is not None:
                progress(num_warmup + num_samples - 1, self)
        self.grad_evals = int(b.n_useful) + int(b.n_leap.sum())
        self.timing = {"first_eval_s": t1 - t0, "capture_s": t2 - t1, "rounds_s": time.perf_counter() - t2}
        self.launched_evals = self.rounds * b.C
        stats = {k: v.clone() for k, v in b.out_stats.items()}
        return b.out_z.clone(), stats, b

    # convenience for progress lines
    @property
    def step_size(self):
        return self.b.eps

    @property
    def stats(self):
        return {"accept_prob": self.b.last_accept, "num_steps": self.b.last_steps}


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
