"""Seeded synthetic inputs of the parity tests: the workloads of dynode_b200/synthetic.py (shared with bench.py)
plus the cases that need the oracle's helpers."""
import numpy as np

from dynode_b200.synthetic import *  # noqa: F401,F403
from dynode_b200.synthetic import (ALL_CASES, AGE3, CONTACT2, EXTRA_CASES, RISK2, make_case,  # noqa: F401
                                   make_seip_case)


def make_seipv_case(B, A=3, K=2, W=3, V=3, NK=2, seed=20260108, t1=200, season=True, intro=True, vaccinate=True):
    """Immune-history family with the vaccination dimension (oracle FAM_SEIPV): spline vaccination rates per
    (age, tier), a second strain that is absent at t = 0 and arrives by external introduction, seasonal reset of the
    top tier.  Per-draw rates and introduction parameters, shared tables."""
    from dynode_b200.seip import SeipModel, immunity_table_vax
    from oracle import oracle as orc
    rng = np.random.Generator(np.random.PCG64(seed))
    H = 1 << K
    r0 = rng.uniform(1.5, 3.0, (B, K))
    inf = rng.uniform(4, 9, (B, K))
    lat = rng.uniform(2, 4, (B, K))
    wane = rng.uniform(20, 90, (B, W))
    beta, sigma, gamma, omega = r0 / inf, 1 / lat, 1 / inf, 1 / wane
    omega[:, -1] = 0.0
    C = np.random.default_rng(300 + A).uniform(0.1, 1.0, (A, A))
    C = C / np.max(np.real(np.linalg.eigvals(C)))
    pop = 1000.0 * (np.arange(A, 0, -1.0) / np.arange(A, 0, -1.0).sum())
    cross = np.full((K, K), 0.45) + 0.55 * np.eye(K)
    eff = np.linspace(0.0, 0.6, V)[:, None] * np.linspace(1.0, 0.7, K)[None, :]  # [V][K]
    imm = immunity_table_vax(K, np.linspace(0.9, 0.2, W), cross, eff)  # [H][V][W][K]
    S0 = np.zeros((A, H, V, W))
    S0[:, 0, 0, W - 1] = 0.99 * pop
    I0 = np.zeros((A, H, V, K))
    I0[:, 0, 0, 0] = 0.01 * pop  # only strain 0 circulates at t = 0
    y0 = np.concatenate([S0.ravel(), np.zeros(A * H * V * K), I0.ravel(), np.zeros(A * H * V * K)])
    # vaccination: ~0.4 % of an age group per day out of tier 0 ramping up after day 30, less from higher tiers
    vbase = np.zeros((A, V, 4))
    vknots = np.zeros((A, V, NK))
    vcoef = np.zeros((A, V, NK))
    if vaccinate:
        vbase[:, :, 0] = 0.004 * np.linspace(1.0, 0.3, V)[None, :] * np.linspace(1.0, 0.5, A)[:, None]
        vbase[:, :, 1] = -1e-6
        if NK > 0:
            vknots[:] = np.linspace(30.0, 90.0, NK)[None, None, :]
            vcoef[:] = 2e-9 * np.where(np.arange(NK) % 2 == 0, 1.0, -1.0)[None, None, :]
    itime = np.tile(np.linspace(60.0, 90.0, K), (B, 1)) + rng.uniform(-5, 5, (B, K))
    iscale = rng.uniform(3.0, 8.0, (B, K))
    ipct = np.zeros((B, K))
    if intro and K > 1:
        ipct[:, 1:] = rng.uniform(0.005, 0.03, (B, K - 1))
    iages = np.zeros((K, A))
    iages[:, : max(1, A - 1)] = 1.0
    tau = 182.5 - 120.0  # the reset peaks at day 120
    theta = np.hstack([beta, sigma, gamma, omega, itime, iscale, ipct])
    shared = np.concatenate([C.ravel(), pop, imm.ravel(), vbase.ravel(), vknots.ravel(), vcoef.ravel(), iages.ravel(),
                             [tau, 1.0 if season else 0.0]])
    return dict(model=SeipModel(A, K, W, V, NK), params=dict(beta=beta, sigma=sigma, gamma=gamma, omega=omega),
                contact=C, pop=pop, immunity=imm, y0=y0, vaccination=(vbase, vknots, vcoef) if vaccinate else None,
                introductions=dict(time=itime, scale=iscale, pct=ipct, ages=iages) if intro else None,
                season_tau=tau if season else None,
                oracle=(orc.SEIPV, orc.seipv_dims(A, W, K, V, NK), theta, shared), t1=t1)
