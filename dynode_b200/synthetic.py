"""Seeded synthetic workloads of BASELINE.json's configurations: what `bench.py` times and what the
parity tests feed to both sides.

Every case gives the same inputs in two forms: the engine's (FlowModel, parameter dict, contact[target][source]) and
a flat one (family id, dims, theta rows, shared table) that a checker can consume -- the CPU oracle's calling
convention; nothing of the oracle is imported here.
"""
import numpy as np

from dynode_b200 import _lib
from dynode_b200.engine import FlowModel

# family ids of the flat form (the numbering of oracle/dynode_oracle.cpp)
O_SIR_1BIN, O_SIR_DENSITY, O_SEIRS_1BIN, O_SEIRS_SEASONAL, O_SIR_AGE, O_SIR_AGE_RISK, O_SEIRS_MULTI = range(7)

CONTACT2 = np.array([[0.7, 0.3], [0.3, 0.7]])
AGE3 = np.array([[0.8, 0.2, 0.0], [0.2, 0.8, 0.0], [0.0, 0.0, 1.0]])
RISK2 = np.array([[0.5, 0.5], [0.5, 0.5]])


def _rates(rng, B, S, r0=(1.5, 3.5), inf=(5, 9), lat=(2, 4), wan=(40, 100)):
    r0 = rng.uniform(*r0, size=(B, S))
    inf = rng.uniform(*inf, size=(B, S))
    lat = rng.uniform(*lat, size=(B, S))
    wan = rng.uniform(*wan, size=(B, S))
    return r0 / inf, 1 / inf, 1 / lat, 1 / wan, r0


def make_case(name, B, seed=20260101):
    """Returns dict(model, params, contact, y0, oracle=(family, dims, theta, shared), t1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if name in ("sir_1bin", "sir_density"):
        beta, gamma, _, _, _ = _rates(rng, B, 1, r0=(1.2, 4.0), inf=(3, 10))
        dens = name == "sir_density"
        y0 = np.array([99.0, 1.0, 0.0]) if dens else np.array([0.9, 0.1, 0.0])
        if dens:
            beta = beta / 100.0
        return dict(model=FlowModel(_lib.FLOW_SIR, _lib.FLAG_DENSITY_DEP if dens else 0, 1, 1),
                    params=dict(beta=beta, gamma=gamma), contact=None, y0=y0,
                    oracle=(O_SIR_DENSITY if dens else O_SIR_1BIN, (1, 1, 1), np.hstack([beta, gamma]), None), t1=150)
    if name in ("seirs_1bin", "seirs_seasonal"):
        beta, gamma, sigma, omega, _ = _rates(rng, B, 1, r0=(1.2, 4.0), inf=(3, 10), lat=(1.5, 5), wan=(40, 200))
        y0 = np.array([0.99, 0.0, 0.01, 0.0])
        if name == "seirs_1bin":
            return dict(model=FlowModel(_lib.FLOW_SEIRS, 0, 1, 1),
                        params=dict(beta=beta, gamma=gamma, sigma=sigma, omega=omega), contact=None, y0=y0,
                        oracle=(O_SEIRS_1BIN, (1, 1, 1), np.hstack([beta, gamma, sigma, omega]), None), t1=365)
        amp = rng.uniform(0, 0.4, size=(B, 1))
        phase = rng.uniform(0, 2 * np.pi, size=(B, 1))
        period = np.full((B, 1), 365.0)
        return dict(model=FlowModel(_lib.FLOW_SEIRS, _lib.FLAG_SEASONAL, 1, 1),
                    params=dict(beta=beta, gamma=gamma, sigma=sigma, omega=omega, season_amp=amp,
                                season_phase=phase, season_period=period), contact=None, y0=y0,
                    oracle=(O_SEIRS_SEASONAL, (1, 1, 1), np.hstack([beta, gamma, sigma, omega, amp, phase, period]), None),
                    t1=365)
    if name in ("sir_age2", "sir_age3", "sir_age4"):
        A = int(name[-1])
        beta, gamma, _, _, _ = _rates(rng, B, 1, r0=(1.5, 3.0), inf=(4, 10))
        if A == 2:
            C = CONTACT2 / np.max(np.real(np.linalg.eigvals(CONTACT2)))
            demo = np.array([0.75, 0.25])
        else:
            C = np.random.default_rng(5 if A == 4 else 50 + A).uniform(0.1, 1.0, (A, A))
            C = C / np.max(np.real(np.linalg.eigvals(C)))
            demo = np.array([0.4, 0.3, 0.2, 0.1]) if A == 4 else np.array([0.5, 0.3, 0.2])
        y0 = np.concatenate([1000 * 0.99 * demo, 1000 * 0.01 * demo, np.zeros(A)])
        return dict(model=FlowModel(_lib.FLOW_SIR, 0, A, 1), params=dict(beta=beta, gamma=gamma), contact=C,
                    y0=y0, oracle=(O_SIR_AGE, (A, 1, 1), np.hstack([beta, gamma]), C), t1=100)
    if name == "sir_age_risk32":
        beta, gamma, _, _, _ = _rates(rng, B, 1, r0=(1.5, 3.0), inf=(4, 10))
        CM = np.einsum("ij,kl->ikjl", AGE3, RISK2)  # [i,j,k,l] source (i,j) -> target (k,l)
        demo = np.array([0.7, 0.2, 0.1])[:, None] * np.array([[0.5, 0.5]] * 3)
        s0 = np.array([[0.99, 1.0], [0.99, 0.99], [1.0, 1.0]])
        i0 = np.array([[0.01, 0.0], [0.01, 0.01], [0.0, 0.0]])
        y0 = np.concatenate([(1000 * demo * s0).ravel(), (1000 * demo * i0).ravel(), np.zeros(6)])
        K = CM.reshape(6, 6).T  # engine layout: contact[target][source]
        return dict(model=FlowModel(_lib.FLOW_SIR, 0, 6, 1), params=dict(beta=beta, gamma=gamma), contact=K,
                    y0=y0, oracle=(O_SIR_AGE_RISK, (3, 2, 1), np.hstack([beta, gamma]), CM), t1=150)
    if name.startswith("seirs_multi_"):
        import re
        m = re.fullmatch(r"seirs_multi_[ag](\d+)s(\d+)", name)
        G, S = int(m.group(1)), int(m.group(2))
        beta, gamma, sigma, omega, r0 = _rates(rng, B, S)
        if G == 2:
            C = CONTACT2
            demo = np.array([0.75, 0.25])
        elif G == 6:
            C = np.kron(AGE3, RISK2)
            demo = (np.array([0.7, 0.2, 0.1])[:, None] * np.array([[0.5, 0.5]] * 3)).ravel()
        else:
            C = np.random.default_rng(100 + G).uniform(0.1, 1.0, (G, G))
            C = C / np.max(np.real(np.linalg.eigvals(C)))
            demo = np.arange(G, 0, -1.0)
            demo = demo / demo.sum()
        dom = r0 / r0.sum(1, keepdims=True)  # initial infections split by r0 (multi-strain example :152-167)
        s_0 = np.broadcast_to(1000 * 0.99 * demo, (B, G))
        i_0 = 1000 * 0.01 * demo[None, :, None] * dom[:, None, :]
        z = np.zeros((B, G * S))
        y0 = np.concatenate([s_0, z, i_0.reshape(B, G * S), z, z], axis=1)
        return dict(model=FlowModel(_lib.FLOW_SEIRS_C, 0, G, S),
                    params=dict(beta=beta, gamma=gamma, sigma=sigma, omega=omega), contact=C, y0=y0,
                    oracle=(O_SEIRS_MULTI, (G, 1, S), np.hstack([beta, gamma, sigma, omega]), C), t1=365)
    raise KeyError(name)


ALL_CASES = ("sir_1bin", "sir_density", "seirs_1bin", "seirs_seasonal", "sir_age2", "sir_age4",
             "sir_age_risk32", "seirs_multi_a2s3", "seirs_multi_g6s3")
# further compiled members of the families (csrc/instances.def 9-17)
EXTRA_CASES = ("sir_age3", "seirs_multi_g1s1", "seirs_multi_g1s2", "seirs_multi_g1s3", "seirs_multi_g2s2",
               "seirs_multi_g3s2", "seirs_multi_g4s2", "seirs_multi_g3s3", "seirs_multi_g4s3")


def make_seip_case(B, A=3, K=2, W=3, seed=20260107, t1=200):
    """Immune-history / waning family (oracle FAM_SEIP, include/dynode_b200_seip.h): per-draw rates, shared
    contact / population / immunity tables, everyone susceptible-and-fully-waned except 1 % infectious."""
    from dynode_b200.seip import SeipModel, immunity_table
    rng = np.random.Generator(np.random.PCG64(seed))
    H = 1 << K
    r0 = rng.uniform(1.5, 3.0, (B, K))
    inf = rng.uniform(4, 9, (B, K))
    lat = rng.uniform(2, 4, (B, K))
    wane = rng.uniform(20, 90, (B, W))
    beta, sigma, gamma, omega = r0 / inf, 1 / lat, 1 / inf, 1 / wane
    omega[:, -1] = 0.0
    C = np.random.default_rng(200 + A).uniform(0.1, 1.0, (A, A))
    C = C / np.max(np.real(np.linalg.eigvals(C)))
    pop = 1000.0 * (np.arange(A, 0, -1.0) / np.arange(A, 0, -1.0).sum())
    cross = np.full((K, K), 0.45) + 0.55 * np.eye(K)
    imm = immunity_table(K, np.linspace(0.9, 0.2, W), cross)
    S0 = np.zeros((A, H, W))
    S0[:, 0, W - 1] = 0.99 * pop
    I0 = np.zeros((A, H, K))
    I0[:, 0, :] = 0.01 * pop[:, None] / K
    y0 = np.concatenate([S0.ravel(), np.zeros(A * H * K), I0.ravel(), np.zeros(A * H * K)])
    theta = np.hstack([beta, sigma, gamma, omega])
    shared = np.concatenate([C.ravel(), pop, imm.ravel()])
    return dict(model=SeipModel(A, K, W), params=dict(beta=beta, sigma=sigma, gamma=gamma, omega=omega), contact=C,
                pop=pop, immunity=imm, y0=y0, oracle=(7, (A, W, K), theta, shared), t1=t1)
