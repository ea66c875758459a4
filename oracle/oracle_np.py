"""Readable numpy twin of the CPU oracle: diffeqsolve(Tsit5, PIDController, SaveAt(ts)) for one
trajectory with an arbitrary Python right-hand side on a tuple-of-arrays state.
TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see oracle/dynode_oracle.cpp header).

It is written independently of the C++ oracle (k = h*f stage form, vectorised over the state)
so that the two restatements check each other, and because it accepts any Python callable it
can integrate the reference's own example RHS functions (tests/golden/make_rhs_golden.py).

Follows: reference src/dynode/simulation/odes.py:107-144 (call into diffrax), SURVEY.md 8a rows
a3 (loop, clip-to-end, save rule), a4 (Tsit5), a5 (dense output), a6 (I-controller),
a7 (initial step), a9 (constant step).
"""

from __future__ import annotations

import numpy as np

C = np.array([161 / 1000, 327 / 1000, 9 / 10,
              0.9800255409045096857298102862870245954942137979563024768854764293221195950761080302604,
              1.0, 1.0])
A = [
    np.array([161 / 1000]),
    np.array([-0.8480655492356988544426874250230774675121177393430391537369234245294192976164141156943e-2,
              0.3354806554923569885444268742502307746751211773934303915373692342452941929761641411569]),
    np.array([2.897153057105493432130432594192938764924887287701866490314866693455023795137503079289,
              -6.359448489975074843148159912383825625952700647415626703305928850207288721235210244366,
              4.362295432869581411017727318190886861027813359713760212991062156752264926097707165077]),
    np.array([5.325864828439256604428877920840511317836476253097040101202360397727981648835607691791,
              -11.74888356406282787774717033978577296188744178259862899288666928009020615663593781589,
              7.495539342889836208304604784564358155658679161518186721010132816213648793440552049753,
              -0.9249506636175524925650207933207191611349983406029535244034750452930469056411389539635e-1]),
    np.array([5.861455442946420028659251486982647890394337666164814434818157239052507339770711679748,
              -12.92096931784710929170611868178335939541780751955743459166312250439928519268343184452,
              8.159367898576158643180400794539253485181918321135053305748355423955009222648673734986,
              -0.7158497328140099722453054252582973869127213147363544882721139659546372402303777878835e-1,
              -0.2826905039406838290900305721271224146717633626879770007617876201276764571291579142206e-1]),
    np.array([0.9646076681806522951816731316512876333711995238157997181903319145764851595234062815396e-1,
              1 / 100,
              0.4798896504144995747752495322905965199130404621990332488332634944254542060153074523509,
              1.379008574103741893192274821856872770756462643091360525934940067397245698027561293331,
              -3.290069515436080679901047585711363850115683290894936158531296799594813811049925401677,
              2.324710524099773982415355918398765796109060233222962411944060046314465391054716027841]),
]
B_SOL = np.concatenate([A[5], [0.0]])
B_HAT = np.array([
    0.9468075576583945807478876255758922856117527357724631226139574065785592789071067303271e-1,
    0.9183565540343253096776363936645313759813746240984095238905939532922955247253608687270e-2,
    0.4877705284247615707855642599631228241516691959761363774365216240304071651579571959813,
    1.234297566930478985655109673884237654035539930748192848315425833500484878378061439761,
    -2.707712349983525454881109975059321670689605166938197378763992255714444407154902012702,
    1.866628418170587035753719399566211498666255505244122593996591602841258328965767580089,
    1 / 66])
B_ERR = B_SOL - B_HAT


def dense_weights(th):
    return np.array([
        -1.0530884977290216 * th * (th - 1.3299890189751412) * (th**2 - 1.4364028541716351 * th + 0.7139816917074209),
        0.1017 * th**2 * (th**2 - 2.1966568338249754 * th + 1.2949852507374631),
        2.490627285651252793 * th**2 * (th**2 - 2.38535645472061657 * th + 1.57803468208092486),
        -16.54810288924490272 * (th - 1.21712927295533244) * (th - 0.61620406037800089) * th**2,
        47.37952196281928122 * (th - 1.203071208372362603) * (th - 0.658047292653547382) * th**2,
        -34.87065786149660974 * (th - 1.2) * (th - 0.666666666666666667) * th**2,
        2.5 * (th - 1) * (th - 0.6) * th**2,
    ])


def _rms(x):
    return float(np.sqrt(np.mean(x * x)))


def solve(ode, y0_tuple, args, t1, *, t0=0.0, rtol=1e-5, atol=1e-6, max_steps=10**6,
          const_dt=0.0, save_ts=None, return_steps=False, jump_ts=()):
    """Integrate `ode(t, state_tuple, args) -> tuple` from t0 to t1.

    Returns (ys_tuple with leading time axis, stats dict).  Unreached save slots stay inf.
    """
    shapes = [np.shape(np.asarray(c)) for c in y0_tuple]
    sizes = [int(np.prod(s)) for s in shapes]
    offs = np.concatenate([[0], np.cumsum(sizes)])

    def pack(tup):
        return np.concatenate([np.asarray(c, dtype=np.float64).ravel() for c in tup])

    def unpack(v):
        return tuple(v[offs[i]:offs[i + 1]].reshape(shapes[i]) for i in range(len(shapes)))

    def f(t, v):
        return pack(ode(t, unpack(v), args))

    if save_ts is None:
        save_ts = np.linspace(t0, t1, int(t1 // 1) + 1)
    save_ts = np.asarray(save_ts, dtype=np.float64)
    y = pack(y0_tuple)
    n = y.size
    out = np.full((save_ts.size, n), np.inf)
    t1 = float(t1)
    tprev = float(t0)
    f0 = f(tprev, y)
    if const_dt > 0:
        tnext = tprev + const_dt
    else:
        scale = atol + np.abs(y) * rtol
        d0, d1 = _rms(y / scale), _rms(f0 / scale)
        if d0 < 1e-5 or d1 < 1e-5:
            h0 = 1e-6
        else:
            h0 = 0.01 * (d0 / d1)
        f1 = f(tprev + h0, y + h0 * f0)
        d2 = _rms((f1 - f0) / scale) / h0
        md = max(d1, d2)
        h1 = max(1e-6, h0 * 1e-3) if md <= 1e-15 else (0.01 / md) ** (1 / 5)
        tnext = tprev + min(100 * h0, h1)
    jumps = np.sort(np.asarray(jump_ts, dtype=np.float64)) if (len(jump_ts) and not const_dt > 0) else None

    def clip_to_jumps(a, b):
        """ClipStepSizeController(jump_ts): searchsorted(side='right') indices of a and b; a jump in (a, b]
        ends the step at prevbefore(jump)."""
        i0, i1 = np.searchsorted(jumps, a, side="right"), np.searchsorted(jumps, b, side="right")
        if i0 < i1:
            return float(np.nextafter(jumps[min(i0, jumps.size - 1)], -np.inf)), True
        return b, False

    made_jump = False
    if jumps is not None:
        tnext, made_jump = clip_to_jumps(tprev, tnext)
    tnext = min(tnext, t1)
    n_steps = n_acc = n_rej = 0
    si = 0
    steps = []
    while tprev < t1 and n_steps < max_steps:
        h = tnext - tprev
        k = np.empty((7, n))
        k[0] = h * f0
        fl = f0
        for s in range(1, 7):
            ys = y + A[s - 1] @ k[:s]
            ts = tnext if C[s - 1] == 1.0 else tprev + C[s - 1] * h
            fl = f(ts, ys)
            k[s] = h * fl
        y1 = ys
        yerr = B_ERR @ k
        if const_dt > 0:
            keep, dt = True, const_dt
        else:
            yerr = np.where(np.isnan(yerr), np.inf, yerr)
            y1c = y if np.isnan(y1).any() else y1
            err = _rms(yerr / (atol + np.maximum(np.abs(y), np.abs(y1c)) * rtol))
            keep = err < 1
            with np.errstate(divide="ignore"):
                inv = np.float64(1.0) / np.float64(err)
            factor = float(np.clip(0.9 * inv ** 0.2, 1.0 if keep else 0.2, 10.0))
            dt = h * factor
        ntprev = tnext if keep else tprev
        next_made_jump = False
        if jumps is not None and keep and made_jump:
            ntprev = float(np.nextafter(tnext, np.inf))
        ntnext = ntprev + dt
        if jumps is not None:
            ntnext, next_made_jump = clip_to_jumps(ntprev, ntnext)
        ntprev = min(ntprev, t1)
        if ntnext > t1 - 1e-10:
            ntnext = t1 if keep else ntprev + 0.5 * (t1 - ntprev)
        n_steps += 1
        if keep:
            n_acc += 1
            while si < save_ts.size and save_ts[si] <= tnext:
                th = (save_ts[si] - tprev) / (1.0 if tnext == tprev else (tnext - tprev))
                out[si] = y + dense_weights(th) @ k
                si += 1
            steps.append((tprev, tnext))
            y, f0 = y1, fl
            if jumps is not None and made_jump:
                f0 = f(ntprev, y)  # no FSAL across a jump
        else:
            n_rej += 1
        if jumps is not None:
            made_jump = next_made_jump
        tprev, tnext = ntprev, ntnext
    stats = dict(result=int(tprev < t1), num_accepted_steps=n_acc, num_rejected_steps=n_rej,
                 num_steps=n_steps)
    ys_t = tuple(out[:, offs[i]:offs[i + 1]].reshape((save_ts.size,) + shapes[i])
                 for i in range(len(shapes)))
    if return_steps:
        return ys_t, stats, steps
    return ys_t, stats
