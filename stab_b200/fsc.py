"""Compressible Falkner-Skan-Cooke mean flow: the profile.<ind> / first.<ind> / second.<ind> tables that
the reference's `getmean` / `getmean2` read (getmean.f90:40-80, getmean2.f90:101-185).

The reference repository does not contain this solver: its test scripts call an external program `fsc`
(thesis/CFtest/run.sh, thesis/TStest/README.md) whose input decks are thesis/TStest/blasius.inp and
thesis/CFtest/fsc.inp (M_e, Re_delta1 / sweep angle / Hartree beta / T0, Tw, muw / converged f''(0), g'(0)).  This module
restates the published similarity problem those decks describe (Pr = 1, mu ~ T, wall at the recovery temperature:
constant total enthalpy) and is pinned against what the reference does ship:
  * TStest/profile.0 (M = 0.3, unswept, beta_h = 0) is reproduced to 1.4e-9 in y and 1e-12 in rho, u, T;
  * the deck's own converged wall values f''(0) = 1.2385480277606, g'(0) = 0.57111548449910 of the crossflow case
    (M = 0.3, 45 deg, beta_h = 1) are reproduced to every printed digit by shooting on the equations below;
  * with the generated crossflow profile the eigenvalues of thesis/CFtest/{time,space}.ref are matched to 3e-10 and the
    eigenfunction rows to 3e-8 (the reference CI's `ndiff` tolerance is 1e-8; the remainder is the unknown last digits
    of the external solver's table: 1e-10 in f''(0) moves the eigenfunction by 2e-7), see tests/test_fsc.py.

Similarity equations (xi: Levy-Lees / Stewartson variable of the chordwise flow, ' = d/dxi):
    f''' + f f'' + beta_h [ (1 - f'^2) + t_s (1 - g^2) ] = 0,   t_s = a sin^2(L) / (1 + a cos^2(L)),  a = (gamma-1)/2 M_e^2
    g''  + f g'                                          = 0
    f(0) = f'(0) = g(0) = 0,  f'(inf) = g(inf) = 1
with the velocities in body-fixed axes normalised by the total edge speed, u = cos(L) f', w = sin(L) g, and
    T/T_e = 1 + a (1 - u^2 - w^2),   rho/rho_e = T_e/T,   v = 0 (parallel flow; getmean.f90:75 zeroes it anyway).
The wall-normal coordinate is y = (1/delta1*) int_0^xi (T/T_e) dxi, where delta1* = int (T/T_e - u_s) dxi is the
displacement thickness of the velocity component along the edge streamline, u_s = cos^2(L) f' + sin^2(L) g: lengths are
in units of delta_1, the reference length of the decks' Re_delta1.  The table has n = 2000 rows on a uniform xi grid over
[0, 20] (the reference profile's grid), integrated with classical RK4 on that grid.
"""
from __future__ import annotations

import numpy as np

GAMMA = 1.4


def _coeffs(Me: float, lam_deg: float, gamma: float):
    lam = np.deg2rad(lam_deg)
    c, s = np.cos(lam), np.sin(lam)
    a = 0.5 * (gamma - 1.0) * Me * Me
    return c, s, a, a * s * s / (1.0 + a * c * c)


def _rhs(st, bh, c, s, a, ts):
    f, fp, fpp, g, gp = st[0], st[1], st[2], st[3], st[4]
    T = 1.0 + a * (1.0 - c * c * fp * fp - s * s * g * g)
    us = c * c * fp + s * s * g
    return np.array([fp, fpp, -f * fpp - bh * ((1.0 - fp * fp) + ts * (1.0 - g * g)), gp, -f * gp, T, T - us])


def _rk4(fpp0, gp0, bh, coef, n, xi_max):
    h = xi_max / (n - 1)
    out = np.empty((n, 7))
    st = np.array([0.0, 0.0, fpp0, 0.0, gp0, 0.0, 0.0])
    out[0] = st
    for i in range(1, n):
        k1 = _rhs(st, bh, *coef)
        k2 = _rhs(st + 0.5 * h * k1, bh, *coef)
        k3 = _rhs(st + 0.5 * h * k2, bh, *coef)
        k4 = _rhs(st + h * k3, bh, *coef)
        st = st + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        out[i] = st
    return out


def shoot(Me: float, lam_deg: float, beta_h: float, guess=None, gamma: float = GAMMA):
    """Wall values (f''(0), g'(0)) that satisfy the edge conditions.  Shooting is done on growing intervals (the
    problem is exponentially ill-conditioned in xi_max) with a Newton iteration on a finite-difference Jacobian."""
    from scipy.integrate import solve_ivp
    coef = _coeffs(Me, lam_deg, gamma)

    def res(p, X):
        sol = solve_ivp(lambda x, y: _rhs(np.append(y, [0.0, 0.0]), beta_h, *coef)[:5], [0.0, X], [0.0, 0.0, p[0], 0.0, p[1]],
                        rtol=1e-12, atol=1e-14, method="DOP853")
        return np.array([sol.y[1, -1] - 1.0, sol.y[3, -1] - 1.0])

    if guess is None:   # incompressible Falkner-Skan-Cooke wall shear, fitted on beta_h in [-0.15, 2]
        b = max(beta_h, -0.19)
        guess = (0.4696 * (1.0 + b / 0.1988) ** 0.53 if b < 0 else 0.4696 + 0.7630 * b ** 0.86, 0.4696 + 0.1010 * b ** 0.75 if b > 0 else 0.4696)
    p = np.array(guess, float)
    for X in (4.0, 6.0, 8.0, 10.0):
        for _ in range(30):
            r = res(p, X)
            if np.abs(r).max() < 1e-13:
                break
            J = np.empty((2, 2))
            for k in range(2):
                dp = np.zeros(2)
                dp[k] = 1e-7 * max(abs(p[k]), 1e-3)
                J[:, k] = (res(p + dp, X) - r) / dp[k]
            step = np.linalg.solve(J, -r)
            nrm = np.abs(step).max()
            if nrm > 0.2:
                step *= 0.2 / nrm
            p = p + step
    return float(p[0]), float(p[1])


def solve(Me: float, lam_deg: float = 0.0, beta_h: float = 0.0, wall=None, n: int = 2000, xi_max: float = 20.0,
          gamma: float = GAMMA, polish: bool = True) -> dict:
    """Integrates the similarity problem on the table grid.  `wall` = (f''(0), g'(0)) as in the fsc deck's last data
    line (converged values); None: they are found by `shoot`.  Returns the table and its y-derivatives."""
    coef = _coeffs(Me, lam_deg, gamma)
    c, s, a, ts = coef
    if wall is None:
        wall = shoot(Me, lam_deg, beta_h, gamma=gamma)
    if polish:
        # The edge conditions are imposed ON THE TABLE GRID at xi_max: a deck prints 14 digits of f''(0), and
        # d f'(xi_max) / d f''(0) ~ 4e2 ... 1e6, so the printed value alone leaves f'(xi_max) - 1 ~ 4e-8.
        p = np.array(wall, float)
        for _ in range(4):
            r = _rk4(p[0], p[1], beta_h, coef, n, xi_max)[-1, [1, 3]] - 1.0
            if np.abs(r).max() < 1e-14:
                break
            J = np.empty((2, 2))
            for k in range(2):
                dp = np.zeros(2)
                dp[k] = 1e-11
                J[:, k] = (_rk4(p[0] + dp[0], p[1] + dp[1], beta_h, coef, n, xi_max)[-1, [1, 3]] - 1.0 - r) / 1e-11
            p = p - np.linalg.solve(J, r)
        wall = (float(p[0]), float(p[1]))
    sol = _rk4(wall[0], wall[1], beta_h, coef, n, xi_max)
    f, fp, fpp, g, gp, Y, D = sol.T
    d1 = D[-1]
    u, w = c * fp, s * g
    T = 1.0 + a * (1.0 - u * u - w * w)
    rho = 1.0 / T
    # y-derivatives by the chain rule: d/dy = (delta1* / T) d/dxi
    fppp = -f * fpp - beta_h * ((1.0 - fp * fp) + ts * (1.0 - g * g))
    gpp = -f * gp
    u1, w1 = c * fpp, s * gp                      # d/dxi
    u2, w2 = c * fppp, s * gpp                    # d2/dxi2
    T1 = -2.0 * a * (u * u1 + w * w1)
    T2 = -2.0 * a * (u1 * u1 + u * u2 + w1 * w1 + w * w2)
    r1 = -T1 / T ** 2
    r2 = -T2 / T ** 2 + 2.0 * T1 * T1 / T ** 3
    m = d1 / T                                    # dxi/dy
    m1 = -d1 * T1 / T ** 2                        # d(dxi/dy)/dxi

    def dy(q1):
        return q1 * m

    def d2y(q1, q2):
        return (q2 * m + q1 * m1) * m
    z = np.zeros(n)
    table = np.stack([Y / d1, rho, u, z, w, T], axis=1)
    first = np.stack([Y / d1, dy(r1), dy(u1), z, dy(w1), dy(T1)], axis=1)
    second = np.stack([Y / d1, d2y(r1, r2), d2y(u1, u2), z, d2y(w1, w2), d2y(T1, T2)], axis=1)
    return dict(table=table, first=first, second=second, wall=wall, delta1_star=d1, edge=(fp[-1], g[-1]), xi=np.linspace(0.0, xi_max, n))


def read_deck(text: str) -> dict:
    """The `fsc` input deck (thesis/CFtest/fsc.inp): M_e, Re_delta1 / Lambda_e (deg) / beta_h / T0, Tw, muw / f'', g' / delta/L."""
    rows = []
    for line in text.splitlines():
        body = line.split("!")[0].replace(",", " ").split()
        try:
            rows.append([float(v) for v in body])
        except ValueError:
            break
        if len(rows) == 6:
            break
    if len(rows) < 5 or not all(rows[:5]):
        raise ValueError("fsc deck: expected 6 data lines")
    return dict(Me=rows[0][0], Re=rows[0][1], lam_deg=rows[1][0], beta_h=rows[2][0], flags=tuple(rows[3]), wall=(rows[4][0], rows[4][1]))


def profile_from_deck(text: str, reshoot: bool = False, **kw) -> dict:
    d = read_deck(text)
    if tuple(int(v) for v in d["flags"][:3]) != (1, 1, 1):
        raise ValueError("fsc deck: only T0, Tw, muw = 1, 1, 1 (wall at the recovery temperature, mu ~ T) is supported")
    return solve(d["Me"], d["lam_deg"], d["beta_h"], wall=None if reshoot else d["wall"], **kw)


def format_table(table: np.ndarray) -> str:
    """profile.<ind> text: rows `y rho u v w T` (getmean.f90:71-80)."""
    return "\n".join(" ".join(f"{v: .13E}" for v in row) for row in table) + "\n"


def write_profile(prefix_dir: str, ind: int, sol: dict, derivatives: bool = False):
    import os
    with open(os.path.join(prefix_dir, f"profile.{ind}"), "w") as fh:
        fh.write(format_table(sol["table"]))
    if derivatives:
        with open(os.path.join(prefix_dir, f"first.{ind}"), "w") as fh:
            fh.write(format_table(sol["first"]))
        with open(os.path.join(prefix_dir, f"second.{ind}"), "w") as fh:
            fh.write(format_table(sol["second"]))
