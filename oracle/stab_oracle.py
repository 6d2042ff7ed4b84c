"""CPU ORACLE for the stab hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, in numpy + scipy-LAPACK, the arithmetic of the reference
`sscollis/stab` Chebyshev-collocation path (operator assembly + dense complex
eigensolve).  It exists so that `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` have something to check the
CUDA path against.  Nothing in `stab_b200/` (the product) may import it.

Parity pin: every function below cites the reference file:line it follows
(paths relative to the reference checkout) and the module is pinned against the
reference's own golden vectors in `tests/test_oracle_golden.py`
(`test/space.1`, `thesis/TStest/time.ref`, `thesis/TStest/space.ref`,
`FSCtest/space.ref`, `CFtest/space.ref`, README known answers) -- see
`tests/golden/README.md`.  The reference itself (Fortran) cannot be compiled in
this image (no Fortran compiler), so LAPACK is scipy's bundled OpenBLAS
(`zgesv`, `zgetrf`, `zgetrs`, `zgeev`), which is also what the reference links
(`gcc.mak:15-25`, unpinned OpenBLAS).  The NR routine `PIKSR2` (private,
unpinned) is replaced by a stable ascending sort, which is what a straight
insertion sort is.
"""
from __future__ import annotations

import dataclasses
import io
import math
import struct
from typing import Optional

import numpy as np

NDOF = 5  # stuff.f90:54
IM = 1j


# ----------------------------------------------------------------------------
# module stuff (stuff.f90:11-59) as a params object
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class Params:
    mattyp: int = 0
    T0: float = 0.0
    Ma: float = 0.0
    Re: float = 0.0
    Pr: float = 1.0
    ny: int = 1
    yi: float = 0.0
    ymax: float = 0.0
    ievec: int = 1
    ider: int = 1
    top: int = 0
    wall: int = 0
    wallt: int = 0
    curve: int = 0
    itype: int = 1
    alpha: complex = 0j
    beta: complex = 0j
    omega: complex = 0j
    ind: int = 0
    x: float = 0.0
    # fluid properties, stuff.f90:45-46
    gamma: float = 1.4
    gamma1: float = 0.4
    cv: float = 716.5
    cp: float = 1003.1
    Rgas: float = 286.6
    datmat: tuple = (1.0, 0.0, 0.0)
    # edge state, stuff.f90:28-29 (filled by finish())
    Te: float = 1.0
    rmue: float = 1.0
    rlme: float = 1.0
    cone: float = 1.0

    def finish(self) -> "Params":
        """input.f90:19-44: material constants and edge properties.

        Quirk q1: Te is computed from T0 *before* Ma is read, i.e. with
        Ma = 0, so Te = T0 for Sutherland; for constant mu Te is never set and every
        use multiplies a zero -- we use 1.0.
        """
        if self.mattyp == 1:
            self.Te = self.T0  # Ma still 0 at input.f90:25
            self.datmat = (1.715336725523065e-05, 273.0, 110.4)  # input.f90:29-31
        else:
            self.Te = 1.0
            self.datmat = (1.0, 0.0, 0.0)
        mu, lm, con, *_ = getmat(np.array([self.Te]), self)  # input.f90:43
        self.rmue, self.rlme, self.cone = float(mu[0]), float(lm[0]), float(con[0])
        return self

    @property
    def navier(self) -> bool:
        """temporal.f90:88-91 / spatial.f90:89-92."""
        return not (self.Re >= 1.0e98 or self.Re == 0.0)


def _tokens(line: str):
    line = line.split("!")[0].replace(",", " ")
    return line.split()


def read_deck(text: str) -> Params:
    """Positional stdin deck: input.f90:15-122 then stab.f90:46-55.

    Only itype 1 (temporal) and 2 (spatial) single-point decks are parsed here;
    the sweep tails (itype 7/8) are handled by `read_sweep_deck`.
    """
    lines = [ln for ln in text.splitlines() if ln.strip()]
    it = iter(lines)
    p = Params()
    p.mattyp = int(_tokens(next(it))[0])
    if p.mattyp == 1:
        p.T0 = float(_tokens(next(it))[0])
    t = _tokens(next(it)); p.Ma, p.Re, p.Pr = float(t[0]), float(t[1]), float(t[2])
    t = _tokens(next(it)); p.ny, p.yi, p.ymax = int(t[0]), float(t[1]), float(t[2])
    p.ievec = int(_tokens(next(it))[0])
    p.ider = 0 if int(_tokens(next(it))[0]) == 0 else 1
    t = _tokens(next(it)); p.top, p.wall, p.wallt, p.curve = (int(v) for v in t[:4])
    p.itype = int(_tokens(next(it))[0])
    p._rest = list(it)  # type: ignore[attr-defined]
    rest = iter(p._rest)  # type: ignore[attr-defined]
    if p.itype in (1, 3):
        t = _tokens(next(rest)); p.alpha = complex(float(t[0]), float(t[1]))
        t = _tokens(next(rest)); p.beta = complex(float(t[0]), float(t[1]))
    elif p.itype in (2, 4):
        t = _tokens(next(rest)); p.omega = complex(float(t[0]), float(t[1]))
        t = _tokens(next(rest)); p.beta = complex(float(t[0]), float(t[1]))
    if p.itype in (1, 2, 3, 4):
        p.ind = int(_tokens(next(rest))[0])
        if p.itype in (2, 4):
            p.x = float(_tokens(next(rest))[0])
    p._tail = list(rest)  # type: ignore[attr-defined]
    return p.finish()


# ----------------------------------------------------------------------------
# getmat.f90:2-37
# ----------------------------------------------------------------------------
PT66 = 6.6666666666666666666e-1  # stuff.f90:35


def getmat(t, p: Params):
    t = np.asarray(t, dtype=np.float64)
    d1, d2, d3 = p.datmat
    if p.mattyp == 0:
        mu = np.full_like(t, d1)
        dmu = np.zeros_like(t)
        d2mu = np.zeros_like(t)
    else:  # Sutherland, getmat.f90:19-25
        mu = d1 * t / d2 * np.sqrt(t / d2) * (d2 + d3) / (t + d3)
        dmu = (d1 * (3.0 * d3 + t) * (d3 + d2) * np.sqrt(t / d2)) / (2.0 * (d3 + t) ** 2 * d2)
        d2mu = (d1 * (3.0 * d3 ** 2 - 6.0 * d3 * t - t ** 2) * (d3 + d2)) / (
            4.0 * (d3 + t) ** 3 * np.sqrt(t / d2) * d2 ** 2)
    con, dcon, d2con = mu * p.cp / p.Pr, dmu * p.cp / p.Pr, d2mu * p.cp / p.Pr
    lm, dlm, d2lm = -PT66 * mu, -PT66 * dmu, -PT66 * d2mu
    return mu, lm, con, dmu, d2mu, dlm, d2lm, dcon, d2con


# ----------------------------------------------------------------------------
# sgengrid.f90:15-45 (algebraic and Streett maps; tanh map reads stdin -> out of scope)
# ----------------------------------------------------------------------------
def sgengrid(ny: int, yi: float, ymax: float):
    dth = math.pi / float(ny - 1)
    eta = np.array([math.cos(float(i) * dth) for i in range(ny)])
    y = np.empty(ny); deta = np.empty(ny); d2eta = np.empty(ny)
    if yi == 0.0:
        raise NotImplementedError("tanh mapping (Yi=0) reads stdin per call: out of scope")
    if ymax == 0.0:  # sgengrid.f90:28-38
        L = yi
        for i in range(ny):
            deta[i] = (eta[i] - 1.0) ** 2 / (2.0 * L)
            d2eta[i] = (eta[i] - 1.0) ** 3 / (2.0 * L ** 2)
            y[i] = L * (1.0 + eta[i]) / (1.0 - eta[i]) if eta[i] != 1.0 else 1.0e99
    else:  # Streett, sgengrid.f90:39-45
        for i in range(ny):
            y[i] = ymax * yi * (1.0 + eta[i]) / (1.0 + 2.0 * yi - eta[i])
            deta[i] = (2.0 * yi + 1.0 - eta[i]) ** 2 / (2.0 * ymax * yi * (yi + 1.0))
            d2eta[i] = -0.5 * (2.0 * yi + 1.0 - eta[i]) ** 3 / (ymax * yi * (yi + 1.0)) ** 2
    return y, eta, deta, d2eta


# ----------------------------------------------------------------------------
# spline.f90:2-75
# ----------------------------------------------------------------------------
def spline(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    n = len(x)
    a = np.zeros(n); b = np.zeros(n); c = np.zeros(n); r = np.zeros(n); fdp = np.zeros(n)
    c[0] = x[1] - x[0]
    for i in range(1, n - 1):
        c[i] = x[i + 1] - x[i]
        a[i] = c[i - 1]
        b[i] = 2.0 * (a[i] + c[i])
        r[i] = 6.0 * ((y[i + 1] - y[i]) / c[i] - (y[i] - y[i - 1]) / c[i - 1])
    b[1] += c[0]          # ALAMDA = 1 (cantilever), spline.f90:29
    b[n - 2] += c[n - 2]  # spline.f90:30
    for i in range(2, n - 1):
        t = a[i] / b[i - 1]
        b[i] -= t * c[i - 1]
        r[i] -= t * r[i - 1]
    fdp[n - 2] = r[n - 2] / b[n - 2]
    for i in range(2, n - 1):
        k = n - 1 - i
        fdp[k] = (r[k] - c[k] * fdp[k + 1]) / b[k]
    fdp[0] = fdp[1]
    fdp[n - 1] = fdp[n - 2]
    return fdp


def speval(x, y, fdp, xx: float) -> float:
    """spline.f90:49-75: linear search for the first interval with xx <= x(i+1)."""
    n = len(x)
    i = int(np.searchsorted(x[1:], xx, side="left"))  # first i with xx <= x[i+1]
    if i > n - 2:
        i = n - 1  # Fortran loop falls through with I = N: reference would read out of bounds
        raise ValueError("speval: xx beyond table")
    dxm = xx - x[i]
    dxp = x[i + 1] - xx
    dl = x[i + 1] - x[i]
    return (fdp[i] * dxp * (dxp * dxp / dl - dl) / 6.0
            + fdp[i + 1] * dxm * (dxm * dxm / dl - dl) / 6.0
            + y[i] * dxp / dl + y[i + 1] * dxm / dl)


def read_profile(text: str) -> np.ndarray:
    """getmean.f90:40-80: rows `y rho u v w T`, '#' lines skipped, v forced to 0."""
    rows = []
    for ln in text.splitlines():
        if not ln.strip() or ln[0] == "#":
            continue
        rows.append([float(v.replace("D", "E").replace("d", "e")) for v in ln.split()[:6]])
    tab = np.array(rows, dtype=np.float64)
    tab[:, 3] = 0.0  # getmean.f90:75 parallel-flow assumption
    return tab


def getmean(tab: np.ndarray, y: np.ndarray) -> np.ndarray:
    """getmean.f90:81-111 -> vm(ny,5)."""
    ym = tab[:, 0]
    ny = len(y)
    vm = np.empty((ny, NDOF))
    for k in range(NDOF):
        vt = tab[:, 1 + k]
        vs = spline(ym, vt)
        for j in range(ny):
            vm[j, k] = speval(ym, vt, vs, y[j]) if y[j] <= ym[-1] else vt[-1]
    return vm


# ----------------------------------------------------------------------------
# chebyd.f90:2-60
# ----------------------------------------------------------------------------
def chebyd(N: int) -> np.ndarray:
    """D(0:N,0:N) with the brute-force Lagrange-weight formula, same operation order."""
    pi = math.acos(-1.0)
    x = np.array([math.cos(pi * float(j) / float(N)) for j in range(N + 1)])
    a = np.ones(N + 1)
    dd = np.zeros(N + 1)
    for k in range(N + 1):  # sequential in k for every j, as chebyd.f90:35-49
        diff = x - x[k]
        mask = np.arange(N + 1) != k
        a[mask] = a[mask] * diff[mask]
        dd[mask] = dd[mask] + 1.0 / diff[mask]
    D = np.empty((N + 1, N + 1))
    for j in range(N + 1):
        for k in range(N + 1):
            if k != j:
                D[j, k] = a[j] / (a[k] * (x[j] - x[k]))
        D[j, j] = dd[j]
    return D


# ----------------------------------------------------------------------------
# circh.f90:35-188 (curve=2), scalar prologue + vector metrics
# ----------------------------------------------------------------------------
def circh(radius_in: float, r: np.ndarray):
    """Returns (s_out, h, dhds, dhdr, dhdsr, dhdrr).  The deck value `x` is the radius
    and the routine overwrites it with 0 (circh.f90:47-49)."""
    n = len(r)
    one = 1.0
    if radius_in == -1.0:
        z = np.zeros(n)
        return radius_in, np.ones(n), z.copy(), z.copy(), z.copy(), z.copy()
    radius = radius_in
    s = 0.0
    infty = 1.0e30
    # xcloc(0, s): circh.f90:212-225
    th1 = math.atan2(math.sqrt(radius ** 2 - 0.0 ** 2), 0.0)
    xl = radius * math.cos(th1 - s)
    yl = math.sqrt(radius ** 2 - xl ** 2)
    th = math.atan2(-xl, math.sqrt(radius ** 2 - xl ** 2))
    bn1 = -math.sin(th)
    bn2 = math.cos(th)
    dydx = -infty if xl == 0.0 else (-xl) / yl
    dxbds = one / math.sqrt(one + dydx ** 2)
    dxdy = -infty if xl == 0.0 else yl / (-xl)
    dybds = (one if xl <= 0.0 else -one) / math.sqrt(dxdy ** 2 + one)
    if yl == 0.0:
        dx, dy = 0.0, -xl
        ddxdx, ddydx, ddxdy, ddydy = -infty, -one, one, 0.0
        d2dxdx2, d2dydx2, d2dxdy2 = -infty, 0.0, 0.0
        d2dydy2 = (xl ** 2 + yl ** 2) / xl ** 3
    elif xl == 0.0:
        dx, dy = yl, 0.0
        ddxdx, ddydx, ddxdy, ddydy = 0.0, -one, one, infty
        d2dxdx2 = -(yl ** 2 + xl ** 2) / yl ** 3
        d2dydx2, d2dxdy2, d2dydy2 = 0.0, 0.0, infty
    else:
        dx, dy = yl, -xl
        ddxdx, ddydx, ddxdy, ddydy = -xl / yl, -one, one, yl / xl
        d2dxdx2 = -(yl ** 2 + xl ** 2) / yl ** 3
        d2dydx2, d2dxdy2 = 0.0, 0.0
        d2dydy2 = (xl ** 2 + yl ** 2) / xl ** 3
    q = dx ** 2 + dy ** 2
    if abs(bn1) > abs(bn2):
        dbn1 = (-ddydy / q ** 0.5 + 0.5 * dy * (2 * dx * ddxdy + 2 * dy * ddydy) / q ** 1.5) * dybds
        dbn2 = (ddxdy / q ** 0.5 - 0.5 * dx * (2 * dx * ddxdy + 2 * dy * ddydy) / q ** 1.5) * dybds
        d2xdy2 = -(xl ** 2 + yl ** 2) / xl ** 3
        sgn = -1.0 if xl <= 0 else 1.0
        d2ybds2 = sgn * (one + dxdy ** 2) ** (-1.5) * dxdy * d2xdy2 * dybds
        d2xbds2 = d2xdy2 * dybds ** 2 + dxdy * d2ybds2
        d2bn1 = ((ddxdy * (dy * ddxdy - dx * ddydy) + dx * (dy * d2dxdy2 - dx * d2dydy2)) / q ** 1.5
                 - (3.0 * dx * (dy * ddxdy - dx * ddydy) * (dx * ddxdy + dy * ddydy)) / q ** 2.5) * dybds ** 2 \
            + (-ddydy / q ** 0.5 + 0.5 * dy * (2 * dx * ddxdy + 2 * dy * ddydy) / q ** 1.5) * d2ybds2
        d2bn2 = ((ddydy * (dy * ddxdy - dx * ddydy) + dy * (dy * d2dxdy2 - dx * d2dydy2)) / q ** 1.5
                 - (3.0 * dy * (dy * ddxdy - dx * ddydy) * (dx * ddxdy + dy * ddydy)) / q ** 2.5) * dybds ** 2 \
            + (ddxdy / q ** 0.5 - 0.5 * dx * (2 * dx * ddxdy + 2 * dy * ddydy) / q ** 1.5) * d2ybds2
    else:
        dbn1 = (-ddydx / q ** 0.5 + 0.5 * dy * (2 * dx * ddxdx + 2 * dy * ddydx) / q ** 1.5) * dxbds
        dbn2 = (ddxdx / q ** 0.5 - 0.5 * dx * (2 * dx * ddxdx + 2 * dy * ddydx) / q ** 1.5) * dxbds
        d2ydx2 = -(xl ** 2 + yl ** 2) / yl ** 3
        d2xbds2 = -(one + dydx ** 2) ** (-1.5) * dydx * d2ydx2 * dxbds
        d2ybds2 = d2ydx2 * dxbds ** 2 + dydx * d2xbds2
        d2bn1 = ((ddxdx * (dy * ddxdx - dx * ddydx) + dx * (dy * d2dxdx2 - dx * d2dydx2)) / q ** 1.5
                 - (3.0 * dx * (dy * ddxdx - dx * ddydx) * (dx * ddxdx + dy * ddydx)) / q ** 2.5) * dxbds ** 2 \
            + (-ddydx / q ** 0.5 + 0.5 * dy * (2 * dx * ddxdx + 2 * dy * ddydx) / q ** 1.5) * d2xbds2
        d2bn2 = ((ddydx * (dy * ddxdx - dx * ddydx) + dy * (dy * d2dxdx2 - dx * d2dydx2)) / q ** 1.5
                 - (3.0 * dy * (dy * ddxdx - dx * ddydx) * (dx * ddxdx + dy * ddydx)) / q ** 2.5) * dxbds ** 2 \
            + (ddxdx / q ** 0.5 - 0.5 * dx * (2 * dx * ddxdx + 2 * dy * ddydx) / q ** 1.5) * d2xbds2
    if xl == 0.0:
        d2bn1 = 0.0
    a = dxbds + r * dbn1
    b = dybds + r * dbn2
    h = np.sqrt(a ** 2 + b ** 2)
    dads = d2xbds2 + r * d2bn1
    dbds = d2ybds2 + r * d2bn2
    dhds = (a * dads + b * dbds) / h
    dhdr = (a * dbn1 + b * dbn2) / h
    dhdrr = (-dhdr ** 2 + dbn1 ** 2 + dbn2 ** 2) / h
    dhdsr = -dhds / h ** 2 * (a * dbn1 + b * dbn2) + (dads * dbn1 + a * d2bn1 + dbds * dbn2 + b * d2bn2) / h
    return s, h, dhds, dhdr, dhdsr, dhdrr


# ----------------------------------------------------------------------------
# derivative operators and mean-flow gradients
# ----------------------------------------------------------------------------
def deriv_ops(ny: int, wallt: int):
    """temporal.f90:135-144 / spatial.f90:156-165."""
    D1 = chebyd(ny - 1)
    D2 = D1 @ D1
    if wallt == 2:
        Dt1 = D1.copy(); Dt1[ny - 1, :] = 0.0
        Dt2 = D1 @ Dt1
    else:
        Dt1, Dt2 = D1, D2
    return D1, D2, Dt1, Dt2


def mean_gradients(vm, D1, D2, deta, d2eta):
    """temporal.f90:148-179 / spatial.f90:169-195 (ider=1)."""
    g2 = D1 @ vm
    g22 = D2 @ vm
    g22 = g22 * (deta ** 2)[:, None] + g2 * d2eta[:, None]
    g2 = g2 * deta[:, None]
    return g2, g22


def _zeros_tables(ny):
    names = ("G", "A", "B", "C", "D", "Vxx", "Vxy", "Vyy", "Vxz", "Vyz", "Vzz")
    return {k: np.zeros((ny, NDOF, NDOF)) for k in names}


def _material(tm, p: Params):
    """getmat call + nondimensionalisation, temporal.f90:262-277 / spatial.f90:307-322."""
    mu, lm, con, dmu, d2mu, dlm, d2lm, dcon, d2con = getmat(tm * p.Te, p)
    mu = mu / p.rmue; dmu = dmu * p.Te / p.rmue; d2mu = d2mu * p.Te ** 2 / p.rmue
    con = con / p.cone; dcon = dcon * p.Te / p.cone; d2con = d2con * p.Te ** 2 / p.cone
    lm = lm / p.rlme; dlm = dlm * p.Te / p.rlme; d2lm = d2lm * p.Te ** 2 / p.rlme
    return mu, dmu, d2mu, lm, dlm, d2lm, con, dcon, d2con


def tables_temporal(vm, g2vm, g22vm, p: Params):
    """The 11 real coefficient tables of temporal.f90:206-598 (index 0-based [i, eq, var])."""
    ny = vm.shape[0]
    t = _zeros_tables(ny)
    G, A, B, C, D = t["G"], t["A"], t["B"], t["C"], t["D"]
    Vxx, Vxy, Vyy, Vxz, Vyz, Vzz = t["Vxx"], t["Vxy"], t["Vyy"], t["Vxz"], t["Vyz"], t["Vzz"]
    rho, u1, u2, u3, tm = (vm[:, k].copy() for k in range(5))
    z = np.zeros(ny)
    gam, gam1, Ma, Re, Pr = p.gamma, p.gamma1, p.Ma, p.Re, p.Pr
    # gradients: only y-derivatives are non-zero (parallel flow), temporal.f90:148-170
    gum = np.zeros((ny, 3, 3))
    gum[:, 0, 1] = g2vm[:, 1]; gum[:, 1, 1] = g2vm[:, 2]; gum[:, 2, 1] = g2vm[:, 3]
    divum = gum[:, 0, 0] + gum[:, 1, 1] + gum[:, 2, 2]
    grho = np.stack([z, g2vm[:, 0], z], axis=1)
    gt = np.stack([z, g2vm[:, 4], z], axis=1)
    fact = 1.0 / (gam * Ma ** 2)
    gp = np.stack([fact * (grho[:, k] * tm + rho * gt[:, k]) for k in range(3)], axis=1)
    g1div = z.copy()
    g2div = g22vm[:, 2].copy()   # g12vm(:,2)=0 + g22vm(:,3), temporal.f90:240
    g3div = z.copy()
    S = 0.5 * (gum + gum.transpose(0, 2, 1))
    S1jj = 0.5 * g22vm[:, 1]                      # temporal.f90:251-252
    S2jj = 0.5 * (g22vm[:, 2] + g22vm[:, 2])      # temporal.f90:254-255
    S3jj = 0.5 * g22vm[:, 3]                      # temporal.f90:257-258
    mu, dmu, d2mu, lm, dlm, d2lm, con, dcon, d2con = _material(tm, p)
    gmu = [dmu * gt[:, k] for k in range(3)]; gdmu = [d2mu * gt[:, k] for k in range(3)]
    gcon = [dcon * gt[:, k] for k in range(3)]; gdcon = [d2con * gt[:, k] for k in range(3)]
    glm = [dlm * gt[:, k] for k in range(3)]; gdlm = [d2lm * gt[:, k] for k in range(3)]
    gdiv = [g1div, g2div, g3div]
    um = [u1, u2, u3]
    gm2 = gam * Ma ** 2

    # continuity, temporal.f90:313-327
    G[:, 0, 0] = 1.0
    A[:, 0, 0] = u1; A[:, 0, 1] = rho
    B[:, 0, 0] = u2; B[:, 0, 2] = rho
    C[:, 0, 0] = u3; C[:, 0, 3] = rho
    D[:, 0, 0] = divum; D[:, 0, 1] = grho[:, 0]; D[:, 0, 2] = grho[:, 1]; D[:, 0, 3] = grho[:, 2]

    # momentum x_k (k = 0,1,2 -> equations 1,2,3): temporal.f90:331-526
    T3 = (A, B, C)
    for k in range(3):
        e = 1 + k
        G[:, e, e] = rho
        for d in range(3):
            T3[d][:, e, e] = rho * um[d]
        T3[k][:, e, 0] = tm / gm2
        T3[k][:, e, 4] = rho / gm2
        D[:, e, 0] = u1 * gum[:, k, 0] + u2 * gum[:, k, 1] + u3 * gum[:, k, 2] + gt[:, k] / gm2
        D[:, e, 1] = rho * gum[:, k, 0]; D[:, e, 2] = rho * gum[:, k, 1]; D[:, e, 3] = rho * gum[:, k, 2]
        D[:, e, 4] = grho[:, k] / gm2
    if p.navier:
        Vd = ((Vxx, Vxy, Vxz), (Vxy, Vyy, Vyz), (Vxz, Vyz, Vzz))  # second-derivative table for (d1,d2)
        Sjj = (S1jj, S2jj, S3jj)
        for k in range(3):
            e = 1 + k
            # (viscous lambda)
            fact = p.rlme / (p.rmue * Re)
            for d in range(3):
                # -fact*g_k(lm) * d(u_d)/dx_d : first-derivative table of direction d, column u_d
                T3[d][:, e, 1 + d] -= fact * glm[k]
            T3[k][:, e, 4] -= fact * dlm * divum
            D[:, e, 4] -= fact * (gdlm[k] * divum + dlm * gdiv[k])
            for d in range(3):
                Vd[k][d][:, e, 1 + d] = fact * lm   # assignment (=), e.g. temporal.f90:363-367
            # (viscous mu)
            fact = 1.0 / Re
            for d in range(3):
                # -fact * g_d(mu) * (du_k/dx_d + du_d/dx_k); the d == k entry is written
                # once with a factor two, as the reference does (e.g. temporal.f90:373)
                if d == k:
                    T3[k][:, e, e] -= fact * 2.0 * gmu[k]
                else:
                    T3[d][:, e, e] -= fact * gmu[d]
                    T3[k][:, e, 1 + d] -= fact * gmu[d]
                T3[d][:, e, 4] -= fact * dmu * 2.0 * S[:, k, d]
            D[:, e, 4] -= fact * 2.0 * (gdmu[0] * S[:, k, 0] + gdmu[1] * S[:, k, 1]
                                        + gdmu[2] * S[:, k, 2] + dmu * Sjj[k])
            for d in range(3):
                if d == k:
                    Vd[k][k][:, e, e] += fact * 2.0 * mu
                else:
                    Vd[d][d][:, e, e] += fact * mu          # mu * Laplacian u_k
                    Vd[k][d][:, e, 1 + d] += fact * mu      # mu * d/dx_k (div u)

    # energy, temporal.f90:530-598
    G[:, 4, 0] = -gam1 * tm / gam
    G[:, 4, 4] = rho / gam
    for d in range(3):
        T3[d][:, 4, 0] = -gam1 * um[d] * tm / gam
        T3[d][:, 4, 4] = rho * um[d] / gam
    D[:, 4, 0] = 1.0 / gam * (u1 * gt[:, 0] + u2 * gt[:, 1] + u3 * gt[:, 2])
    for d in range(3):
        D[:, 4, 1 + d] = rho * gt[:, d] - gam1 * Ma ** 2 * gp[:, d]
    D[:, 4, 4] = -gam1 / gam * (u1 * grho[:, 0] + u2 * grho[:, 1] + u3 * grho[:, 2])
    if p.navier:
        fact = 1.0 / (Pr * Re)
        for d in range(3):
            T3[d][:, 4, 4] -= fact * (gcon[d] + dcon * gt[:, d])
        D[:, 4, 4] -= fact * (gdcon[0] * gt[:, 0] + gdcon[1] * gt[:, 1] + gdcon[2] * gt[:, 2]
                              + dcon * (0.0 + g22vm[:, 4] + 0.0))
        Vxx[:, 4, 4] = fact * con; Vyy[:, 4, 4] = fact * con; Vzz[:, 4, 4] = fact * con
        fact = gam1 * Ma ** 2 * p.rlme / (Re * p.rmue)
        for d in range(3):
            T3[d][:, 4, 1 + d] -= fact * 2.0 * lm * divum
        D[:, 4, 4] -= fact * dlm * divum * divum
        fact = 2.0 * gam1 * Ma ** 2 / Re
        for d in range(3):
            for k in range(3):
                T3[d][:, 4, 1 + k] -= fact * 2.0 * mu * S[:, k, d]
        D[:, 4, 4] -= fact * dmu * np.sum(S ** 2, axis=(1, 2))
    return t


def assemble_temporal(p: Params, vm, deta, d2eta, g2vm=None, g22vm=None):
    """A0, B0 of temporal.f90:604-752 for p.alpha, p.beta.  Returns (A0, B0, tables)."""
    ny = p.ny
    n = NDOF * ny
    D1, D2, Dt1, Dt2 = deriv_ops(ny, p.wallt)
    if p.ider or g2vm is None:
        g2vm, g22vm = mean_gradients(vm, D1, D2, deta, d2eta)
    t = tables_temporal(vm, g2vm, g22vm, p)
    al, be = p.alpha, p.beta
    Dh = (t["D"] + IM * al * t["A"] + IM * be * t["C"]
          + al ** 2 * t["Vxx"] + al * be * t["Vxz"] + be ** 2 * t["Vzz"])
    Bh = t["B"] - IM * al * t["Vxy"] - IM * be * t["Vyz"]
    Bh = Bh * deta[:, None, None] - t["Vyy"] * d2eta[:, None, None]
    Vyy = t["Vyy"] * (deta ** 2)[:, None, None]
    G = t["G"]
    A0 = np.zeros((ny, NDOF, ny, NDOF), dtype=np.complex128)
    B0 = np.zeros((ny, NDOF, ny, NDOF), dtype=np.complex128)
    # B0, temporal.f90:630-662
    for i in range(ny):
        if i == 0 or i == ny - 1:
            for d in range(NDOF):
                B0[i, d, i, d] = IM
        else:
            B0[i, :, i, :] = IM * G[i]
    if p.wallt == 2:
        B0[ny - 1, 4, ny - 1, :] = IM * G[ny - 1, 4, :]
    elif p.wallt != 0:
        raise ValueError("Illegal value of wallt")
    # A0, temporal.f90:666-752
    def first_order_row(i, eq):
        A0[i, eq, :, :] += Bh[i, eq, None, :] * D1[i, :, None]
        A0[i, eq, i, :] += Dh[i, eq, :]
    first_order_row(0, 0)
    for i in range(1, ny - 1):
        A0[i, :, :, :] += (Bh[i][:, None, :] * D1[i][None, :, None]
                           - Vyy[i][:, None, :] * D2[i][None, :, None])
        A0[i, :, i, :] += Dh[i]
    first_order_row(ny - 1, 0)
    if p.wallt == 2:
        i = ny - 1
        # quirk q4: D1 (not D2) multiplies Vyy for dof 1-4, temporal.f90:733-735
        A0[i, 4, :, :4] += Bh[i, 4, None, :4] * D1[i, :, None] - Vyy[i, 4, None, :4] * D1[i, :, None]
        A0[i, 4, i, :4] += Dh[i, 4, :4]
        A0[i, 4, :, 4] += Bh[i, 4, 4] * Dt1[i, :] - Vyy[i, 4, 4] * Dt2[i, :]
        A0[i, 4, i, 4] += Dh[i, 4, 4]
    return A0.reshape(n, n), B0.reshape(n, n), t


def stable_sort_by_imag(vals: np.ndarray) -> np.ndarray:
    """PIKSR2 replacement: ascending, stable (temporal.f90:844-855)."""
    return np.argsort(vals.imag, kind="stable")


def scale_columns_maxabs(evec: np.ndarray) -> np.ndarray:
    """temporal.f90:867-879: divide by the first-encountered entry of maximum |.|."""
    out = evec.copy()
    mag = np.abs(out)
    idx = np.argmax(mag, axis=0)  # first maximum
    sc = out[idx, np.arange(out.shape[1])]
    nz = sc != 0
    out[:, nz] = out[:, nz] / sc[nz]
    return out


def _zgeev(M: np.ndarray, want_vectors: bool, as_coded: bool):
    from scipy.linalg import lapack
    n = M.shape[0]
    kw = dict(compute_vl=0, compute_vr=1 if want_vectors else 0, overwrite_a=1)
    if as_coded:
        kw["lwork"] = 2 * n  # temporal.f90:793
    w, vl, vr, info = lapack.zgeev(np.asfortranarray(M), **kw)
    return w, vr, info


def solve_temporal(p: Params, vm, deta, d2eta, g2vm=None, g22vm=None, want_vectors=True,
                   as_coded=True):
    """temporal.f90:604-879.  Returns dict(omg sorted, evec scaled, info, M=B0^-1 A0)."""
    from scipy.linalg import lapack
    A0, B0, _ = assemble_temporal(p, vm, deta, d2eta, g2vm, g22vm)
    lu, piv, M, info = lapack.zgesv(np.asfortranarray(B0), np.asfortranarray(A0))  # :774
    if info != 0:
        return dict(info=info)
    Mkeep = M.copy()
    w, vr, info = _zgeev(M, want_vectors, as_coded)
    order = stable_sort_by_imag(w)
    omg = w[order]
    out = dict(omg=omg, info=info, M=Mkeep, A0=A0, B0=B0)
    if want_vectors:
        out["evec"] = scale_columns_maxabs(vr[:, order])
    return out


# ----------------------------------------------------------------------------
# spatial
# ----------------------------------------------------------------------------
def tables_spatial(vm, g2vm, g22vm, hm, p: Params):
    """spatial.f90:223-673.  hm = (h, dhds, dhdr, dhdsr, dhdrr)."""
    ny = vm.shape[0]
    t = _zeros_tables(ny)
    G, A, B, C, D = t["G"], t["A"], t["B"], t["C"], t["D"]
    Vxx, Vxy, Vyy, Vxz, Vyz, Vzz = t["Vxx"], t["Vxy"], t["Vyy"], t["Vxz"], t["Vyz"], t["Vzz"]
    h, dhds, dhdr, dhdsr, dhdrr = hm
    rho, u1, u3, tm = vm[:, 0].copy(), vm[:, 1].copy(), vm[:, 3].copy(), vm[:, 4].copy()
    u2 = np.zeros(ny)  # spatial.f90:149
    z = np.zeros(ny)
    gam, gam1, Ma, Re, Pr = p.gamma, p.gamma1, p.Ma, p.Re, p.Pr
    gm2 = gam * Ma ** 2
    gum = np.zeros((ny, 3, 3))
    gum[:, 0, 1] = g2vm[:, 1]; gum[:, 1, 1] = g2vm[:, 2]; gum[:, 2, 1] = g2vm[:, 3]
    g11 = np.zeros((ny, 5)); g12 = g11; g13 = g11; g23 = g11; g33 = g11
    g22 = g22vm
    divum = (gum[:, 0, 0] + u2 * dhdr) / h + gum[:, 1, 1] + gum[:, 2, 2]
    grho = np.stack([z, g2vm[:, 0], z], axis=1)
    gt = np.stack([z, g2vm[:, 4], z], axis=1)
    fact = 1.0 / gm2
    gp = np.stack([fact * (grho[:, k] * tm + rho * gt[:, k]) for k in range(3)], axis=1)
    g1div = (-dhds / h ** 3 * (gum[:, 0, 0] + u2 * dhdr)
             + 1.0 / h ** 2 * (g11[:, 1] + gum[:, 1, 0] * dhdr + u2 * dhdsr)
             + 1.0 / h * (g12[:, 2] + g13[:, 3]))
    g2div = (-dhdr / h ** 2 * (gum[:, 0, 0] + u2 * dhdr)
             + 1.0 / h * (g12[:, 1] + gum[:, 1, 1] * dhdr + u2 * dhdrr)
             + (g22[:, 2] + g23[:, 3]))
    g3div = 1.0 / h * (g13[:, 1] + gum[:, 1, 2] * dhdr) + g23[:, 2] + g33[:, 3]
    S = np.zeros((ny, 3, 3))
    S[:, 0, 0] = (gum[:, 0, 0] + u2 * dhdr) / h
    S[:, 0, 1] = 0.5 * ((gum[:, 1, 0] - u1 * dhdr) / h + gum[:, 0, 1])
    S[:, 0, 2] = 0.5 * (gum[:, 2, 0] / h + gum[:, 0, 2])
    S[:, 1, 0] = S[:, 0, 1]
    S[:, 1, 1] = gum[:, 1, 1]
    S[:, 1, 2] = 0.5 * (gum[:, 2, 1] + gum[:, 1, 2])
    S[:, 2, 0] = S[:, 0, 2]
    S[:, 2, 1] = S[:, 1, 2]
    S[:, 2, 2] = gum[:, 2, 2]
    S1jj = (-0.5 * (dhdr ** 2 + dhdrr * h) / h ** 2 * u1 + 0.5 * dhdr * gum[:, 0, 1] / h + 0.5 * g22[:, 1]
            - dhds * gum[:, 0, 0] / h ** 3 + g11[:, 1] / h ** 2 + 0.5 * g33[:, 1]
            + (h * dhdsr - dhdr * dhds) / h ** 3 * u2 + 3.0 * dhdr * gum[:, 1, 0] / (2.0 * h ** 2)
            + 0.5 * g12[:, 2] / h + 0.5 * g13[:, 3] / h)
    S2jj = (0.5 * (dhdr * dhds - h * dhdsr) / h ** 3 * u1 - 3.0 * dhdr * gum[:, 0, 0] / (2.0 * h ** 2)
            + 0.5 * g12[:, 1] / h - dhdr ** 2 * u2 / h ** 2 + dhdr * gum[:, 1, 1] / h + g22[:, 2]
            - 0.5 * dhds * gum[:, 1, 0] / h ** 3 + 0.5 * g11[:, 2] / h ** 2 + 0.5 * g33[:, 2] + 0.5 * g23[:, 3])
    S3jj = (0.5 * g13[:, 1] / h + 0.5 * g23[:, 2] + 0.5 * dhdr * gum[:, 1, 2] / h + 0.5 * dhdr * gum[:, 2, 1] / h
            + 0.5 * g22[:, 3] + 0.5 * g11[:, 3] / h ** 2 + g33[:, 3] - 0.5 * dhds * gum[:, 2, 0] / h ** 3)
    LapT = 1.0 / h * (-dhds / h ** 2 * gt[:, 0] + 1.0 / h * g11[:, 4] + h * g22[:, 4]
                      + gt[:, 1] * dhdr + h * g33[:, 4])
    mu, dmu, d2mu, lm, dlm, d2lm, con, dcon, d2con = _material(tm, p)
    g1mu, g2mu, g3mu = (dmu * gt[:, k] for k in range(3))
    g1dmu, g2dmu, g3dmu = (d2mu * gt[:, k] for k in range(3))
    g1con, g2con, g3con = (dcon * gt[:, k] for k in range(3))
    g1dcon, g2dcon, g3dcon = (d2con * gt[:, k] for k in range(3))
    g1lm, g2lm, g3lm = (dlm * gt[:, k] for k in range(3))
    g1dlm, g2dlm, g3dlm = (d2lm * gt[:, k] for k in range(3))

    # continuity, spatial.f90:358-372
    G[:, 0, 0] = 1.0
    A[:, 0, 0] = u1 / h; A[:, 0, 1] = rho / h
    B[:, 0, 0] = u2; B[:, 0, 2] = rho
    C[:, 0, 0] = u3; C[:, 0, 3] = rho
    D[:, 0, 0] = divum; D[:, 0, 1] = grho[:, 0] / h
    D[:, 0, 2] = grho[:, 1] + rho * dhdr / h; D[:, 0, 3] = grho[:, 2]
    # x1 momentum, spatial.f90:376-450
    G[:, 1, 1] = rho
    A[:, 1, 0] = tm / (h * gm2); A[:, 1, 1] = rho * u1 / h; A[:, 1, 4] = rho / (h * gm2)
    B[:, 1, 1] = rho * u2
    C[:, 1, 1] = rho * u3
    D[:, 1, 0] = (u1 / h * (gum[:, 0, 0] + u2 * dhdr) + u2 * gum[:, 0, 1] + u3 * gum[:, 0, 2]
                  + gt[:, 0] / (h * gm2))
    D[:, 1, 1] = rho * (gum[:, 0, 0] + u2 * dhdr) / h
    D[:, 1, 2] = rho * (gum[:, 0, 1] + u1 * dhdr / h)
    D[:, 1, 3] = rho * gum[:, 0, 2]
    D[:, 1, 4] = grho[:, 0] / (h * gm2)
    if p.navier:
        fact = p.rlme / (p.rmue * Re)
        A[:, 1, 1] -= fact * (g1lm / h ** 2 - lm / h ** 3 * dhds)
        A[:, 1, 2] -= fact * lm / h ** 2 * dhdr
        A[:, 1, 4] -= fact * dlm * divum / h
        B[:, 1, 2] -= fact * (g1lm / h)
        C[:, 1, 3] -= fact * (g1lm / h)
        D[:, 1, 2] -= fact * (g1lm * dhdr / h ** 2 - lm / h ** 3 * dhds * dhdr + lm / h ** 2 * dhdsr)
        D[:, 1, 4] -= fact * (g1dlm * divum / h + dlm * g1div)
        Vxx[:, 1, 1] = fact * lm / h ** 2
        Vxy[:, 1, 2] = fact * lm / h
        Vxz[:, 1, 3] = fact * lm / h
        fact = 1.0 / Re
        A[:, 1, 1] -= fact * (2.0 * g1mu / h ** 2 - 2.0 * mu * dhds / h ** 3)
        A[:, 1, 2] -= fact * (g2mu / h + mu * 3.0 * dhdr / h ** 2)
        A[:, 1, 3] -= fact * g3mu / h
        A[:, 1, 4] -= fact * dmu * 2.0 * S[:, 0, 0] / h
        B[:, 1, 1] -= fact * (g2mu + mu * dhdr / h)
        B[:, 1, 4] -= fact * dmu * 2.0 * S[:, 0, 1]
        C[:, 1, 1] -= fact * g3mu
        C[:, 1, 4] -= fact * dmu * 2.0 * S[:, 0, 2]
        D[:, 1, 1] -= fact * (g2mu / h * (-dhdr) - mu * (dhdr ** 2 + dhdrr * h) / h ** 2)
        D[:, 1, 2] -= fact * (2.0 * g1mu / h ** 2 * dhdr + 2.0 * mu * (dhdsr * h - dhds * dhdr) / h ** 3)
        D[:, 1, 4] -= fact * 2.0 * (g1dmu / h * S[:, 0, 0] + g2dmu * S[:, 0, 1] + g3dmu * S[:, 0, 2]
                                    + dmu * S1jj)
        Vxx[:, 1, 1] += fact * 2.0 * mu / h ** 2
        Vxy[:, 1, 2] += fact * mu / h
        Vyy[:, 1, 1] += fact * mu
        Vxz[:, 1, 3] += fact * mu / h
        Vzz[:, 1, 1] += fact * mu
    # x2 momentum, spatial.f90:454-526
    G[:, 2, 2] = rho
    A[:, 2, 2] = rho * u1 / h
    B[:, 2, 0] = tm / gm2; B[:, 2, 2] = rho * u2; B[:, 2, 4] = rho / gm2
    C[:, 2, 2] = rho * u3
    D[:, 2, 0] = (u1 / h * (gum[:, 1, 0] - u1 * dhdr) + u2 * gum[:, 1, 1] + u3 * gum[:, 1, 2]
                  + gt[:, 1] / gm2)
    D[:, 2, 1] = rho * (gum[:, 1, 0] - 2.0 * u1 * dhdr) / h
    D[:, 2, 2] = rho * gum[:, 1, 1]
    D[:, 2, 3] = rho * gum[:, 1, 2]
    D[:, 2, 4] = grho[:, 1] / gm2
    if p.navier:
        fact = p.rlme / (p.rmue * Re)
        A[:, 2, 1] -= fact * (g2lm / h - lm * dhdr / h ** 2)
        B[:, 2, 2] -= fact * (g2lm + lm * dhdr / h)
        B[:, 2, 4] -= fact * dlm * divum
        C[:, 2, 3] -= fact * g2lm
        D[:, 2, 2] -= fact * (g2lm / h * dhdr - lm * dhdr / h ** 2 * dhdr + lm / h * dhdrr)
        D[:, 2, 4] -= fact * (g2dlm * divum + dlm * g2div)
        Vxy[:, 2, 1] = fact * lm / h
        Vyy[:, 2, 2] = fact * lm
        Vyz[:, 2, 3] = fact * lm
        fact = 1.0 / Re
        A[:, 2, 1] += fact * mu * 3.0 * dhdr / h ** 2
        A[:, 2, 2] -= fact * (g1mu / h ** 2 - mu * dhds / h ** 3)
        A[:, 2, 4] -= fact * dmu * 2.0 * S[:, 1, 0] / h
        B[:, 2, 1] -= fact * g1mu / h
        B[:, 2, 2] -= fact * (2.0 * g2mu + 2.0 * mu * dhdr / h)
        B[:, 2, 3] -= fact * g3mu
        B[:, 2, 4] -= fact * dmu * 2.0 * S[:, 1, 1]
        C[:, 2, 2] -= fact * g3mu
        C[:, 2, 4] -= fact * dmu * 2.0 * S[:, 1, 2]
        D[:, 2, 1] -= fact * (g1mu / h ** 2 * (-dhdr) + mu * (dhds * dhdr - h * dhdsr) / h ** 3)
        D[:, 2, 2] += fact * 2.0 * mu * dhdr ** 2 / h ** 2
        D[:, 2, 4] -= fact * 2.0 * (g1dmu / h * S[:, 1, 0] + g2dmu * S[:, 1, 1] + g3dmu * S[:, 1, 2]
                                    + dmu * S2jj)
        Vxx[:, 2, 2] += fact * mu / h ** 2
        Vxy[:, 2, 1] += fact * mu / h
        Vyy[:, 2, 2] += fact * 2.0 * mu
        Vyz[:, 2, 3] += fact * mu
        Vzz[:, 2, 2] += fact * mu
    # x3 momentum, spatial.f90:530-595
    G[:, 3, 3] = rho
    A[:, 3, 3] = rho * u1 / h
    B[:, 3, 3] = rho * u2
    C[:, 3, 0] = tm / gm2; C[:, 3, 3] = rho * u3; C[:, 3, 4] = rho / gm2
    D[:, 3, 0] = u1 * gum[:, 2, 0] / h + u2 * gum[:, 2, 1] + u3 * gum[:, 2, 2] + gt[:, 2] / gm2
    D[:, 3, 1] = rho * gum[:, 2, 0] / h
    D[:, 3, 2] = rho * gum[:, 2, 1]
    D[:, 3, 3] = rho * gum[:, 2, 2]
    D[:, 3, 4] = grho[:, 2] / gm2
    if p.navier:
        fact = p.rlme / (p.rmue * Re)
        A[:, 3, 1] -= fact * g3lm / h
        B[:, 3, 2] -= fact * g3lm
        C[:, 3, 3] -= fact * g3lm
        C[:, 3, 2] -= fact * lm / h * dhdr
        C[:, 3, 4] -= fact * dlm * divum
        D[:, 3, 2] -= fact * (g2lm / h * dhdr)
        D[:, 3, 4] -= fact * (g3dlm * divum + dlm * g3div)
        Vxz[:, 3, 1] = fact * lm / h
        Vyz[:, 3, 2] = fact * lm
        Vzz[:, 3, 3] = fact * lm
        fact = 1.0 / Re
        A[:, 3, 3] -= fact * (g1mu / h ** 2 - mu * dhds / h ** 3)
        A[:, 3, 4] -= fact * dmu * 2.0 * S[:, 2, 0] / h
        B[:, 3, 3] -= fact * (g2mu + mu * dhdr / h)
        B[:, 3, 4] -= fact * dmu * 2.0 * S[:, 2, 1]
        C[:, 3, 1] -= fact * g1mu / h
        C[:, 3, 2] -= fact * (g2mu + mu * dhdr / h)
        C[:, 3, 3] -= fact * 2.0 * g3mu
        C[:, 3, 4] -= fact * dmu * 2.0 * S[:, 2, 2]
        D[:, 3, 4] -= fact * 2.0 * (g1dmu / h * S[:, 2, 0] + g2dmu * S[:, 2, 1] + g3dmu * S[:, 2, 2]
                                    + dmu * S3jj)
        Vxx[:, 3, 3] += fact * mu / h ** 2
        Vyy[:, 3, 3] += fact * mu
        Vxz[:, 3, 1] += fact * mu / h
        Vyz[:, 3, 2] += fact * mu
        Vzz[:, 3, 3] += fact * 2.0 * mu
    # energy, spatial.f90:599-671
    G[:, 4, 4] = rho
    A[:, 4, 1] = rho * gam1 * tm / h; A[:, 4, 4] = rho * u1 / h
    B[:, 4, 2] = rho * gam1 * tm; B[:, 4, 4] = rho * u2
    C[:, 4, 3] = rho * gam1 * tm; C[:, 4, 4] = rho * u3
    D[:, 4, 0] = u1 / h * gt[:, 0] + u2 * gt[:, 1] + u3 * gt[:, 2] + gam1 * tm * divum
    D[:, 4, 1] = rho * gt[:, 0] / h
    D[:, 4, 2] = rho * gt[:, 1] + rho * gam1 * tm * dhdr / h
    D[:, 4, 3] = rho * gt[:, 2]
    D[:, 4, 4] = rho * gam1 * divum
    if p.navier:
        fact = gam / (Pr * Re)
        A[:, 4, 4] -= fact * (g1con / h ** 2 + dcon * gt[:, 0] / h ** 2 - con * dhds / h ** 3)
        B[:, 4, 4] -= fact * (g2con + dcon * gt[:, 1] + con * dhdr / h)
        C[:, 4, 4] -= fact * (g3con + dcon * gt[:, 2])
        D[:, 4, 4] -= fact * (g1dcon * gt[:, 0] / h ** 2 + g2dcon * gt[:, 1] + g3dcon * gt[:, 2]
                              + dcon * LapT)
        Vxx[:, 4, 4] = fact * con / h ** 2
        Vyy[:, 4, 4] = fact * con
        Vzz[:, 4, 4] = fact * con
        fact = gam * gam1 * Ma ** 2 * p.rlme / (Re * p.rmue)
        A[:, 4, 1] -= fact * 2.0 * lm * divum / h
        B[:, 4, 2] -= fact * 2.0 * lm * divum
        C[:, 4, 3] -= fact * 2.0 * lm * divum
        D[:, 4, 2] -= fact * 2.0 * lm * divum * dhdr / h
        D[:, 4, 4] -= fact * dlm * divum * divum
        fact = gam * gam1 * Ma ** 2 / Re
        for k in range(3):
            A[:, 4, 1 + k] -= fact * 4.0 * mu * S[:, k, 0] / h
            B[:, 4, 1 + k] -= fact * 4.0 * mu * S[:, k, 1]
            C[:, 4, 1 + k] -= fact * 4.0 * mu * S[:, k, 2]
        D[:, 4, 1] += fact * 4.0 * mu * S[:, 1, 0] * dhdr / h
        D[:, 4, 2] -= fact * 4.0 * mu * S[:, 0, 0] * dhdr / h
        D[:, 4, 4] -= fact * 2.0 * dmu * np.sum(S ** 2, axis=(1, 2))
    return t


def curvature_metrics(p: Params, y: np.ndarray):
    """spatial.f90:110-125.  Returns (x_out, (h, dhds, dhdr, dhdsr, dhdrr))."""
    ny = len(y)
    if p.curve == 2:
        s, *hm = circh(p.x, y)
        return s, tuple(hm)
    if p.curve == 1:
        raise NotImplementedError("calch needs NR RTFLSP: out of scope")
    z = np.zeros(ny)
    return p.x, (np.ones(ny), z.copy(), z.copy(), z.copy(), z.copy())


def assemble_spatial(p: Params, vm, deta, d2eta, hm, g2vm=None, g22vm=None):
    """C0, C1, C2 of spatial.f90:681-959 for p.omega, p.beta."""
    ny = p.ny
    n = NDOF * ny
    D1, D2, Dt1, Dt2 = deriv_ops(ny, p.wallt)
    if p.ider or g2vm is None:
        g2vm, g22vm = mean_gradients(vm, D1, D2, deta, d2eta)
    t = tables_spatial(vm, g2vm, g22vm, hm, p)
    be, om = p.beta, p.omega
    G = t["G"]
    Dh = t["D"] + IM * be * t["C"] + be ** 2 * t["Vzz"]
    Bh = t["B"] - IM * be * t["Vyz"]
    Bh = Bh * deta[:, None, None] - t["Vyy"] * d2eta[:, None, None]
    Vyy = t["Vyy"] * (deta ** 2)[:, None, None]
    C0 = np.zeros((ny, NDOF, ny, NDOF), dtype=np.complex128)
    C1 = np.zeros_like(C0); C2 = np.zeros_like(C0)

    def first_order_row(M, i, eq, Bh_, Dh_):
        M[i, eq, :, :] += Bh_[i, eq, None, :] * D1[i, :, None]
        M[i, eq, i, :] += Dh_[i, eq, :]

    w = ny - 1
    if p.top == 1:
        first_order_row(C0, 0, 0, Bh, Dh)
    for i in range(1, ny - 1):
        C0[i] += Bh[i][:, None, :] * D1[i][None, :, None] - Vyy[i][:, None, :] * D2[i][None, :, None]
        C0[i, :, i, :] += Dh[i]
    first_order_row(C0, w, 0, Bh, Dh)
    if p.wallt == 2:
        C0[w, 4, :, :4] += Bh[w, 4, None, :4] * D1[w, :, None] - Vyy[w, 4, None, :4] * D2[w, :, None]
        C0[w, 4, w, :4] += Dh[w, 4, :4]
        C0[w, 4, :, 4] += Bh[w, 4, 4] * Dt1[w, :] - Vyy[w, 4, 4] * Dt2[w, :]
        C0[w, 4, w, 4] += Dh[w, 4, 4]
    # time term + BCs, spatial.f90:792-835
    if p.top == 1:
        C0[0, 0, 0, 0] -= IM * om
    else:
        C0[0, 0, 0, 0] = -1.0
    for d in range(1, 5):
        C0[0, d, 0, d] = -1.0
    for i in range(1, ny - 1):
        C0[i, :, i, :] -= IM * om * G[i]
    C0[w, 0, w, 0] -= IM * om
    for d in (1, 2, 3):
        C0[w, d, w, d] -= 1.0
    if p.wallt == 0:
        C0[w, 4, w, 4] -= 1.0
    elif p.wallt == 2:
        C0[w, 4, w, :] -= IM * om * G[w, 4, :]
    else:
        raise ValueError("Illegal value of wallt")
    # C1, spatial.f90:839-938
    Dh1 = IM * t["A"] + be * t["Vxz"]
    Bh1 = (-IM * t["Vxy"]) * deta[:, None, None]
    if p.top == 1:
        first_order_row(C1, 0, 0, Bh1, Dh1)
    for i in range(1, ny - 1):
        C1[i] += Bh1[i][:, None, :] * D1[i][None, :, None]
        C1[i, :, i, :] += Dh1[i]
    first_order_row(C1, w, 0, Bh1, Dh1)
    if p.wallt == 2:
        C1[w, 4, :, :4] += Bh1[w, 4, None, :4] * D1[w, :, None]
        C1[w, 4, w, :4] += Dh1[w, 4, :4]
        C1[w, 4, :, 4] += Bh1[w, 4, 4] * Dt1[w, :]
        C1[w, 4, w, 4] += Dh1[w, 4, 4]
    # C2, spatial.f90:942-959
    for i in range(1, ny - 1):
        C2[i, :, i, :] = t["Vxx"][i]
    if p.wallt == 2:
        C2[w, 4, w, :] = t["Vxx"][w, 4, :]
    return C0.reshape(n, n), C1.reshape(n, n), C2.reshape(n, n), t


def companion_spatial(C0, C1, C2):
    """spatial.f90:978-1016: B0 = [[-C0^-1 C1, -C0^-1 C2], [I, 0]]."""
    from scipy.linalg import lapack
    n = C0.shape[0]
    lu, piv, info = lapack.zgetrf(np.asfortranarray(C0))
    M1, i1 = lapack.zgetrs(lu, piv, np.asfortranarray(-C1))
    M2, i2 = lapack.zgetrs(lu, piv, np.asfortranarray(-C2))
    B0 = np.zeros((2 * n, 2 * n), dtype=np.complex128, order="F")
    B0[:n, :n] = M1
    B0[:n, n:] = M2
    B0[n:, :n] = np.eye(n)
    return B0, info


def invert_spatial_eigs(lam: np.ndarray) -> np.ndarray:
    """spatial.f90:1065-1069: alpha = 1/lambda, 0 where lambda == 0."""
    out = np.zeros_like(lam)
    nz = lam != 0
    out[nz] = 1.0 / lam[nz]
    return out


def solve_spatial(p: Params, vm, deta, d2eta, hm, g2vm=None, g22vm=None, want_vectors=None,
                  as_coded=True):
    """spatial.f90:681-1084.  evec is NOT rescaled (spatial.f90:1100-1116 commented out)."""
    if want_vectors is None:
        want_vectors = p.ievec == 1
    C0, C1, C2, _ = assemble_spatial(p, vm, deta, d2eta, hm, g2vm, g22vm)
    B0, info = companion_spatial(C0, C1, C2)
    Bkeep = B0.copy()
    lam, vr, info = _zgeev(B0, want_vectors, as_coded)
    alp = invert_spatial_eigs(lam)
    order = stable_sort_by_imag(alp)
    out = dict(alp=alp[order], lam=lam[order], info=info, B0=Bkeep, C0=C0, C1=C1, C2=C2)
    if want_vectors:
        out["evec"] = vr[:, order]
    return out


# ----------------------------------------------------------------------------
# drivers: one deck -> one result (stab.f90:46-55 + temporal/spatial)
# ----------------------------------------------------------------------------
def prepare(p: Params, profile_text: str):
    """grid + mean flow (+ curvature metrics for spatial)."""
    y, eta, deta, d2eta = sgengrid(p.ny, p.yi, p.ymax)
    vm = getmean(read_profile(profile_text), y)
    return dict(y=y, eta=eta, deta=deta, d2eta=d2eta, vm=vm)


def run_deck(p: Params, profile_text: str, want_vectors=True, as_coded=True):
    g = prepare(p, profile_text)
    if p.itype == 1:
        res = solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=want_vectors,
                             as_coded=as_coded)
        res["x_out"] = p.x
    elif p.itype == 2:
        x_out, hm = curvature_metrics(p, g["y"])
        res = solve_spatial(p, g["vm"], g["deta"], g["d2eta"], hm, want_vectors=want_vectors,
                            as_coded=as_coded)
        res["x_out"] = x_out
        res["hm"] = hm
    else:
        raise NotImplementedError(p.itype)
    res.update(g)
    return res


# ----------------------------------------------------------------------------
# sweep enumeration (mtemporal.f90:25-39, mspatial.f90:68-96)
# ----------------------------------------------------------------------------
def nint(v: float) -> int:
    return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))


def mtemporal_points(amin, amax, ainc, bmin, bmax, binc):
    """Returns list of (iver, alpha, beta); upper end excluded (quirk q6)."""
    na = max(nint((amax - amin) / ainc), 1)
    nb = max(nint((bmax - bmin) / binc), 1)
    pts = []
    iver = 0
    for ia in range(1, na + 1):
        for ib in range(1, nb + 1):
            iver += 1
            pts.append((iver, amin + float(ia - 1) * ainc, bmin + float(ib - 1) * binc))
    return pts


def mspatial_points(omin, omax, oinc, bmin, bmax, binc):
    """Per station: list of (omega, beta); upper end included (mspatial.f90:68-73)."""
    if oinc == 0.0:
        oinc = 1.0
    if binc == 0.0:
        binc = 1.0
    no = nint((omax - omin) / oinc) + 1
    nb = nint((bmax - bmin) / binc) + 1
    return [(omin + float(io) * oinc, bmin + float(ib) * binc) for io in range(no) for ib in range(nb)]


def makename(base: str, iver: int) -> str:
    """mtemporal.f90:53-76."""
    if iver >= 10000:
        raise ValueError("Error in MakeName:  iver too large")
    return f"{base}.{iver}"


# ----------------------------------------------------------------------------
# gfortran sequential-unformatted records (temporal.f90:883-890, spatial.f90:1120-1126)
# ----------------------------------------------------------------------------
def _rec(payload: bytes) -> bytes:
    n = struct.pack("<i", len(payload))
    return n + payload + n


def write_eig_file(p: Params, res: dict, itype: int, with_vectors: bool) -> bytes:
    ny = p.ny
    buf = io.BytesIO()
    buf.write(_rec(struct.pack("<10i", p.ind, ny, NDOF, itype, p.ievec, p.curve, p.top, p.wall,
                               p.wallt, 1 if p.ider else 0)))
    buf.write(_rec(struct.pack("<9d", p.omega.real, p.omega.imag, p.alpha.real, p.alpha.imag,
                               p.beta.real, p.beta.imag, p.Re, p.Ma, p.Pr)))
    r3 = np.concatenate([[res["x_out"]], res["y"], res["eta"], res["deta"], res["d2eta"],
                         [p.yi, p.ymax]]).astype("<f8")
    buf.write(_rec(r3.tobytes()))
    vals = res["omg"] if itype == 1 else res["alp"]
    buf.write(_rec(np.asarray(vals, dtype="<c16").tobytes()))
    if with_vectors:
        buf.write(_rec(np.asfortranarray(res["evec"]).astype("<c16").tobytes(order="F")))
    return buf.getvalue()


def read_eig_file(data: bytes) -> dict:
    """Reader side of getevec.f90:77-91."""
    off = 0
    recs = []
    while off < len(data):
        (ln,) = struct.unpack_from("<i", data, off)
        recs.append(data[off + 4: off + 4 + ln])
        off += 8 + ln
    hdr = struct.unpack("<10i", recs[0])
    ind, ny, ndof, itype, ievec, curve, top, wall, wallt, ider = hdr
    r2 = struct.unpack("<9d", recs[1])
    r3 = np.frombuffer(recs[2], dtype="<f8")
    nmax = ndof * ny if itype == 1 else 2 * ndof * ny
    out = dict(ind=ind, ny=ny, ndof=ndof, itype=itype, ievec=ievec, curve=curve, top=top, wall=wall,
               wallt=wallt, ider=ider, omega=complex(r2[0], r2[1]), alpha=complex(r2[2], r2[3]),
               beta=complex(r2[4], r2[5]), Re=r2[6], Ma=r2[7], Pr=r2[8],
               x=r3[0], y=r3[1:1 + ny], eta=r3[1 + ny:1 + 2 * ny], deta=r3[1 + 2 * ny:1 + 3 * ny],
               d2eta=r3[1 + 3 * ny:1 + 4 * ny], yi=r3[1 + 4 * ny], ymax=r3[2 + 4 * ny],
               eval=np.frombuffer(recs[3], dtype="<c16"))
    if len(recs) > 4:
        out["evec"] = np.frombuffer(recs[4], dtype="<c16").reshape((nmax, nmax), order="F")
    return out


# ----------------------------------------------------------------------------
# getevec post-processing (getevec.f90:154-222)
# ----------------------------------------------------------------------------
def select_mode(evals: np.ndarray, value: complex) -> int:
    """getevec.f90:162-171: nearest eigenvalue, first one wins ties (0-based index)."""
    return int(np.argmin(np.abs(value - evals)))


def getevec_rows(y, evec_col, ny: int):
    """getevec.f90:179-222: rescale over the first ny*ndof rows by the REAL part of the
    max-|.| entry (quirk q7) and return rows wall -> freestream: [y, Re/Im x 5]."""
    v = np.array(evec_col[: ny * NDOF], dtype=np.complex128)
    scale = 0.0
    for i in range(ny * NDOF):
        if abs(v[i]) > abs(scale):
            scale = v[i].real
    if scale != 0.0:
        v = v / scale
    rows = np.empty((ny, 11))
    for r, i in enumerate(range(ny - 1, -1, -1)):
        rows[r, 0] = y[i]
        blk = v[i * NDOF:(i + 1) * NDOF]
        rows[r, 1::2] = blk.real
        rows[r, 2::2] = blk.imag
    return rows


def _fmt_e21(v: float) -> str:
    """Fortran 1pe21.13E3."""
    if v == 0.0:
        return " 0.0000000000000E+000" if not math.copysign(1, v) < 0 else "-0.0000000000000E+000"
    s = f"{v:.13E}"
    mant, exp = s.split("E")
    return f"{mant}E{exp[0]}{int(exp[1:]):03d}".rjust(21)


def getevec_text(p_like: dict, eigval: complex, rows: np.ndarray, itype: int) -> str:
    """Text of `time.N` / `space.N` (getevec.f90:196-222)."""
    def e13(v):
        s = f"{v:.6E}"; m, e = s.split("E"); return f"{m}E{e[0]}{int(e[1:]):02d}".rjust(13)
    out = [f"# Re = {e13(p_like['Re'])}, Ma = {e13(p_like['Ma'])}, Pr = {e13(p_like['Pr'])}"]
    def cl(tag, c):
        return f"# {tag} = ({_fmt_e21(c.real)},{_fmt_e21(c.imag)})"
    if itype == 1:
        out += [cl("Omega", eigval), cl("Alpha", p_like["alpha"]), cl("Beta ", p_like["beta"])]
    else:
        out += [cl("Omega", p_like["omega"]), cl("Alpha", eigval), cl("Beta ", p_like["beta"])]
    for r in rows:
        out.append("".join(_fmt_e21(v) + " " for v in r).rstrip())   # trailing 1x emits nothing (cf. thesis/TStest/time.ref)
    return "\n".join(out) + "\n"
