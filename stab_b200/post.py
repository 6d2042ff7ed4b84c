"""Post-processing on in-memory batched results: the selection / rescaling / text format of
`getevec` (getevec.f90:154-227) and the mode trackers of `getalpha` / `getab` (getalpha.f90:137-149,
getab.f90:188-206), so that a sweep turns into growth-rate curves without writing 6.5-26 MB per point,
and the reference's own CI check (`ndiff -abserr 1e-8 time.1 time.ref`) can be run on GPU results.
Host-side only (numpy); nothing here touches the device."""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

NDOF = 5


def fmt_e21(v: float) -> str:
    """Fortran edit descriptor 1pe21.13E3."""
    if v == 0.0 or not math.isfinite(v):
        sign = "-" if math.copysign(1.0, v) < 0 else " "
        return f"{sign}0.0000000000000E+000" if v == 0.0 else f"{v:21}"
    mant, exp = f"{v:.13E}".split("E")
    return f"{mant}E{'+' if int(exp) >= 0 else '-'}{abs(int(exp)):03d}".rjust(21)


def fmt_e13(v: float) -> str:
    """Fortran 1pe13.6."""
    return f"{v:13.6E}"


def select_mode(evals: np.ndarray, value: Optional[complex] = None, index: Optional[int] = None) -> int:
    """getevec.f90:154-174: by 1-based index (-i) or nearest eigenvalue, first one wins ties (-v).  Returns 0-based."""
    if index is not None:
        if index < 1 or index > len(evals):
            raise ValueError("illegal eigenfunction index")
        return index - 1
    d = np.abs(value - np.asarray(evals))
    return int(np.argmin(d))


def rescale(evec_col: np.ndarray, ny: int) -> np.ndarray:
    """getevec.f90:179-189: divide the first ny*ndof rows by the max-|.| entry -- `scale` is declared REAL there,
    so only its real part survives (quirk q7)."""
    v = np.array(evec_col[: ny * NDOF], dtype=np.complex128)
    scale = 0.0
    for i in range(ny * NDOF):
        if abs(v[i]) > abs(scale):
            scale = v[i].real
    return v / scale if scale != 0.0 else v


def getevec_text(itype: int, Re: float, Ma: float, Pr: float, omega: complex, alpha: complex, beta: complex, y: np.ndarray,
                 evals: np.ndarray, evec: np.ndarray, value: Optional[complex] = None, index: Optional[int] = None) -> str:
    """The `time.N` / `space.N` file of getevec.f90:193-222 for one selected mode."""
    ny = len(y)
    j = select_mode(evals, value, index)
    v = rescale(evec[:, j], ny)
    cl = lambda z: f"({fmt_e21(z.real)},{fmt_e21(z.imag)})"
    lines = [f"# Re = {fmt_e13(Re)}, Ma = {fmt_e13(Ma)}, Pr = {fmt_e13(Pr)}"]
    if itype in (1, 3):
        lines += [f"# Omega = {cl(complex(evals[j]))}", f"# Alpha = {cl(complex(alpha))}", f"# Beta  = {cl(complex(beta))}"]
    else:
        lines += [f"# Omega = {cl(complex(omega))}", f"# Alpha = {cl(complex(evals[j]))}", f"# Beta  = {cl(complex(beta))}"]
    for i in range(ny - 1, -1, -1):                      # wall -> freestream
        blk = v[i * NDOF:(i + 1) * NDOF]
        vals = [y[i]] + [c for z in blk for c in (z.real, z.imag)]
        lines.append("".join(fmt_e21(x) + " " for x in vals).rstrip())
    return "\n".join(lines) + "\n"


def track_nearest(spectra: Sequence[np.ndarray], start: complex) -> np.ndarray:
    """getalpha.f90:137-149: follow one mode through a sweep by the eigenvalue nearest to the previous one."""
    out = np.empty(len(spectra), dtype=np.complex128)
    prev = start
    for k, ev in enumerate(spectra):
        j = int(np.argmin(np.abs(np.asarray(ev) - prev)))
        out[k] = prev = ev[j]
    return out


def track_extrapolated(spectra: Sequence[np.ndarray], params: Sequence[float], start: complex) -> np.ndarray:
    """getab.f90:188-206: as above, but the guess for point k is the linear extrapolation of the last two picks."""
    out = np.empty(len(spectra), dtype=np.complex128)
    e1 = e2 = None
    p1 = p2 = None
    for k, (ev, prm) in enumerate(zip(spectra, params)):
        if e1 is None:
            guess = start
        elif e2 is None or p2 == p1:
            guess = e1
        else:
            guess = e1 + (e2 - e1) / (p2 - p1) * (prm - p1)
        j = int(np.argmin(np.abs(np.asarray(ev) - guess)))
        e2, p2 = e1, p1
        e1, p1 = ev[j], prm
        out[k] = ev[j]
    return out
