"""Host-side mirror of the reference's interface for the hot path.

The reference keeps its run state in module `stuff` and calls `temporal(name, ind)` /
`spatial(name, ind)` (temporal.f90:2, spatial.f90:2) once per point, from `stab.f90:46-92` or
from the sweep loops of `mtemporal.f90:25-39` / `mspatial.f90:68-96`.  Here module `stuff`
becomes a `Case` (deck + profile), the per-point subroutines keep their names and argument
meaning, and the sweep drivers hand the whole point list to ONE batched GPU call.  Results are
written in the reference's unformatted record format (`evec.dat` / `eig.<iver>`).
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional

import numpy as np

from . import binding as B


def _tokens(line: str):
    return line.split("!")[0].replace(",", " ").split()


@dataclasses.dataclass
class Case:
    """Run state of module stuff (stuff.f90:11-59) + grid + mean flow on the grid."""

    params: B.Params
    itype: int = 1
    alpha: complex = 0j
    beta: complex = 0j
    omega: complex = 0j
    ind: int = 0
    x: float = 0.0
    T0: float = 0.0
    tail: List[str] = dataclasses.field(default_factory=list)   # unread deck lines (sweep ranges)
    y: Optional[np.ndarray] = None
    eta: Optional[np.ndarray] = None
    deta: Optional[np.ndarray] = None
    d2eta: Optional[np.ndarray] = None
    vm: Optional[np.ndarray] = None
    h5: Optional[np.ndarray] = None
    g2vm: Optional[np.ndarray] = None     # analytic mean derivatives (ider=0, getmean2.f90), else None
    g22vm: Optional[np.ndarray] = None
    x_out: float = 0.0

    @property
    def ny(self) -> int:
        return self.params.ny

    def load_profile(self, path: str, first: Optional[str] = None, second: Optional[str] = None) -> "Case":
        """sgengrid (sgengrid.f90:15-45) + getmean (getmean.f90:27-111) [+ circh for curve=2].
        With ider=0 the reference calls getmean2 (getmean2.f90:26-187, temporal.f90:99-103), which reads
        `first.<ind>` / `second.<ind>` and treats them exactly like the profile (same spline, v := 0,
        constant beyond the table): pass their paths as `first`, `second`."""
        p = self.params
        self.y, self.eta, self.deta, self.d2eta = B.sgengrid(p.ny, p.yi, p.ymax)
        self.vm = B.getmean(B.read_profile(path), self.y)
        if p.ider == 0:
            if first is None or second is None:
                raise B.StabGpuError("ider=0 needs the first.<ind> and second.<ind> derivative tables (getmean2)")
            self.g2vm = B.getmean(B.read_profile(first), self.y)
            self.g22vm = B.getmean(B.read_profile(second), self.y)
        self.h5, self.x_out = None, self.x
        if self.itype in (2, 8):
            if p.curve == 2:                      # spatial.f90:112-118
                self.x_out, self.h5 = B.circh(self.x, self.y)
            elif p.curve == 1:
                raise B.StabGpuError("curve=1 (calch, parabolic cylinder) is outside the supported path")
        return self


def read_deck(text: str) -> Case:
    """Positional stdin deck: input.f90:15-122 then stab.f90:46-92 (itype 1, 2, 7, 8)."""
    lines = [ln for ln in text.splitlines() if ln.strip()]
    it = iter(lines)
    p = B.Params.default()
    p.mattyp = int(_tokens(next(it))[0])
    T0 = 0.0
    if p.mattyp == 1:
        T0 = float(_tokens(next(it))[0])
    t = _tokens(next(it)); p.Ma, p.Re, p.Pr = float(t[0]), float(t[1]), float(t[2])
    t = _tokens(next(it)); p.ny, p.yi, p.ymax = int(t[0]), float(t[1]), float(t[2])
    p.ievec = int(_tokens(next(it))[0])
    p.ider = 0 if int(_tokens(next(it))[0]) == 0 else 1
    t = _tokens(next(it)); p.top, p.wall, p.wallt, p.curve = (int(v) for v in t[:4])
    itype = int(_tokens(next(it))[0])
    B.edge_properties(p, T0)
    c = Case(params=p, itype=itype, T0=T0)
    if itype in (1, 3):
        t = _tokens(next(it)); c.alpha = complex(float(t[0]), float(t[1]))
        t = _tokens(next(it)); c.beta = complex(float(t[0]), float(t[1]))
    elif itype in (2, 4):
        t = _tokens(next(it)); c.omega = complex(float(t[0]), float(t[1]))
        t = _tokens(next(it)); c.beta = complex(float(t[0]), float(t[1]))
    if itype in (1, 2, 3, 4):
        c.ind = int(_tokens(next(it))[0])
        if itype in (2, 4):
            c.x = float(_tokens(next(it))[0])
            p.x = c.x
    c.tail = list(it)
    return c


def _write(case: Case, name: str, itype: int, omega, alpha, beta, eig, evec):
    B.write_eig_file(name, case.params, itype, case.ind, omega, alpha, beta, case.x_out, case.y, case.eta, case.deta,
                     case.d2eta, eig, evec)


def temporal(case: Case, name: Optional[str] = "evec.dat", want_vectors: bool = True):
    """temporal.f90:2 for the point (case.alpha, case.beta).  Returns dict(omg, evec, info)."""
    omg, ev, info = B.temporal_batch(case.params, case.vm, case.deta, case.d2eta, [case.alpha], [case.beta],
                                     g2vm=case.g2vm, g22vm=case.g22vm, want_vectors=want_vectors)
    if info[0] != 0:                                  # temporal.f90:776-785,806-809: stop
        raise B.StabGpuError(f"temporal: eigensolver failure, info = {int(info[0])}")
    res = dict(omg=omg[0], evec=None if ev is None else ev[0], info=int(info[0]))
    if name:
        _write(case, name, 1, case.omega, case.alpha, case.beta, res["omg"], res["evec"])
    return res


def spatial(case: Case, name: Optional[str] = "evec.dat", want_vectors: Optional[bool] = None):
    """spatial.f90:2 for the point (case.omega, case.beta); vectors only if ievec=1 (spatial.f90:1042-1048)."""
    if want_vectors is None:
        want_vectors = case.params.ievec == 1
    alp, ev, info = B.spatial_batch(case.params, case.vm, case.deta, case.d2eta, [case.omega], [case.beta], h5=case.h5,
                                    g2vm=case.g2vm, g22vm=case.g22vm, want_vectors=want_vectors)
    res = dict(alp=alp[0], evec=None if ev is None else ev[0], info=int(info[0]))
    if name:
        _write(case, name, 2, case.omega, case.alpha, case.beta, res["alp"], res["evec"])
    return res


def makename(base: str, iver: int) -> str:
    """mtemporal.f90:53-76."""
    if iver >= 10000:
        raise ValueError("Error in MakeName:  iver too large")
    return f"{base}.{iver}"


def mtemporal(case: Case, amin, amax, ainc, bmin, bmax, binc, outdir: Optional[str] = None, want_vectors: bool = True,
              rank: int = 0, world: int = 1):
    """mtemporal.f90:25-39: (alpha, beta) sweep, one `eig.<iver>` per point; the whole sweep is one
    batched call.  With world > 1 this rank solves the contiguous shard of stabgpu_shard_range."""
    a, b = B.mtemporal_points(amin, amax, ainc, bmin, bmax, binc)
    lo, hi = B.shard_range(a.size, rank, world)
    omg, ev, info = B.temporal_batch(case.params, case.vm, case.deta, case.d2eta, a[lo:hi] + 0j, b[lo:hi] + 0j,
                                     g2vm=case.g2vm, g22vm=case.g22vm, want_vectors=want_vectors)
    if outdir is not None:
        for k in range(hi - lo):
            _write(case, os.path.join(outdir, makename("eig", lo + k + 1)), 1, case.omega, complex(a[lo + k]),
                   complex(b[lo + k]), omg[k], None if ev is None else ev[k])
    return dict(alpha=a[lo:hi], beta=b[lo:hi], omg=omg, evec=ev, info=info, lo=lo, hi=hi)


def mspatial(case: Case, omin, omax, oinc, bmin, bmax, binc, outdir: Optional[str] = None,
             want_vectors: Optional[bool] = None, rank: int = 0, world: int = 1):
    """mspatial.f90:68-96 for one station (the profile in `case`)."""
    if want_vectors is None:
        want_vectors = case.params.ievec == 1
    o, b = B.mspatial_points(omin, omax, oinc, bmin, bmax, binc)
    lo, hi = B.shard_range(o.size, rank, world)
    alp, ev, info = B.spatial_batch(case.params, case.vm, case.deta, case.d2eta, o[lo:hi] + 0j, b[lo:hi] + 0j, h5=case.h5,
                                    g2vm=case.g2vm, g22vm=case.g22vm, want_vectors=want_vectors)
    if outdir is not None:
        for k in range(hi - lo):
            _write(case, os.path.join(outdir, makename("eig", lo + k + 1)), 2, complex(o[lo + k]), case.alpha,
                   complex(b[lo + k]), alp[k], None if ev is None else ev[k])
    return dict(omega=o[lo:hi], beta=b[lo:hi], alp=alp, evec=ev, info=info, lo=lo, hi=hi)


def read_delta(path: str):
    """delta.dat of mspatial.f90:31-66: rows `x_body delta`, lines starting with '#' skipped.  Returns (xb, delta)."""
    xb, delta = [], []
    with open(path) as fh:
        for line in fh:
            if not line.strip() or line.lstrip().startswith("#"):
                continue
            t = _tokens(line)
            xb.append(float(t[0]))
            delta.append(float(t[1]))
    return np.array(xb), np.array(delta)


def mspatial_stations(case: Case, omin, omax, oinc, bmin, bmax, binc, ind1: int, ind2: int, ind_inc: int, workdir: str = ".",
                      outdir: Optional[str] = None, want_vectors: Optional[bool] = None):
    """The full mspatial driver (mspatial.f90:20-96): for the stations ind1..ind2 (1-based rows of `delta.dat`), x = xb(ind),
    Yi = 2 delta(ind), mean flow from `profile.<ind>`; at every station the (omega, beta) grid in the reference's loop
    order; `iver` keeps counting across stations.  One batched GPU call per station (the grid and the mean flow change
    with the station)."""
    xb, delta = read_delta(os.path.join(workdir, "delta.dat"))
    if ind1 < 1 or ind2 > xb.size:
        raise B.StabGpuError("ERROR: illegal index...check delta.dat")          # mspatial.f90:46-50
    if ind_inc == 0:
        ind_inc = 1
    out, iver = [], 0
    for ind in range(ind1, ind2 + 1, ind_inc):
        st = dataclasses.replace(case, params=case.params.copy(), itype=2, ind=ind, x=float(xb[ind - 1]))
        st.params.yi = 2.0 * float(delta[ind - 1])
        st.params.x = st.x
        st.load_profile(os.path.join(workdir, f"profile.{ind}"))
        r = mspatial(st, omin, omax, oinc, bmin, bmax, binc, outdir=None, want_vectors=want_vectors)
        if outdir is not None:
            for k in range(r["omega"].size):
                _write(st, os.path.join(outdir, makename("eig", iver + k + 1)), 2, complex(r["omega"][k]), st.alpha,
                       complex(r["beta"][k]), r["alp"][k], None if r["evec"] is None else r["evec"][k])
        r.update(ind=ind, x=st.x, yi=st.params.yi, iver0=iver + 1)
        iver += r["omega"].size
        out.append(r)
    return out


def stab(deck_text: str, workdir: str = ".", name: str = "evec.dat"):
    """The main program's dispatch (stab.f90:40-92) for the Chebyshev path: itype 1 temporal, 2 spatial, 7 mtemporal,
    8 mspatial; mean flow from `profile.<ind>` (and `first.<ind>`, `second.<ind>` when ider=0) in `workdir`, results
    written there as `evec.dat` / `eig.<iver>`.  The finite-difference, Stokes and bump solvers (itype 3-6, 9) are outside
    the supported path."""
    case = read_deck(deck_text)

    def load(c: Case):
        f = [os.path.join(workdir, f"{b}.{c.ind}") for b in ("profile", "first", "second")]
        return c.load_profile(f[0], f[1], f[2]) if c.params.ider == 0 else c.load_profile(f[0])

    if case.itype == 1:
        return temporal(load(case), os.path.join(workdir, name))
    if case.itype == 2:
        return spatial(load(case), os.path.join(workdir, name))
    rows = [[float(v) for v in _tokens(ln)] for ln in case.tail if _tokens(ln)]
    if case.itype == 7:                                   # mtemporal(ind): `ind` is never read for this itype (stays 0)
        case.itype = 1
        (amin, amax, ainc), (bmin, bmax, binc) = rows[0][:3], rows[1][:3]
        return mtemporal(load(case), amin, amax, ainc, bmin, bmax, binc, outdir=workdir, want_vectors=case.params.ievec == 1)
    if case.itype == 8:
        case.itype = 2
        (omin, omax, oinc), (bmin, bmax, binc) = rows[0][:3], rows[1][:3]
        ind1, ind2, ind_inc = (int(v) for v in rows[2][:3])
        if len(rows) > 3 and int(rows[3][0]) != 0:
            raise B.StabGpuError("mspatial: dtype=1 (finite differences) is outside the supported path")
        return mspatial_stations(case, omin, omax, oinc, bmin, bmax, binc, ind1, ind2, ind_inc, workdir=workdir, outdir=workdir)
    raise B.StabGpuError(f"itype = {case.itype} is outside the supported path (1, 2, 7, 8)")

