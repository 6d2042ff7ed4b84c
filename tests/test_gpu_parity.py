"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI of
include/stabgpu.h, against the CPU oracle on the same inputs and against the reference's golden
vectors (tests/golden).  Tolerances (north star): eigenvalues 1e-10 relative on the physical
modes, eigenvectors 1e-8 after phase normalisation; element-wise operator entries ~1e-13."""
import io
import os

import numpy as np
import pytest

import stab_oracle as so
from conftest import golden_text
from helpers import eigpair_residuals, match_spectra, oracle_case, spectrum_parity, to_params

import stab_b200 as sb

pytestmark = pytest.mark.gpu


def _rows(name):
    return np.loadtxt(io.StringIO(golden_text(name)), comments="#")


def _rand(n, seed, batch=1):
    r = np.random.default_rng(seed)
    return r.standard_normal((batch, n, n)) + 1j * r.standard_normal((batch, n, n))


# ---- stage 1: assembly --------------------------------------------------------------------------
@pytest.mark.parametrize("over", [dict(ny=32), dict(ny=24, wallt=2), dict(ny=24, mattyp=1, T0=300.0, beta=0.2 + 0j),
                                  dict(ny=16, Re=0.0), dict(ny=128)])
def test_temporal_assembly(over):
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", **over)
    A0r, B0r, _ = so.assemble_temporal(p, g["vm"], g["deta"], g["d2eta"])
    A0, B0 = sb.temporal_assemble(to_params(p), g["vm"], g["deta"], g["d2eta"], p.alpha, p.beta)
    assert np.array_equal(B0, B0r)
    # entries are sums of products of O(N^4) derivative-matrix entries: compare row-scaled
    rs = np.abs(A0r).max(axis=1, keepdims=True) + 1e-300
    assert (np.abs(A0 - A0r) / rs).max() < 1e-11


@pytest.mark.parametrize("deck,prof,over", [
    ("ts_spatial_ny32.inp", "ts_profile.0", dict()),
    ("ts_spatial_ny32.inp", "ts_profile.0", dict(ny=24, top=1, wallt=2)),
    ("fsc_spatial_ny64.inp", "fsc_profile.0", dict()),
    ("cf_spatial_ny96.inp", "cf_profile.0", dict(ny=48)),
])
def test_spatial_assembly(deck, prof, over):
    p, g = oracle_case(deck, prof, **over)
    ref = so.assemble_spatial(p, g["vm"], g["deta"], g["d2eta"], g["hm"])[:3]
    got = sb.spatial_assemble(to_params(p), g["vm"], g["deta"], g["d2eta"], p.omega, p.beta, h5=g["h5"])
    for a, b in zip(got, ref):
        rs = np.abs(b).max(axis=1, keepdims=True) + 1e-300
        assert (np.abs(a - b) / np.maximum(rs, 1e-30)).max() < 1e-11


# ---- stage 3: balancing + Hessenberg ------------------------------------------------------------
def test_balance_bitwise_and_hessenberg_similarity():
    from scipy.linalg import lapack
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=32)
    M = so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=False)["M"]
    for A in (M, _rand(96, 5)[0] * np.logspace(-5, 5, 96)[None, :]):
        n = A.shape[0]
        ba, lo, hi, sc, info = lapack.zgebal(np.asfortranarray(A), scale=1, permute=1)
        bal, scale, ilo, ihi, Hh, tau = sb.debug_stages(A)
        assert (ilo, ihi) == (lo, hi)
        assert np.array_equal(bal, ba)
        q, info = lapack.zunghr(np.asfortranarray(Hh), tau[:-1], lo=ilo, hi=ihi)
        H = np.triu(Hh, -1)
        assert np.abs(q.conj().T @ q - np.eye(n)).max() < 1e-13
        assert np.abs(q @ H @ q.conj().T - ba).max() < 1e-13 * n * np.abs(ba).max()


@pytest.mark.parametrize("mode", [1, 5, 2, 0])
@pytest.mark.parametrize("n", [33, 97, 320, 640])
def test_hessenberg_matches_zgehrd(n, mode):
    """Stage 3b against LAPACK's ZGEHRD on the same balanced matrix: H, the reflectors and tau agree to
    rounding (mode 1: batched blocked + DMMA tensor-core updates, 2: same with scalar GEMM, 0: v1 unblocked)."""
    from scipy.linalg import lapack
    A = _rand(n, 900 + n)[0]
    A[5, :] = 0; A[:, 11] = 0                  # make ZGEBAL isolate something: ilo > 0 or ihi < n-1
    sb.set_hess_mode(mode)
    try:
        bal, scale, ilo, ihi, Hh, tau = sb.debug_stages(A)
    finally:
        sb.set_hess_mode(1)
    ba, lo, hi, sc, info = lapack.zgebal(np.asfortranarray(A), scale=1, permute=1)
    assert (ilo, ihi) == (lo, hi) and np.array_equal(bal, ba)
    ref, rtau, info = lapack.zgehrd(ba, lo=lo, hi=hi)
    scale_ = np.abs(ba).max()
    assert np.abs(np.triu(Hh, -1) - np.triu(ref, -1)).max() < 1e-11 * scale_ * np.sqrt(n)
    assert np.abs(Hh - ref).max() < 1e-10 * scale_ * np.sqrt(n)
    assert np.abs(tau[:-1] - rtau).max() < 1e-10


# ---- stages 3-6 on caller-supplied matrices (config Cr) ----------------------------------------
@pytest.mark.parametrize("n,batch", [(8, 3), (64, 4), (200, 3), (320, 2)])
def test_zgeev_batch_random(n, batch):
    A = _rand(n, 100 + n, batch)
    w, V, info = sb.zgeev_batch(A, want_vectors=True)
    assert np.all(info == 0)
    for b in range(batch):
        ref = np.linalg.eigvals(A[b])
        _, d = match_spectra(ref, w[b])
        assert d.max() < 1e-11 * np.abs(ref).max()
        assert eigpair_residuals(A[b], w[b], V[b]).max() < 1e-13
        assert np.abs(np.linalg.norm(V[b], axis=0) - 1).max() < 1e-12
    w2, _, info2 = sb.zgeev_batch(A, want_vectors=False)
    assert np.all(info2 == 0)
    for b in range(batch):
        _, d = match_spectra(w[b], w2[b])
        assert d.max() < 1e-11 * np.abs(w[b]).max()


@pytest.mark.parametrize("n,batch", [(1, 3), (2, 3), (3, 2), (5, 1), (31, 2), (33, 2), (65, 3), (129, 2), (257, 1), (513, 1), (639, 1), (16, 700)])
def test_zgeev_batch_ragged_orders(n, batch):
    """Orders around every tile / warp / panel boundary (32, 64, 128, 256, 512, 640), the trivial orders, and a batch larger
    than one wave of CTAs: eigenvalues against numpy, residuals at rounding level."""
    A = _rand(n, 900 + n, batch)
    w, V, info = sb.zgeev_batch(A, want_vectors=True)
    assert np.all(info == 0)
    for b in range(min(batch, 8)):
        ref = np.linalg.eigvals(A[b])
        _, d = match_spectra(ref, w[b])
        assert d.max() < 1e-11 * np.abs(ref).max()
        assert eigpair_residuals(A[b], w[b], V[b]).max() < max(1e-13, 4 * n * np.finfo(float).eps)
        assert np.abs(np.linalg.norm(V[b], axis=0) - 1).max() < 1e-12


def test_eigenvector_paths_agree():
    """Register-resident inverse iteration + tensor-core back-transformation (default) against the v1
    warp kernel (per-vector reflector application): same vectors to rounding."""
    A = _rand(200, 77, 2)
    w1, V1, i1 = sb.zgeev_batch(A, want_vectors=True)
    sb.set_evec_mode(0)
    try:
        w0, V0, i0 = sb.zgeev_batch(A, want_vectors=True)
    finally:
        sb.set_evec_mode(1)
    assert np.array_equal(w0, w1)
    assert np.abs(V0 - V1).max() < 1e-10


@pytest.mark.parametrize("n", [641, 700, 1024, 1280])
def test_zgeev_batch_two_warp_eigenvectors(n):
    """Orders 640 < n <= 1280 (spatial companion at Ny=128, temporal Ny=256): inverse iteration with two warps per
    eigenvalue (k_invit2) + tensor-core back-transformation; residuals, ZGEEV normalisation, and the v1 kernel's vectors."""
    A = _rand(n, 300 + n, 2)
    w, V, info = sb.zgeev_batch(A, want_vectors=True)
    assert np.all(info == 0)
    for b in range(2):
        ref = np.linalg.eigvals(A[b])
        _, d = match_spectra(ref, w[b])
        assert d.max() < 1e-11 * np.abs(ref).max()
        assert eigpair_residuals(A[b], w[b], V[b]).max() < 4 * n * np.finfo(float).eps      # backward stable: O(n eps)
        assert np.abs(np.linalg.norm(V[b], axis=0) - 1).max() < 1e-12
    if n in (641, 700):
        sb.set_evec_mode(0)
        try:
            w0, V0, i0 = sb.zgeev_batch(A[:1], want_vectors=True)
        finally:
            sb.set_evec_mode(1)
        assert np.array_equal(w0[0], w[0])
        assert np.abs(V0[0] - V[0]).max() < 1e-9


def test_spatial_companion_vectors_ny72():
    """Spatial problem with eigenvectors at companion order 2n = 720 (> 640: two-warp inverse iteration), block-triangular
    structure (many exactly-zero eigenvalues, leading-block orders kr < n): residuals of every finite mode against the
    oracle's companion matrix and the TS mode's eigenfunction against the oracle's."""
    p, g = oracle_case("ts_spatial_ny96.inp", "ts_profile.0", ny=72, ievec=1)
    alp, ev, info = sb.spatial_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], [p.omega], [p.beta], h5=g["h5"], want_vectors=True)
    assert info[0] == 0
    ref = so.solve_spatial(p, g["vm"], g["deta"], g["d2eta"], g["hm"], want_vectors=True)
    target = complex(2.2804739411367E-001, -6.5163146912049E-003)     # TStest/README.md:10 (Ny=96 value)
    j = so.select_mode(alp[0], target)
    jr = so.select_mode(ref["alp"], target)
    assert abs(alp[0][j] - ref["alp"][jr]) < 1e-10
    n = 5 * p.ny
    x, xr = ev[0][:n, j], ref["evec"][:n, jr]
    x, xr = x / x[np.argmax(np.abs(x))], xr / xr[np.argmax(np.abs(xr))]
    assert np.abs(x - xr).max() < 1e-8
    M = ref["B0"]
    fin = np.abs(alp[0]) > 1e-8
    lam = np.where(fin, 1.0 / np.where(fin, alp[0], 1.0), 0.0)

    def resid(vecs, lams):
        R = M @ vecs - vecs * lams[None, :]
        return np.linalg.norm(R, axis=0) / (np.linalg.norm(M) * np.linalg.norm(vecs, axis=0))
    res = resid(ev[0], lam)
    # Inverse iteration accepts a vector once it has grown by 0.1/sqrt(N) (ZLAEIN's criterion), which bounds the backward
    # error by ~10 N^1.5 ulp ||H||_inf: attained only by the near-defective continuous-branch modes at alpha ~ omega/c
    # (LAPACK's ZTREVC path gives ~5e-14 there); the discrete TS mode and the bulk of the spectrum sit at rounding level.
    N = 2 * n
    assert res[fin].max() < 10 * N ** 1.5 * np.finfo(float).eps * np.sqrt(N)
    assert res[j] < 1e-13
    assert np.median(res[fin]) < 1e-15


def test_zgeev_batch_structured():
    """Matrices with isolated eigenvalues (balancing permutes), defective blocks and zero rows."""
    n = 40
    r = np.random.default_rng(7)
    T = np.triu(r.standard_normal((n, n)) + 1j * r.standard_normal((n, n)))
    P = np.eye(n)[r.permutation(n)]
    A1 = P @ T @ P.T                       # fully permutable to triangular: ZGEBAL isolates everything
    A2 = r.standard_normal((n, n)) + 0j
    A2[5, :] = 0; A2[:, 9] = 0; A2[17, :] = 0     # zero rows/columns -> zero eigenvalues
    J = np.diag(np.full(n, 2.0 + 1j)) + np.diag(np.ones(n - 1), 1) * 1e-3   # nearly defective
    w, V, info = sb.zgeev_batch(np.stack([A1, A2, J]), want_vectors=True)
    assert np.all(info == 0)
    for A, wb, Vb in zip((A1, A2, J), w, V):
        ref = np.linalg.eigvals(A)
        _, d = match_spectra(ref, wb)
        assert d.max() < 1e-7 * max(np.abs(ref).max(), 1)
        assert eigpair_residuals(A, wb, Vb).max() < 1e-12
    assert np.sum(np.abs(w[1]) < 1e-12) >= 3


# ---- the hot path: temporal ---------------------------------------------------------------------
def _check_temporal_point(p, g, omg, ev, phys_tol=1e-10, vec_tol=1e-8):
    r = so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=ev is not None)
    ref = r["omg"]
    n = ref.size
    assert np.all(np.diff(omg.imag) >= 0)                      # sorted by Im (temporal.f90:844-855)
    # homogeneous Dirichlet rows give exactly-zero eigenvalues: 4 at the freestream, 4 at
    # the wall (3 when the wall energy equation is kept, wallt=2) -- same count as the oracle
    assert np.sum(omg == 0) == np.sum(ref == 0) >= (8 if p.wallt == 0 else 7)
    perm, d = match_spectra(ref, omg)
    phys = np.abs(ref) < 2.0
    # condition-aware parity over the WHOLE spectrum; 1e-10 relative wherever that is attainable
    diag = spectrum_parity(r["M"], ref, omg, rel_tol=phys_tol)
    assert diag["n_attainable"] >= 5
    # the least-stable discrete mode (what a stability analysis reads) to 1e-10 relative
    jm = np.argmax(np.where(phys, ref.imag, -np.inf))
    assert d[jm] / abs(ref[jm]) < phys_tol
    if ev is not None:
        res = eigpair_residuals(r["M"], omg, ev)
        assert res.max() < 1e-11
        # scaling of temporal.f90:867-879: the max-|.| entry is exactly 1
        k = np.argmax(np.abs(ev), axis=0)
        assert np.all(ev[k, np.arange(n)] == 1.0)
        # eigenvectors of well-separated physical modes agree with the oracle's
        sep = np.array([np.min(np.abs(np.delete(ref, j) - ref[j])) for j in range(n)])
        good = phys & (sep > 1e-3) & (ref != 0)
        assert good.sum() > 5
        dv = np.abs(ev[:, perm][:, good] - r["evec"][:, good]).max(axis=0)
        assert dv.max() < vec_tol
    return r


@pytest.mark.parametrize("over", [dict(ny=32), dict(ny=48, wallt=2), dict(ny=40, mattyp=1, T0=300.0, beta=0.15 + 0j)])
def test_temporal_point_vs_oracle(over):
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", **over)
    omg, ev, info = sb.temporal_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], [p.alpha], [p.beta], want_vectors=True)
    assert info[0] == 0
    _check_temporal_point(p, g, omg[0], ev[0])


def test_temporal_sweep_batch_vs_oracle():
    """mtemporal enumeration (config C2 at reduced Ny): every point of the batch matches the oracle."""
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=64)
    a, b = sb.mtemporal_points(0.05, 0.45, 0.4 / 8, 0.0, 0.1, 0.1)
    assert a.size == 8
    omg, ev, info = sb.temporal_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], a + 0j, b + 0j, want_vectors=False)
    assert np.all(info == 0)
    for k in range(a.size):
        p.alpha, p.beta = complex(a[k]), complex(b[k])
        _check_temporal_point(p, g, omg[k], None)


def test_temporal_golden_thesis_time_ref():
    # thesis/TStest/time.ref:2, README.md:25; the reference CI tolerance is abs 1e-8 (run.sh:36-39)
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0")
    omg, ev, info = sb.temporal_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], [p.alpha], [p.beta], want_vectors=True)
    assert info[0] == 0
    target = complex(1.1467880189410E-001, 2.3844535276599E-003)
    j = so.select_mode(omg[0], target)
    assert j == 479                                   # compbl/run.sh:13 selects sorted index 480
    assert abs(omg[0][j] - target) < 1e-10
    rows = so.getevec_rows(g["y"], ev[0][:, j], p.ny)
    assert np.abs(rows - _rows("ts_temporal_ny96.time.ref")).max() < 1e-8


def _cf_profile_text():
    from stab_b200 import fsc
    return fsc.format_table(fsc.profile_from_deck(golden_text("cf_thesis_fsc.inp"))["table"])


def test_temporal_golden_thesis_crossflow_time_ref():
    """thesis/CFtest (Collis thesis Ch. 4 crossflow vortex, M=0.3, Re=400, 45 deg sweep, beta_h=1) on the mean flow
    GENERATED by stab_b200/fsc.py: eigenvalue of time.ref:2 / run.sh:13 and the 96-row eigenfunction."""
    p, g = oracle_case("cf_thesis_temporal_ny96.inp", None, profile_text=_cf_profile_text())
    omg, ev, info = sb.temporal_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], [p.alpha], [p.beta], want_vectors=True)
    assert info[0] == 0
    target = complex(6.3418480187508E-007, 6.5335847258858E-003)
    j = so.select_mode(omg[0], target)
    assert abs(omg[0][j] - target) < 1e-9
    rows = so.getevec_rows(g["y"], ev[0][:, j], p.ny)
    assert np.abs(rows - _rows("cf_thesis_temporal_ny96.time.ref")).max() < 5e-8     # see tests/test_fsc.py on the 5e-8
    ref = so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=True)
    jr = so.select_mode(ref["omg"], target)
    assert abs(omg[0][j] - ref["omg"][jr]) < 1e-12
    assert np.abs(rows - so.getevec_rows(g["y"], ref["evec"][:, jr], p.ny)).max() < 1e-9


def test_temporal_crossflow_alpha_beta_grid_ny128():
    """BASELINE configs[2]: crossflow-vortex temporal sweep over an (alpha, beta) grid at Ny=128 on the
    generated Falkner-Skan-Cooke profile; mtemporal's enumeration; two points against the oracle, the unstable
    crossflow mode to 1e-10 relative."""
    p, g = oracle_case("cf_thesis_temporal_ny96.inp", None, profile_text=_cf_profile_text(), ny=128)
    a, b = sb.mtemporal_points(-0.5, 0.0, 0.125, 0.1, 0.6, 0.125)
    assert a.size == 4 * 4                       # both upper ends excluded (mtemporal.f90:25-39, quirk q6)
    omg, ev, info = sb.temporal_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], a + 0j, b + 0j, want_vectors=False)
    assert np.all(info == 0) and ev is None
    for k in (5, 13):
        p.alpha, p.beta = complex(a[k]), complex(b[k])
        _check_temporal_point(p, g, omg[k], None, phys_tol=1e-10)
    # the sweep contains amplified crossflow modes (Im omega > 0) around (alpha, beta) = (-0.25, 0.35)
    k = int(np.argmin(np.abs(a + 0.25) + np.abs(b - 0.35)))
    phys = np.abs(omg[k]) < 1.0
    assert omg[k][phys].imag.max() > 1e-3


def test_temporal_ny128_full_size_properties():
    """BASELINE config size (Ny=128, n=640): oracle comparison on 2 points + size-independent properties."""
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=128)
    al = np.array([0.15, 0.308620690, 0.4]) + 0j
    be = np.zeros(3) + 0j
    omg, ev, info = sb.temporal_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], al, be, want_vectors=True)
    assert np.all(info == 0)
    for k in (1,):
        p.alpha = complex(al[k])
        r = _check_temporal_point(p, g, omg[k], ev[k], phys_tol=1e-10)
    for k in range(3):
        p.alpha = complex(al[k])
        A0, B0, _ = so.assemble_temporal(p, g["vm"], g["deta"], g["d2eta"])
        # generalized residual on the ORIGINAL pencil and trace identity sum(omega) = tr(B0^-1 A0)
        R = A0 @ ev[k] - (B0 @ ev[k]) * omg[k][None, :]
        rel = np.linalg.norm(R, axis=0) / (np.linalg.norm(A0) * np.linalg.norm(ev[k], axis=0))
        assert rel.max() < 1e-11
        M = np.linalg.solve(B0, A0)
        assert abs(omg[k].sum() - np.trace(M)) < 1e-9 * np.abs(omg[k]).sum()


def test_bad_point_is_reported_and_does_not_poison_the_batch():
    """A NaN sweep value (or an Inf matrix entry) fails THAT point with a LAPACK-style info > 0 at once -- no iteration to
    the limit -- and leaves the other points of the batch bit-identical to a clean call (temporal.f90:806-809 stops on
    info /= 0; a batched caller needs the per-point status instead)."""
    import time
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=32)
    al = np.array([0.2, np.nan, 0.3]) + 0j
    t0 = time.perf_counter()
    omg, ev, info = sb.temporal_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], al, al * 0, want_vectors=True)
    assert time.perf_counter() - t0 < 10.0
    assert info[0] == 0 and info[2] == 0 and info[1] > 0
    ref, _, i2 = sb.temporal_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], al[[0, 2]], al[[0, 2]] * 0, want_vectors=False)
    assert np.array_equal(ref[0], omg[0]) and np.array_equal(ref[1], omg[2]) and not np.isnan(ev[0]).any()
    p2, g2 = oracle_case("ts_spatial_ny32.inp", "ts_profile.0", ny=24)
    om = np.array([0.08, np.nan, 0.1]) + 0j
    alp, _, info = sb.spatial_batch(to_params(p2), g2["vm"], g2["deta"], g2["d2eta"], om, om * 0, h5=g2["h5"])
    assert info[0] == 0 and info[2] == 0 and info[1] > 0 and not np.isnan(alp[[0, 2]]).any()
    A = _rand(40, 3, 3)
    A[1, 3, 4] = np.inf
    w, V, info = sb.zgeev_batch(A, want_vectors=True)
    assert info[0] == 0 and info[2] == 0 and info[1] > 0 and not np.isnan(w[[0, 2]]).any()


def test_temporal_re_ma_overrides():
    """Per-point Re/Ma overrides (neutral-curve sweeps, config C5) equal separate calls."""
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=32)
    q = to_params(p)
    Re = np.array([500.0, 1000.0, 2000.0])
    Ma = np.array([0.3, 0.3, 0.5])
    al = np.full(3, p.alpha)
    omg, _, info = sb.temporal_batch(q, g["vm"], g["deta"], g["d2eta"], al, np.zeros(3) + 0j, Re_pt=Re, Ma_pt=Ma)
    assert np.all(info == 0)
    for k in range(3):
        p.Re, p.Ma = float(Re[k]), float(Ma[k])
        ref = so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=False)["omg"]
        _, d = match_spectra(ref, omg[k])
        phys = np.abs(ref) < 2.0
        assert (d[phys] / np.maximum(np.abs(ref[phys]), 1e-3)).max() < 1e-10


# ---- the hot path: spatial ----------------------------------------------------------------------
def _check_spatial(p, g, alp, ev, target, rows_ref=None, rows_flip=False, tol_rows=1e-9):
    r = so.solve_spatial(p, g["vm"], g["deta"], g["d2eta"], g["hm"], want_vectors=False)
    assert np.all(np.diff(alp.imag) >= 0)
    j = so.select_mode(alp, target)
    assert abs(alp[j] - target) < 1e-10 * max(abs(target), 1)
    ref = r["alp"]
    # finite spectrum as multisets (lambda = 0 <-> alpha reported as 0 is rounding dependent, q8)
    fin = np.abs(ref) > 1e-8
    mine = alp[np.abs(alp) > 1e-8]
    win = fin & (np.abs(ref) < 5 * max(abs(target), 1))
    perm, d = match_spectra(ref[win], mine) if mine.size >= win.sum() else (None, None)
    if d is not None:
        assert np.median(d / np.abs(ref[win])) < 1e-9
    if rows_ref is not None:
        n = 5 * p.ny
        rows = so.getevec_rows(g["y"], ev[:, j], p.ny)
        if rows_flip:
            rows = rows[::-1]
        assert np.abs(rows - rows_ref).max() < tol_rows


def test_spatial_golden_ny32_space1():
    # test/space.1:3 (deck test/input.dat): eigenvalue and the 32-row eigenfunction
    p, g = oracle_case("ts_spatial_ny32.inp", "ts_profile.0")
    alp, ev, info = sb.spatial_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], [p.omega], [p.beta], h5=g["h5"], want_vectors=True)
    assert info[0] == 0
    _check_spatial(p, g, alp[0], ev[0], complex(2.2805022654496E-001, -6.5136925762007E-003), _rows("ts_spatial_ny32.space.ref"))


def test_spatial_golden_fsc_curve2():
    # FSCtest/space.ref:1 -- Streett map + circh metrics; rows freestream -> wall
    p, g = oracle_case("fsc_spatial_ny64.inp", "fsc_profile.0")
    alp, ev, info = sb.spatial_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], [p.omega], [p.beta], h5=g["h5"], want_vectors=True)
    assert info[0] == 0
    _check_spatial(p, g, alp[0], ev[0], complex(-4.6108596548503E-001, -7.0050272583092E-003),
                   _rows("fsc_spatial_ny64.space.ref"), rows_flip=True)


def test_spatial_golden_cf_ny64():
    # CFtest/README.md:12, CFtest/space.ref (M=0.8, Re=1e5, beta=35)
    p, g = oracle_case("cf_spatial_ny96.inp", "cf_profile.0", ny=64)
    alp, ev, info = sb.spatial_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], [p.omega], [p.beta], h5=g["h5"], want_vectors=True)
    assert info[0] == 0
    _check_spatial(p, g, alp[0], ev[0], complex(-3.7392297537875E+001, -2.9982641664422E-001),
                   _rows("cf_spatial_ny64.space.ref"), tol_rows=1e-8)


@pytest.mark.parametrize("deck,prof,over", [
    ("ts_spatial_ny32.inp", "ts_profile.0", dict()),                 # n = 160 = 5 panels
    ("ts_spatial_ny32.inp", "ts_profile.0", dict(ny=27)),            # n = 135: ragged last panel
    ("ts_spatial_ny32.inp", "ts_profile.0", dict(ny=5)),             # n = 25 < one panel
    ("cf_spatial_ny96.inp", "cf_profile.0", dict(ny=64)),            # M=0.8, Re=1e5, beta=35: row interchanges matter
    ("fsc_spatial_ny64.inp", "fsc_profile.0", dict(ny=128)),         # BASELINE size n = 640
])
def test_spatial_lu_reduce_blocked(deck, prof, over):
    """Stage 2 (ZGETRF + 2 x ZGETRS, spatial.f90:978-1008): the blocked DMMA LU reduce against LAPACK on the same C0, C1, C2
    and against the v1 one-CTA kernel; error bound = the backward-stable bound eps * cond(C0) relative to |M|."""
    import scipy.linalg as sla
    p, g = oracle_case(deck, prof, **over)
    P = to_params(p)
    C0, C1, C2 = sb.spatial_assemble(P, g["vm"], g["deta"], g["d2eta"], p.omega, p.beta, h5=g["h5"])
    n = C0.shape[0]
    lu, piv = sla.lu_factor(C0)
    ref = sla.lu_solve((lu, piv), np.hstack([-C1, -C2]))
    M, info = sb.debug_spatial_reduce(P, g["vm"], g["deta"], g["d2eta"], p.omega, p.beta, h5=g["h5"])
    assert info == 0 and M.shape == (n, 2 * n)
    sb.set_lu_mode(0)
    try:
        M0, info0 = sb.debug_spatial_reduce(P, g["vm"], g["deta"], g["d2eta"], p.omega, p.beta, h5=g["h5"])
    finally:
        sb.set_lu_mode(1)
    assert info0 == 0
    # residual of the defining equation, the quantity partial pivoting bounds: C0 M = -[C1 C2]
    rhs = np.hstack([-C1, -C2])
    scale = np.linalg.norm(C0, 1) * np.abs(ref).max() + np.abs(rhs).max()
    assert np.abs(C0 @ M - rhs).max() / scale < 1e-13
    assert np.abs(C0 @ M0 - rhs).max() / scale < 1e-13
    cond = np.linalg.cond(C0, 1)
    tol = 50 * n * np.finfo(float).eps * cond
    assert np.abs(M - ref).max() / np.abs(ref).max() < max(tol, 1e-13)
    assert np.abs(M - M0).max() / np.abs(ref).max() < max(tol, 1e-13)


def test_spatial_omega_sweep_readme_values():
    # TStest/README.md:9 (Ny=64); batch over omega, eigenvalues only (ievec=0)
    p, g = oracle_case("ts_spatial_ny96.inp", "ts_profile.0", ny=64, ievec=0)
    o, b = sb.mspatial_points(0.06, 0.10, 0.02, 0.0, 0.0, 0.0)
    alp, ev, info = sb.spatial_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], o + 0j, b + 0j, h5=g["h5"])
    assert np.all(info == 0) and ev is None
    target = complex(2.2804739410500E-001, -6.5163146761218E-003)
    j = so.select_mode(alp[1], target)
    assert abs(alp[1][j] - target) < 5e-11


def test_spatial_ny128_companion_full_size():
    """BASELINE configs[3] size: spatial companion problem at Ny=128 (order 2n = 1280), omega sweep,
    eigenvalues only; every finite mode in the physical window against the oracle."""
    p, g = oracle_case("ts_spatial_ny96.inp", "ts_profile.0", ny=128, ievec=0)
    om = np.array([0.06, 0.08, 0.10]) + 0j
    alp, ev, info = sb.spatial_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], om, om * 0, h5=g["h5"])
    assert np.all(info == 0)
    for k in (1,):
        p.omega = complex(om[k])
        ref = so.solve_spatial(p, g["vm"], g["deta"], g["d2eta"], g["hm"], want_vectors=False)["alp"]
        assert np.all(np.diff(alp[k].imag) >= 0)
        fin = np.abs(ref) > 1e-8
        mine = alp[k][np.abs(alp[k]) > 1e-8]
        assert abs(int(fin.sum()) - mine.size) <= 4                     # lambda = 0 <-> alpha := 0 is rounding dependent (q8)
        win = fin & (np.abs(ref) < 2.0)
        _, d = match_spectra(ref[win], mine)
        assert (d / np.abs(ref[win])).max() < 1e-9
        assert np.median(d / np.abs(ref[win])) < 1e-11
    # the TS mode of TStest/README.md:9-10 converges to the Ny=96 value within the author's Ny=64<->96 scatter
    target = complex(2.2804739411367E-001, -6.5163146912049E-003)
    j = so.select_mode(alp[1], target)
    assert abs(alp[1][j] - target) < 1e-9


def test_temporal_ny256_neutral_curve_points():
    """BASELINE configs[4] size (Ny=256, n=1280, eigenvalues only, per-point Re overrides): the least
    stable discrete mode and every well-conditioned mode against the oracle."""
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=256)
    al = np.array([0.25, 0.308620690]) + 0j
    Re = np.array([800.0, 1000.0])
    omg, _, info = sb.temporal_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], al, al * 0, Re_pt=Re)
    assert np.all(info == 0)
    p.alpha, p.Re = complex(al[1]), float(Re[1])
    ref = so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=False)["omg"]
    assert np.sum(omg[1] == 0) == np.sum(ref == 0) >= 8
    _, d = match_spectra(ref, omg[1])
    phys = np.abs(ref) < 2.0
    jm = np.argmax(np.where(phys, ref.imag, -np.inf))
    assert abs(ref[jm] - complex(1.1467880189e-01, 2.38445353e-03)) < 1e-8      # thesis/TStest/README.md:24-25
    assert d[jm] / abs(ref[jm]) < 1e-10
    # Ny=256 is beyond what 1e-10 on the full spectrum can mean: median and window checks
    assert np.median(d[phys] / np.maximum(np.abs(ref[phys]), 1e-3)) < 1e-10


def test_cli_harness_writes_reference_records(tmp_path):
    """host/stabgpu_cli: `stab < temporal.inp` on the GPU -- stdin deck + profile.<ind> in, evec.dat / eig.<iver> out."""
    import shutil, subprocess
    cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stab_b200", "stabgpu_cli")
    shutil.copy(os.path.join(os.path.dirname(__file__), "golden", "ts_profile.0"), tmp_path / "profile.0")
    deck = golden_text("ts_temporal_ny96.inp").replace("96", "40", 1)
    r = subprocess.run([cli], input=deck, capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    back = so.read_eig_file(open(tmp_path / "evec.dat", "rb").read())
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=40)
    ref = so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=True)
    assert back["ny"] == 40 and back["itype"] == 1 and back["alpha"] == p.alpha
    _, d = match_spectra(ref["omg"], back["eval"])
    phys = np.abs(ref["omg"]) < 2.0
    assert (d[phys] / np.maximum(np.abs(ref["omg"][phys]), 1e-3)).max() < 1e-10
    assert eigpair_residuals(ref["M"], back["eval"], back["evec"]).max() < 1e-11
    # itype 7: the (alpha, beta) sweep of mtemporal.f90 -> eig.1 .. eig.4.  The reference reads NO index line for this
    # itype (stab.f90:46-84 reads `ind` for itype 1-6 only; mtemporal(ind) runs with ind = 0): the same deck goes through
    # the Python mirror of stab.f90 and through the CLI, and the records must be byte-identical.
    lines = deck.splitlines()
    sweep = "\n".join(lines[:6] + ["7", "0.1 0.5 0.1", "0.0 0.0 1.0"]) + "\n"      # itype 7, alpha range, beta range
    for sub, ievec in (("cli0", "0"), ("cli1", "1")):
        d = tmp_path / sub
        d.mkdir()
        shutil.copy(tmp_path / "profile.0", d / "profile.0")
        dk = sweep.splitlines()
        dk[3] = ievec                                                    # deck line 4: ievec
        dk = "\n".join(dk) + "\n"
        r = subprocess.run([cli], input=dk, capture_output=True, text=True, cwd=d)
        assert r.returncode == 0, r.stderr + r.stdout
        assert sorted(f for f in os.listdir(d) if f.startswith("eig.")) == ["eig.1", "eig.2", "eig.3", "eig.4"]
        b3 = so.read_eig_file(open(d / "eig.3", "rb").read())
        assert abs(b3["alpha"] - 0.3) < 1e-15 and b3["ind"] == 0
        # the header's ievec and the presence of record 5 agree (getevec.f90:90 reads evec iff ievec == 1)
        assert ("evec" in b3) == (ievec == "1")
        py = tmp_path / (sub + "_py")
        py.mkdir()
        shutil.copy(tmp_path / "profile.0", py / "profile.0")
        sb.stab(dk, workdir=str(py))
        for k in (1, 2, 3, 4):
            assert (d / f"eig.{k}").read_bytes() == (py / f"eig.{k}").read_bytes()
    if ievec == "1":
        assert eigpair_residuals(so.solve_temporal(*_point(p, g, 0.3))["M"], b3["eval"], b3["evec"]).max() < 1e-11


def _point(p, g, alpha):
    import copy
    q = copy.copy(p)
    q.alpha = complex(alpha)
    return q, g["vm"], g["deta"], g["d2eta"]


def test_cli_itype8_and_ider0(tmp_path):
    """host/stabgpu_cli: itype 8 (mspatial.f90:20-96, station loop over delta.dat) and ider = 0 (getmean2: first.<ind> /
    second.<ind>) produce the same records as the Python mirror of stab.f90."""
    import shutil, subprocess
    cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stab_b200", "stabgpu_cli")
    prof = golden_text("ts_profile.0")
    for sub in ("cli", "py"):
        d = tmp_path / sub
        d.mkdir()
        for ind in (1, 2):
            (d / f"profile.{ind}").write_text(prof)
        (d / "delta.dat").write_text("# x_body  delta\n 10.0 0.5\n 20.0 0.45\n")
    sp = golden_text("ts_spatial_ny32.inp").splitlines()
    ms = "\n".join(sp[:6] + ["8", "0.06 0.08 0.02", "0.0 0.0 0.0", "1 2 1", "0"]) + "\n"
    r = subprocess.run([cli], input=ms, capture_output=True, text=True, cwd=tmp_path / "cli")
    assert r.returncode == 0, r.stderr + r.stdout
    sb.stab(ms, workdir=str(tmp_path / "py"))
    names = sorted(f for f in os.listdir(tmp_path / "cli") if f.startswith("eig."))
    assert names == ["eig.1", "eig.2", "eig.3", "eig.4"]
    for f in names:
        assert (tmp_path / "cli" / f).read_bytes() == (tmp_path / "py" / f).read_bytes()
    rec = so.read_eig_file((tmp_path / "cli" / "eig.4").read_bytes())
    assert rec["ind"] == 2 and rec["x"] == 20.0 and rec["yi"] == 0.9 and abs(rec["omega"] - 0.08) < 1e-15
    # ider = 0: derivative tables of the same profile next to it
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=24)
    tab = so.read_profile(prof)
    import scipy.interpolate as si
    d1 = np.column_stack([tab[:, 0]] + [si.CubicSpline(tab[:, 0], tab[:, k])(tab[:, 0], 1) for k in range(1, 6)])
    d2 = np.column_stack([tab[:, 0]] + [si.CubicSpline(tab[:, 0], tab[:, k])(tab[:, 0], 2) for k in range(1, 6)])
    deck = golden_text("ts_temporal_ny96.inp").splitlines()
    small = deck[:2] + ["24 1.0 0.0"] + deck[3:]
    small[4] = "0"                                                   # deck line 5: ider
    dk = "\n".join(small) + "\n"
    for sub in ("cli_i0", "py_i0"):
        d = tmp_path / sub
        d.mkdir()
        (d / "profile.0").write_text(prof)
        np.savetxt(d / "first.0", d1, fmt="%.16e")
        np.savetxt(d / "second.0", d2, fmt="%.16e")
    r = subprocess.run([cli], input=dk, capture_output=True, text=True, cwd=tmp_path / "cli_i0")
    assert r.returncode == 0, r.stderr + r.stdout
    sb.stab(dk, workdir=str(tmp_path / "py_i0"))
    assert (tmp_path / "cli_i0" / "evec.dat").read_bytes() == (tmp_path / "py_i0" / "evec.dat").read_bytes()


def test_mspatial_stations_delta_dat(tmp_path):
    """mspatial.f90:20-96: station loop over delta.dat (x = xb(ind), Yi = 2 delta(ind), profile.<ind>), the (omega, beta) grid
    at every station, `iver` counting across stations; each station equals a direct spatial call on its own grid."""
    prof = golden_text("ts_profile.0")
    for ind in (1, 2, 3):
        (tmp_path / f"profile.{ind}").write_text(prof)
    (tmp_path / "delta.dat").write_text("# x_body  delta\n 10.0 0.5\n# comment in the middle\n 20.0 0.45\n 30.0 0.6\n")
    case = sb.read_deck(golden_text("ts_spatial_ny32.inp"))
    out = sb.mspatial_stations(case, 0.06, 0.08, 0.02, 0.0, 0.0, 0.0, 2, 3, 1, workdir=str(tmp_path), outdir=str(tmp_path), want_vectors=False)
    assert [r["ind"] for r in out] == [2, 3] and [r["iver0"] for r in out] == [1, 3]
    assert [r["yi"] for r in out] == [0.9, 1.2] and [r["x"] for r in out] == [20.0, 30.0]
    assert sorted(f for f in os.listdir(tmp_path) if f.startswith("eig.")) == ["eig.1", "eig.2", "eig.3", "eig.4"]
    for r in out:
        assert np.all(r["info"] == 0) and r["omega"].tolist() == [0.06, 0.08]          # upper end included (mspatial.f90:72, q6)
        p, g = oracle_case("ts_spatial_ny32.inp", "ts_profile.0", yi=r["yi"], x=r["x"])
        p.omega = 0.08 + 0j
        ref = so.solve_spatial(p, g["vm"], g["deta"], g["d2eta"], g["hm"], want_vectors=False)["alp"]
        fin = (np.abs(ref) > 1e-8) & (np.abs(ref) < 2.0)
        _, d = match_spectra(ref[fin], r["alp"][1])
        assert np.median(d / np.abs(ref[fin])) < 1e-10
    recs = so.read_eig_file((tmp_path / "eig.4").read_bytes())
    assert abs(recs["omega"] - 0.08) < 1e-15 and recs["ind"] == 3 and recs["x"] == 30.0 and recs["yi"] == 1.2
    assert np.array_equal(recs["eval"], out[1]["alp"][1])


def test_stab_main_dispatch(tmp_path):
    """stab.f90:40-92 through the host mirror: itype 1 writes evec.dat, itype 7 one eig.<iver> per (alpha, beta) point,
    itype 8 runs the station loop, anything else is refused."""
    (tmp_path / "profile.0").write_text(golden_text("ts_profile.0"))
    (tmp_path / "profile.1").write_text(golden_text("ts_profile.0"))
    (tmp_path / "delta.dat").write_text(" 0.0 0.5\n")
    deck = golden_text("ts_temporal_ny96.inp").splitlines()         # positional deck: line 3 = Ny Yi Ymax, line 7 = itype
    small = deck[:2] + ["24 1.0 0.0"] + deck[3:]
    r = sb.stab("\n".join(small), workdir=str(tmp_path))
    rec = so.read_eig_file((tmp_path / "evec.dat").read_bytes())
    assert r["info"] == 0 and rec["itype"] == 1 and rec["ny"] == 24 and np.array_equal(rec["eval"], r["omg"])
    sweep = small[:6] + ["7", "0.25 0.35 0.05", "0.0 0.1 1.0"]
    out = sb.stab("\n".join(sweep), workdir=str(tmp_path))
    assert np.allclose(out["alpha"], [0.25, 0.30]) and np.all(out["info"] == 0)
    assert abs(so.read_eig_file((tmp_path / "eig.2").read_bytes())["alpha"] - 0.30) < 1e-15
    sp = golden_text("ts_spatial_ny32.inp").splitlines()
    ms = sp[:6] + ["8", "0.08 0.08 0.0", "0.0 0.0 0.0", "1 1 1", "0"]
    st = sb.stab("\n".join(ms), workdir=str(tmp_path))
    assert len(st) == 1 and st[0]["yi"] == 1.0 and np.all(st[0]["info"] == 0)
    target = complex(2.2805022654496E-001, -6.5136925762007E-003)              # test/space.1:3: Yi = 2 * 0.5 = 1 is the deck's own grid
    assert np.abs(st[0]["alp"][0] - target).min() < 1e-10
    with pytest.raises(sb.StabGpuError):
        sb.stab("\n".join(small[:6] + ["5", "0"]), workdir=str(tmp_path))


def test_ider0_analytic_derivatives_getmean2(tmp_path):
    """ider=0 (getmean2.f90:26-187, temporal.f90:99-103): the mean derivatives come from first.<ind> /
    second.<ind> tables instead of D1/D2.  Tables here are finite differences of the shipped profile."""
    prof = np.loadtxt(io.StringIO(golden_text("ts_profile.0")), comments="#")
    yv = prof[:, 0]
    first = np.column_stack([yv] + [np.gradient(prof[:, k], yv) for k in range(1, 6)])
    second = np.column_stack([yv] + [np.gradient(first[:, k], yv) for k in range(1, 6)])
    fmt = "%.15e"
    np.savetxt(tmp_path / "first.0", first, fmt=fmt); np.savetxt(tmp_path / "second.0", second, fmt=fmt)
    c = sb.read_deck(golden_text("ts_temporal_ny96.inp"))
    c.params.ny = 32; c.params.ider = 0
    c.load_profile(os.path.join(os.path.dirname(__file__), "golden", "ts_profile.0"), str(tmp_path / "first.0"), str(tmp_path / "second.0"))
    # oracle: same tables through its getmean (the derivative files are treated exactly like the profile)
    p, gg = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=32, ider=0)
    g2 = so.getmean(first, gg["y"]); g22 = so.getmean(second, gg["y"])
    assert np.abs(c.g2vm - g2).max() <= 1e-13 * np.abs(g2).max() and np.abs(c.g22vm - g22).max() <= 1e-13 * np.abs(g22).max()
    A0r, B0r, _ = so.assemble_temporal(p, gg["vm"], gg["deta"], gg["d2eta"], g2, g22)
    A0, B0 = sb.temporal_assemble(c.params, c.vm, c.deta, c.d2eta, c.alpha, c.beta, g2vm=c.g2vm, g22vm=c.g22vm)
    rs = np.abs(A0r).max(axis=1, keepdims=True) + 1e-300
    assert (np.abs(A0 - A0r) / rs).max() < 1e-11 and np.array_equal(B0, B0r)
    res = sb.temporal(c, None, want_vectors=False)
    ref = so.solve_temporal(p, gg["vm"], gg["deta"], gg["d2eta"], g2, g22, want_vectors=False)["omg"]
    _, d = match_spectra(ref, res["omg"])
    phys = np.abs(ref) < 2.0
    assert (d[phys] / np.maximum(np.abs(ref[phys]), 1e-3)).max() < 1e-9


def test_temporal_polish_shift_invert():
    """Stage 4 of the north star (new functionality, no reference code): shift-invert inverse iteration on the
    pencil polishes a mode to the oracle's eigenpair from a 0.1 % perturbed shift."""
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=64)
    r = so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=True)
    phys = np.abs(r["omg"]) < 2.0
    j = int(np.argmax(np.where(phys, r["omg"].imag, -np.inf)))
    target = r["omg"][j]
    lam, x, resid, iters = sb.temporal_polish(to_params(p), g["vm"], g["deta"], g["d2eta"], p.alpha, p.beta, target * (1 + 1e-3))
    assert 0 < iters <= 8 and resid < 1e-10        # floor ~ eps * cond(A0 - sigma B0) at this Ny
    assert abs(lam - target) < 1e-10 * abs(target)
    assert np.abs(x - r["evec"][:, j]).max() < 1e-8
    R = r["A0"] @ x - lam * (r["B0"] @ x)
    assert np.linalg.norm(R) / (np.linalg.norm(r["A0"]) * np.linalg.norm(x)) < 1e-12


# ---- reference-facing API: files ----------------------------------------------------------------
def test_api_temporal_writes_reference_format(tmp_path):
    c = sb.read_deck(golden_text("ts_temporal_ny96.inp"))
    c.params.ny = 24
    c.load_profile(os.path.join(os.path.dirname(__file__), "golden", "ts_profile.0"))
    name = str(tmp_path / "evec.dat")
    res = sb.temporal(c, name)
    back = so.read_eig_file(open(name, "rb").read())
    assert back["ny"] == 24 and back["itype"] == 1
    assert np.array_equal(back["eval"], res["omg"])
    assert np.array_equal(back["evec"], res["evec"])
    out = sb.mtemporal(c, 0.1, 0.3, 0.1, 0.0, 0.0, 1.0, outdir=str(tmp_path), want_vectors=False)
    assert sorted(f for f in os.listdir(tmp_path) if f.startswith("eig.")) == ["eig.1", "eig.2"]
    b2 = so.read_eig_file(open(str(tmp_path / "eig.2"), "rb").read())
    assert np.array_equal(b2["eval"], out["omg"][1]) and b2["alpha"] == complex(out["alpha"][1])
