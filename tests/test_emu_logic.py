"""CPU logic tests of the kernel SOURCE (tests/emu: the .cuh files traced single-threaded on the
host) against the oracle.  These do not replace the GPU parity tests; they catch algorithmic
errors in a container without a GPU.  The emulation build is never part of the product."""
import ctypes as C

import numpy as np
import pytest
from scipy.linalg import lapack

import stab_oracle as so
from helpers import cptr, eigpair_residuals, emu, match_spectra, oracle_case, to_params


def _grid_bufs(p, g):
    vmc = np.ascontiguousarray(g["vm"].T)
    return vmc, np.ascontiguousarray(g["deta"]), np.ascontiguousarray(g["d2eta"])


def emu_temporal_matrix(p, g, apply_b0inv):
    q = to_params(p)
    n = 5 * p.ny
    vmc, de, d2e = _grid_bufs(p, g)
    M = np.empty((n, n), dtype=np.complex128)
    B0 = np.empty((n, n), dtype=np.complex128)
    al = np.array([p.alpha], dtype=np.complex128)
    be = np.array([p.beta], dtype=np.complex128)
    rc = emu().emu_temporal_matrix(C.byref(q), cptr(vmc), None, None, cptr(de), cptr(d2e), cptr(al), cptr(be),
                                   int(apply_b0inv), cptr(M), cptr(B0))
    assert rc == 0
    return M.T.copy(), B0.T.copy()


@pytest.mark.parametrize("over", [dict(ny=24), dict(ny=20, wallt=2), dict(ny=20, mattyp=1, T0=300.0, beta=0.2 + 0j),
                                  dict(ny=16, Re=0.0)])
def test_temporal_assembly_matches_oracle(over):
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", **over)
    A0r, B0r, _ = so.assemble_temporal(p, g["vm"], g["deta"], g["d2eta"])
    A0, B0 = emu_temporal_matrix(p, g, 0)
    scale = np.abs(A0r).max()
    assert np.abs(A0 - A0r).max() <= 1e-13 * scale
    assert np.array_equal(B0, B0r)
    # fused B0^-1: equals ZGESV's result to rounding
    M, _ = emu_temporal_matrix(p, g, 1)
    Mr = np.linalg.solve(B0r, A0r)
    assert np.abs(M - Mr).max() <= 1e-12 * np.abs(Mr).max()


def emu_spatial(p, g, companion=False):
    q = to_params(p)
    n = 5 * p.ny
    vmc, de, d2e = _grid_bufs(p, g)
    h5c = np.ascontiguousarray(g["h5"].T)
    Cs = [np.empty((n, n), dtype=np.complex128) for _ in range(3)]
    comp = np.empty((2 * n, 2 * n), dtype=np.complex128) if companion else None
    om = np.array([p.omega], dtype=np.complex128)
    be = np.array([p.beta], dtype=np.complex128)
    info = C.c_int(0)
    rc = emu().emu_spatial_matrices(C.byref(q), cptr(vmc), None, None, cptr(de), cptr(d2e), cptr(h5c), cptr(om), cptr(be),
                                    cptr(Cs[0]), cptr(Cs[1]), cptr(Cs[2]), None if comp is None else cptr(comp), C.byref(info))
    assert rc == 0
    return [c.T.copy() for c in Cs], (None if comp is None else comp.T.copy()), info.value


@pytest.mark.parametrize("deck,prof,over", [
    ("ts_spatial_ny32.inp", "ts_profile.0", dict(ny=20)),
    ("ts_spatial_ny32.inp", "ts_profile.0", dict(ny=16, top=1, wallt=2)),
    ("fsc_spatial_ny64.inp", "fsc_profile.0", dict(ny=20)),
    ("cf_spatial_ny96.inp", "cf_profile.0", dict(ny=18)),
])
def test_spatial_assembly_matches_oracle(deck, prof, over):
    p, g = oracle_case(deck, prof, **over)
    C0r, C1r, C2r, _ = so.assemble_spatial(p, g["vm"], g["deta"], g["d2eta"], g["hm"])
    (C0, C1, C2), comp, info = emu_spatial(p, g, companion=True)
    for a, b in ((C0, C0r), (C1, C1r), (C2, C2r)):
        assert np.abs(a - b).max() <= 1e-13 * max(np.abs(b).max(), 1.0)
    Br, _ = so.companion_spatial(C0r, C1r, C2r)
    assert info == 0
    assert np.abs(comp - Br).max() <= 1e-9 * np.abs(Br).max()


def _rand(n, seed):
    r = np.random.default_rng(seed)
    return r.standard_normal((n, n)) + 1j * r.standard_normal((n, n))


def emu_balance(A):
    n = A.shape[0]
    Ac = np.ascontiguousarray(A.T)
    scale = np.empty(n)
    ilo, ihi = C.c_int(0), C.c_int(0)
    emu().emu_balance(cptr(Ac), n, cptr(scale), C.byref(ilo), C.byref(ihi))
    return Ac.T.copy(), scale, ilo.value, ihi.value


def test_balance_matches_zgebal():
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=20)
    r = so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=False)
    for A in (r["M"], _rand(30, 1) * np.logspace(-6, 6, 30)[None, :]):
        ba, lo, hi, sc, info = lapack.zgebal(np.asfortranarray(A), scale=1, permute=1)
        bal, scale, ilo, ihi = emu_balance(A)
        assert (ilo, ihi) == (lo, hi)
        n = A.shape[0]
        ref_scale = sc.copy()
        # LAPACK stores 1-based permutation partners outside [ilo, ihi]
        for j in list(range(0, lo)) + list(range(hi + 1, n)):
            ref_scale[j] -= 1
        assert np.array_equal(scale, ref_scale)
        assert np.array_equal(bal, ba)


def emu_hessenberg(A, ilo, ihi):
    n = A.shape[0]
    Ac = np.ascontiguousarray(A.T)
    tau = np.empty(n, dtype=np.complex128)
    emu().emu_hessenberg(cptr(Ac), n, ilo, ihi, cptr(tau))
    return Ac.T.copy(), tau


def test_hessenberg_is_unitary_similarity():
    A = _rand(40, 2)
    Hh, tau = emu_hessenberg(A, 0, 39)
    H = np.triu(Hh, -1)
    # same layout as ZGEHRD: rebuild Q with ZUNGHR
    q, info = lapack.zunghr(np.asfortranarray(Hh), tau[:-1])
    assert info == 0
    assert np.abs(q.conj().T @ q - np.eye(40)).max() < 1e-13
    assert np.abs(q @ H @ q.conj().T - A).max() < 1e-12 * np.abs(A).max() * 40


def emu_hqr(H, ilo, ihi, W=24, ns=4, steps=16, nw=12, nibble=14):
    n = H.shape[0]
    Hc = np.ascontiguousarray(H.T)
    w = np.empty(n, dtype=np.complex128)
    info = emu().emu_hqr(cptr(Hc), n, ilo, ihi, cptr(w), W, ns, steps, nw, nibble)
    return w, info


@pytest.mark.parametrize("n,W,ns,steps,nw", [(12, 24, 4, 16, 12), (50, 24, 4, 16, 12), (90, 32, 6, 20, 16), (70, 20, 3, 7, 10),
                                             (150, 48, 16, 40, 24), (260, 96, 16, 64, 44), (100, 40, 16, 31, 20),
                                             (150, 48, 16, 40, 0), (180, 64, 16, 32, 32), (180, 64, 16, 32, 33), (220, 64, 16, 32, 45)])
def test_hqr_eigenvalues(n, W, ns, steps, nw):
    A = _rand(n, 3 + n)
    H = np.triu(A, -1)
    w, info = emu_hqr(H, 0, n - 1, W, ns, steps, nw)
    assert info == 0
    ref = np.linalg.eigvals(H)
    _, d = match_spectra(ref, w)
    assert d.max() < 1e-11 * np.abs(ref).max() * max(1.0, n / 50)


def emu_eig_pipeline(M, want_vectors=True, scale_rows=0):
    n = M.shape[0]
    bal, scale, ilo, ihi = emu_balance(M)
    Hh, tau = emu_hessenberg(bal, ilo, ihi)
    H = np.triu(Hh, -1)
    w, info = emu_hqr(H, ilo, ihi)
    V = None
    if want_vectors:
        Hc = np.ascontiguousarray(Hh.T)
        Vc = np.empty((n, n), dtype=np.complex128)
        hnorm = float(np.abs(H).sum(axis=1).max())
        lam = np.ascontiguousarray(w)
        bad = emu().emu_evec(cptr(Hc), n, ilo, ihi, cptr(tau), cptr(scale), cptr(lam), n, C.c_double(hnorm), scale_rows, cptr(Vc))
        assert bad == 0
        V = Vc.T.copy()
    return w, V, info


def test_full_pipeline_temporal_small():
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=20)
    r = so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=True)
    M, _ = emu_temporal_matrix(p, g, 1)
    w, V, info = emu_eig_pipeline(M)
    assert info == 0
    ref = r["omg"]
    perm, d = match_spectra(ref, w)
    scale = np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max())
    assert (d / scale).max() < 1e-8
    phys = np.abs(ref) < 2.0
    assert (d[phys] / np.maximum(np.abs(ref[phys]), 1e-2)).max() < 1e-9
    res = eigpair_residuals(r["M"], w, V)
    assert res.max() < 1e-12
    # eigenvectors: unit 2-norm, largest component real (ZGEEV's convention)
    assert np.abs(np.linalg.norm(V, axis=0) - 1).max() < 1e-12
    k = np.argmax(np.abs(V), axis=0)
    assert np.abs(V[k, np.arange(V.shape[1])].imag).max() == 0.0


def test_full_pipeline_random_vectors():
    A = _rand(48, 11)
    w, V, info = emu_eig_pipeline(A)
    assert info == 0
    assert eigpair_residuals(A, w, V).max() < 1e-13
    wr, vr = np.linalg.eig(A)
    perm, d = match_spectra(wr, w)
    assert d.max() < 1e-11 * np.abs(wr).max()
    # same vectors up to the normalisation both sides apply
    Vm = V[:, perm]
    dots = np.abs(np.sum(vr.conj() * Vm, axis=0))
    assert np.abs(dots - 1).max() < 1e-9


def test_full_pipeline_structured_matrices():
    """Balancing isolates eigenvalues (permutations), decoupled blocks: vectors from the leading
    block only (ZHSEIN's KR), ZGEBAK with ILO == IHI."""
    n = 40
    r = np.random.default_rng(7)
    T = np.triu(r.standard_normal((n, n)) + 1j * r.standard_normal((n, n)))
    P = np.eye(n)[r.permutation(n)]
    A1 = P @ T @ P.T
    A2 = r.standard_normal((n, n)) + 0j
    A2[5, :] = 0; A2[:, 9] = 0; A2[17, :] = 0
    for A in (A1, A2):
        w, V, info = emu_eig_pipeline(A)
        assert info == 0
        assert eigpair_residuals(A, w, V).max() < 1e-12
        _, d = match_spectra(np.linalg.eigvals(A), w)
        assert d.max() < 1e-9


def emu_hess_blocked(A, ilo, ihi):
    n = A.shape[0]
    Ac = np.ascontiguousarray(A.T)
    tau = np.empty(n, dtype=np.complex128)
    P = (n - 1 + 31) // 32
    T = np.zeros((P, 32, 32), dtype=np.complex128)
    emu().emu_hess_blocked(cptr(Ac), n, ilo, ihi, cptr(tau), cptr(T))
    return Ac.T.copy(), tau, T.transpose(0, 2, 1)


@pytest.mark.parametrize("n,ilo,ihi", [(40, 0, 39), (100, 0, 99), (97, 3, 90), (70, 0, 33), (33, 0, 32), (150, 5, 149)])
def test_hess_blocked_matches_zgehrd(n, ilo, ihi):
    """Same recurrences as LAPACK's blocked ZGEHRD -> H, the reflectors and tau agree to rounding."""
    A = _rand(n, 50 + n)
    A[ihi + 1:, :ihi + 1] = 0          # what ZGEBAL's permutation leaves: upper triangular outside [ilo, ihi]
    A[:, :ilo] = np.triu(A[:, :ilo])
    A[ihi + 1:, ihi + 1:] = np.triu(A[ihi + 1:, ihi + 1:])
    Hh, tau, T = emu_hess_blocked(A, ilo, ihi)
    ref, rtau, info = lapack.zgehrd(np.asfortranarray(A), lo=ilo, hi=ihi)
    assert info == 0
    q, info = lapack.zunghr(np.asfortranarray(Hh), tau[:-1], lo=ilo, hi=ihi)
    H = np.triu(Hh, -1)
    assert np.abs(q.conj().T @ q - np.eye(n)).max() < 1e-13
    assert np.abs(q @ H @ q.conj().T - A).max() < 1e-13 * n * np.abs(A).max()
    assert np.abs(Hh - ref).max() < 1e-11 * np.abs(A).max()
    assert np.abs(tau[:-1] - rtau).max() < 1e-11


