"""CPU tests: the C-ABI library loads and exports every symbol of include/stabgpu.h; the host-side
pieces (grid, spline, Chebyshev matrix, curvature metrics, sweep enumeration, record writer)
agree with the oracle; compute entry points fail loudly without a GPU."""
import os
import re

import numpy as np
import pytest

import stab_oracle as so
from conftest import GOLDEN, ROOT, golden_text
from helpers import oracle_case, to_params

import stab_b200 as sb


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "stabgpu.h")).read()
    names = sorted(set(re.findall(r"\b(stabgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    L = sb.lib()
    for nm in names:
        assert hasattr(L, nm), nm


def test_params_struct_layout_matches_header():
    import ctypes as C
    # 8 ints + 16 doubles, no padding surprises
    assert C.sizeof(sb.Params) == 8 * 4 + 16 * 8
    p = sb.Params.default()
    assert (p.gamma, p.gamma1, p.cp) == (1.4, 0.4, 1003.1)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=8)
    with pytest.raises(sb.StabGpuError, match="no CUDA device"):
        sb.temporal_batch(to_params(p), g["vm"], g["deta"], g["d2eta"], [p.alpha], [p.beta])
    with pytest.raises(sb.StabGpuError):
        sb.zgeev_batch(np.eye(4, dtype=complex))


@pytest.mark.parametrize("ny,yi,ymax", [(32, 1.0, 0.0), (64, 0.05, 26.028), (96, 0.01, 0.1)])
def test_sgengrid(ny, yi, ymax):
    got = sb.sgengrid(ny, yi, ymax)
    ref = so.sgengrid(ny, yi, ymax)
    for a, b in zip(got, ref):
        assert np.allclose(a, b, rtol=1e-15, atol=0)   # x**3: pow() in numpy vs x*x*x here (gfortran expands integer powers)


@pytest.mark.parametrize("N", [15, 63, 127])
def test_chebyd_bitwise(N):
    assert np.array_equal(sb.chebyd(N), so.chebyd(N))


@pytest.mark.parametrize("prof", ["ts_profile.0", "cf_profile.0", "fsc_profile.0"])
def test_read_profile_and_getmean(prof):
    tab = sb.read_profile(os.path.join(GOLDEN, prof))
    ref = so.read_profile(golden_text(prof))
    assert np.array_equal(tab, ref)
    y = so.sgengrid(48, 1.0, 0.0)[0] if prof == "ts_profile.0" else so.sgengrid(48, 0.05, tab[-1, 0] * 1.2)[0]
    vm = sb.getmean(tab, y)
    vr = so.getmean(ref, y)
    assert np.abs(vm - vr).max() <= 1e-14 * np.abs(vr).max()
    assert np.all(vm[:, 2] == 0.0)          # v := 0 (getmean.f90:75)


@pytest.mark.parametrize("wallt", [0, 2])
def test_mean_gradients(wallt):
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=40, wallt=wallt)
    D1, D2, Dt2w, g2, g22 = sb.mean_gradients(g["vm"], g["deta"], g["d2eta"], wallt)
    rD1, rD2, rDt1, rDt2 = so.deriv_ops(p.ny, wallt)
    assert np.array_equal(D1, rD1)
    assert np.abs(D2 - rD2).max() <= 1e-12 * np.abs(rD2).max()
    assert np.abs(Dt2w - rDt2[-1]).max() <= 1e-12 * np.abs(rDt2).max()
    r2, r22 = so.mean_gradients(g["vm"], rD1, rD2, g["deta"], g["d2eta"])
    # D1/D2 rows have entries up to O(N^2)/O(N^4) that cancel: compare against that rounding scale
    s1 = (np.abs(rD1) @ np.abs(g["vm"])) * np.abs(g["deta"])[:, None]
    s2 = (np.abs(rD2) @ np.abs(g["vm"])) * (g["deta"] ** 2)[:, None] + s1
    assert np.all(np.abs(g2 - r2) <= 8e-16 * p.ny * s1 + 1e-300)
    assert np.all(np.abs(g22 - r22) <= 8e-16 * p.ny * s2 + 1e-300)


def test_circh_matches_oracle():
    y = so.sgengrid(64, 0.05, 26.028)[0]
    x_out, h5 = sb.circh(500.0, y)
    rx, *hm = so.circh(500.0, y)
    assert x_out == rx == 0.0
    ref = np.stack(hm, axis=1)
    assert np.abs(h5 - ref).max() <= 1e-15 * np.abs(ref).max() + 1e-30


@pytest.mark.parametrize("radius", [500.0, 100.0, 1.0e4, 37.5, -500.0])
def test_circh_closed_form_against_literal(radius):
    """The product evaluates circh's metrics in closed form (circle: h = 1 + r/R, dh/dr = 1/R, the rest 0); the oracle restates
    circh.f90:35-188 literally.  They agree to 1 ulp in h, 2 ulp in dh/dr and to the literal evaluation's own rounding residue
    (<= 4 eps / R^2) in the three vanishing terms, for the radii of the reference decks (FSCtest / CFtest: 500) and others."""
    y = so.sgengrid(48, 0.05, 30.0)[0]
    x_out, h5 = sb.circh(radius, y)
    rx, h, dhds, dhdr, dhdsr, dhdrr = so.circh(radius, y)
    assert x_out == rx == 0.0
    assert np.all(np.abs(h5[:, 0] - h) <= np.spacing(h))
    assert np.all(np.abs(h5[:, 2] - dhdr) <= 2 * np.spacing(dhdr))
    eps = np.finfo(float).eps
    for k, lit, bound in ((1, dhds, 1e-30), (3, dhdsr, 1e-30), (4, dhdrr, 4 * eps / radius ** 2)):   # d2h/dr2: cancellation of 1/R^2 terms
        assert np.all(h5[:, k] == 0.0) and np.abs(lit).max() <= bound


def test_sweep_enumeration_and_shards():
    a, b = sb.mtemporal_points(0.1, 0.5, 0.1, 0.0, 0.0, 1.0)
    ref = so.mtemporal_points(0.1, 0.5, 0.1, 0.0, 0.0, 1.0)
    assert a.size == len(ref) == 4 and np.array_equal(a, [r[1] for r in ref])
    o, b2 = sb.mspatial_points(0.02, 0.05, 0.01, 0.0, 0.0, 0.0)
    rs = so.mspatial_points(0.02, 0.05, 0.01, 0.0, 0.0, 0.0)
    assert np.array_equal(o, [r[0] for r in rs]) and o.size == 4
    # shards: contiguous, disjoint, cover everything
    for npts, world in ((10, 4), (256, 8), (3, 8), (10000, 8)):
        got = [sb.shard_range(npts, r, world) for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == npts
        assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in got]
        assert max(sizes) - min(sizes) <= 1


def test_eig_file_is_byte_identical_to_oracle_writer(tmp_path):
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0", ny=12)
    r = so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"])
    r.update(g); r["x_out"] = p.x
    blob = so.write_eig_file(p, r, 1, True)
    path = str(tmp_path / "evec.dat")
    sb.write_eig_file(path, to_params(p), 1, p.ind, p.omega, p.alpha, p.beta, p.x, g["y"], g["eta"], g["deta"], g["d2eta"],
                      r["omg"], r["evec"])
    assert open(path, "rb").read() == blob


def test_deck_reader_matches_oracle():
    for deck in ("ts_temporal_ny96.inp", "ts_spatial_ny32.inp", "fsc_spatial_ny64.inp", "cf_spatial_ny96.inp"):
        c = sb.read_deck(golden_text(deck))
        p = so.read_deck(golden_text(deck))
        q = c.params
        assert (q.ny, q.mattyp, q.wallt, q.top, q.curve, q.ider, q.ievec) == (p.ny, p.mattyp, p.wallt, p.top, p.curve, p.ider, p.ievec)
        assert (q.Ma, q.Re, q.Pr, q.yi, q.ymax) == (p.Ma, p.Re, p.Pr, p.yi, p.ymax)
        assert (c.itype, c.alpha, c.beta, c.omega, c.ind, c.x) == (p.itype, p.alpha, p.beta, p.omega, p.ind, p.x)
        assert (q.Te, q.rmue, q.rlme, q.cone) == (p.Te, p.rmue, p.rlme, p.cone)


def test_post_getevec_text_matches_oracle_and_golden_format():
    """stab_b200.post reproduces getevec's file (getevec.f90:193-227) character for character vs the oracle's writer,
    and numerically the reference golden file thesis/TStest/time.ref (abs 1e-8, the reference CI tolerance)."""
    import io
    p, g = oracle_case("ts_temporal_ny96.inp", "ts_profile.0")
    r = so.run_deck(p, golden_text("ts_profile.0"))
    target = complex(1.1467880189410E-001, 2.3844535276599E-003)
    txt = sb.post.getevec_text(1, p.Re, p.Ma, p.Pr, p.omega, p.alpha, p.beta, g["y"], r["omg"], r["evec"], value=target)
    j = so.select_mode(r["omg"], target)
    ref_txt = so.getevec_text(dict(Re=p.Re, Ma=p.Ma, Pr=p.Pr, omega=p.omega, alpha=p.alpha, beta=p.beta), r["omg"][j],
                              so.getevec_rows(g["y"], r["evec"][:, j], p.ny), 1)
    assert txt == ref_txt
    mine = np.loadtxt(io.StringIO(txt), comments="#")
    gold = np.loadtxt(io.StringIO(golden_text("ts_temporal_ny96.time.ref")), comments="#")
    assert np.abs(mine - gold).max() < 1e-8
    hdr = txt.splitlines()[1]
    assert hdr.startswith("# Omega = ( 1.146788018") and "E-001" in hdr


def test_post_mode_tracking():
    rng = np.random.default_rng(0)
    prm = np.linspace(0.1, 0.4, 12)
    mode = 0.3 * prm + 0.01j * (1 - (prm - 0.25) ** 2 * 40)
    spectra = [np.concatenate([rng.standard_normal(20) + 1j * rng.standard_normal(20) + 3, [m]]) for m in mode]
    for s in spectra:
        rng.shuffle(s)
    assert np.allclose(sb.post.track_nearest(spectra, mode[0]), mode)
    assert np.allclose(sb.post.track_extrapolated(spectra, prm, mode[0]), mode)


def test_cli_harness_fails_loudly_without_gpu(tmp_path):
    """host/stabgpu_cli (the C++ front end: `stab < temporal.inp`) parses the deck and the profile, then refuses to run
    without a CUDA device -- no CPU fallback."""
    import shutil, subprocess, torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cli = os.path.join(ROOT, "stab_b200", "stabgpu_cli")
    if not os.path.exists(cli):
        pytest.skip("stabgpu_cli not built")
    shutil.copy(os.path.join(GOLDEN, "ts_profile.0"), tmp_path / "profile.0")
    r = subprocess.run([cli], input=golden_text("ts_temporal_ny96.inp"), capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "no CUDA device" in r.stderr
