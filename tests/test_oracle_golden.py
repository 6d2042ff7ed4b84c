"""Pin the CPU oracle against the reference's own golden vectors.

Every number/target below is quoted from a reference file (path:line relative to the
reference checkout); the data files were copied into tests/golden/ by make_golden.py.
"""
import numpy as np
import pytest

import stab_oracle as so
from conftest import golden_text


def _run(deck, prof, **over):
    p = so.read_deck(golden_text(deck))
    for k, v in over.items():
        setattr(p, k, v)
    return p, so.run_deck(p, golden_text(prof))


def _ref_rows(name):
    return np.loadtxt(__import__("io").StringIO(golden_text(name)), comments="#")


def test_spatial_ny32_eigenvalue_and_eigenfunction():
    # test/space.1:3 (deck test/input.dat)
    p, r = _run("ts_spatial_ny32.inp", "ts_profile.0")
    target = complex(2.2805022654496E-001, -6.5136925762007E-003)
    j = so.select_mode(r["alp"], target)
    assert abs(r["alp"][j] - target) < 1e-12
    rows = so.getevec_rows(r["y"], r["evec"][:, j], p.ny)
    assert np.abs(rows - _ref_rows("ts_spatial_ny32.space.ref")).max() < 1e-10


def test_temporal_ny96_thesis_time_ref():
    # thesis/TStest/time.ref:2, README.md:25; CI tolerance is abs 1e-8 (run.sh:36-39)
    p, r = _run("ts_temporal_ny96.inp", "ts_profile.0")
    target = complex(1.1467880189410E-001, 2.3844535276599E-003)
    j = so.select_mode(r["omg"], target)
    assert j == 479                      # compbl/run.sh:13 selects sorted index 480 (1-based)
    assert abs(r["omg"][j] - target) < 1e-10
    rows = so.getevec_rows(r["y"], r["evec"][:, j], p.ny)
    assert np.abs(rows - _ref_rows("ts_temporal_ny96.time.ref")).max() < 1e-8
    # temporal eigenvectors are scaled so that the max-|.| entry is exactly 1 (temporal.f90:867-879)
    col = r["evec"][:, j]
    assert col[np.argmax(np.abs(col))] == 1.0
    # >= 8 exactly-zero eigenvalues from the homogeneous Dirichlet rows
    assert np.sum(r["omg"] == 0) >= 8
    # sorted ascending by imaginary part
    assert np.all(np.diff(r["omg"].imag) >= 0)


def test_temporal_ny64_readme_value():
    # thesis/TStest/README.md:24
    p, r = _run("ts_temporal_ny96.inp", "ts_profile.0", ny=64)
    target = complex(1.1467880189148E-001, 2.3844535289045E-003)
    j = so.select_mode(r["omg"], target)
    assert abs(r["omg"][j] - target) < 1e-10


def test_spatial_ny96_thesis_space_ref():
    # thesis/TStest/space.ref:3 (profile there comes from `fsc`; the shipped one reproduces it)
    p, r = _run("ts_spatial_thesis_ny96.inp", "ts_profile.0")
    target = complex(2.2804739411180E-001, -6.5163146952626E-003)
    j = so.select_mode(r["alp"], target)
    # compbl/run.sh:22 selects sorted index 371 (1-based) on its own profile; the number of
    # lambda=0 (alpha=inf -> 0) modes ahead of it is rounding dependent
    assert j in (370, 371)
    assert abs(r["alp"][j] - target) < 1e-10
    rows = so.getevec_rows(r["y"], r["evec"][:, j], p.ny)
    assert np.abs(rows - _ref_rows("ts_spatial_thesis_ny96.space.ref")).max() < 1e-8


@pytest.mark.parametrize("ny,target", [
    (64, complex(2.2804739410500E-001, -6.5163146761218E-003)),   # TStest/README.md:9
    (96, complex(2.2804739411367E-001, -6.5163146912049E-003)),   # TStest/README.md:10
])
def test_spatial_readme_values(ny, target):
    p, r = _run("ts_spatial_ny96.inp", "ts_profile.0", ny=ny, ievec=0)
    j = so.select_mode(r["alp"], target)
    assert abs(r["alp"][j] - target) < 5e-11


def test_fsc_spatial_curve2_streett():
    # FSCtest/space.ref:1: "# 250 alpha_r alpha_i 0 0", rows freestream -> wall (older layout)
    p, r = _run("fsc_spatial_ny64.inp", "fsc_profile.0")
    target = complex(-4.6108596548503E-001, -7.0050272583092E-003)
    j = so.select_mode(r["alp"], target)
    assert j in (248, 249)          # header says 250 (1-based); +-1 from rounding-dependent zero modes
    assert abs(r["alp"][j] - target) < 1e-12
    assert r["x_out"] == 0.0             # circh overwrites x (circh.f90:47-49)
    rows = so.getevec_rows(r["y"], r["evec"][:, j], p.ny)[::-1]
    ref = _ref_rows("fsc_spatial_ny64.space.ref")
    assert np.abs(rows - ref).max() < 1e-10


def test_cf_spatial_curve2_ny64():
    # CFtest/README.md:12 (SGI, Ny=64) and CFtest/space.ref (64 rows, wall -> freestream)
    p, r = _run("cf_spatial_ny96.inp", "cf_profile.0", ny=64)
    target = complex(-3.7392297537875E+001, -2.9982641664422E-001)
    j = so.select_mode(r["alp"], target)
    assert abs(r["alp"][j] - target) / abs(target) < 1e-11
    rows = so.getevec_rows(r["y"], r["evec"][:, j], p.ny)
    ref = _ref_rows("cf_spatial_ny64.space.ref")
    assert np.abs(rows - ref).max() < 1e-9


def test_frozen_oracle_outputs():
    """The oracle today reproduces what it produced when the fixtures were made."""
    fz = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "oracle_frozen.npz"))
    p, r = _run("ts_temporal_ny96.inp", "ts_profile.0", ny=24)
    assert np.allclose(r["M"], fz["temporal_ny24_M"], rtol=1e-13, atol=1e-13)
    d = np.abs(np.sort_complex(r["omg"]) - np.sort_complex(fz["temporal_ny24_omg"]))
    assert d.max() < 1e-7 * np.abs(r["omg"]).max()


def test_record_roundtrip_and_layout():
    p, r = _run("ts_temporal_ny96.inp", "ts_profile.0", ny=16)
    blob = so.write_eig_file(p, r, 1, True)
    n = 5 * p.ny
    # record sizes of SURVEY section 5: 40, 72, (3+4ny)*8, 16n, 16n^2, each framed by 2 int32
    assert len(blob) == sum(8 + s for s in (40, 72, (3 + 4 * p.ny) * 8, 16 * n, 16 * n * n))
    back = so.read_eig_file(blob)
    assert back["ny"] == p.ny and back["itype"] == 1
    assert np.array_equal(back["eval"], r["omg"])
    assert np.array_equal(back["evec"], r["evec"])
    assert back["alpha"] == p.alpha and back["Re"] == p.Re


def test_sweep_enumeration_quirks():
    # mtemporal.f90:25 excludes the upper end; mspatial.f90:72 includes it
    pts = so.mtemporal_points(0.1, 0.5, 0.1, 0.0, 0.0, 1.0)
    assert [iv for iv, _, _ in pts] == [1, 2, 3, 4]
    assert abs(pts[-1][1] - 0.4) < 1e-15
    sp = so.mspatial_points(0.02, 0.05, 0.01, 0.0, 0.0, 0.0)
    assert len(sp) == 4
    assert so.makename("eig", 12) == "eig.12"


def test_chebyd_properties():
    D = so.chebyd(16)
    x = np.cos(np.pi * np.arange(17) / 16)
    assert np.abs(D @ np.ones(17)).max() < 1e-11            # derivative of a constant
    assert np.abs(D @ x ** 3 - 3 * x ** 2).max() < 1e-10    # exact for low-order polynomials
