#!/usr/bin/env python
"""Throughput and stage times of the BASELINE.json configs other than the bench workload (C2 at Ny=64/96,
C3 crossflow (alpha, beta) grid, C4 spatial omega sweep with and without vectors, C5 neutral-curve points at Ny=256).
These are parity-test cases, not bench lines; this script records what they cost on one B200.
usage: python profiles/configs_bench.py [--cpu] > profiles/r02_configs.json   (--cpu adds the oracle port on the host cores)"""
import json
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import stab_b200 as sb
from stab_b200 import fsc

G = os.path.join(R, "tests", "golden")


def case(deck, ny, profile=None, table=None):
    c = sb.read_deck(open(os.path.join(G, deck)).read())
    c.params.ny = ny
    if table is not None:
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            open(os.path.join(d, "profile.0"), "w").write(fsc.format_table(table))
            c.load_profile(os.path.join(d, "profile.0"))
    else:
        c.load_profile(os.path.join(G, profile))
    return c


def run(name, kind, c, s1, s2, vec, Re=None, reps=2):
    P = len(s1)
    pl = sb.Plan(kind, c.params, c.vm, c.deta, c.d2eta, P, want_vectors=vec, h5=c.h5 if kind == 2 else None)
    pl.upload(s1 + 0j, s2 + 0j, Re_pt=Re)
    pl.execute()
    t0 = time.perf_counter()
    for _ in range(reps):
        pl.execute()
    dt = (time.perf_counter() - t0) / reps
    info = pl.info()
    out = dict(config=name, kind="temporal" if kind == 1 else "spatial", ny=c.params.ny, order=pl.N, points=P, eigenvectors=vec,
               ms=round(dt * 1e3, 1), eigensolves_per_s=round(P / dt, 1), stages_ms={k: round(v, 1) for k, v in pl.stage_times().items()},
               failed_points=int(np.count_nonzero(info)))
    pl.destroy()
    return out


def _cpu_worker(args):
    """one worker = one host core: the oracle port (scipy LAPACK, 1 BLAS thread) on `pts` points of a config"""
    deck, ny, kind, pts, vec = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(R, "oracle"))
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    import stab_oracle as so
    p = so.read_deck(open(os.path.join(G, deck)).read())
    p.ny = ny
    p.finish()
    g = so.prepare(p, open(os.path.join(G, "ts_profile.0")).read())
    hm = so.curvature_metrics(p, g["y"])[1] if kind == 2 else None
    t0 = time.perf_counter()
    for s1, re in pts:
        if re is not None:
            p.Re = float(re); p.finish()
        if kind == 1:
            p.alpha = complex(s1)
            so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=vec, as_coded=False)
        else:
            p.omega = complex(s1)
            so.solve_spatial(p, g["vm"], g["deta"], g["d2eta"], hm, want_vectors=vec, as_coded=False)
    return time.perf_counter() - t0


def cpu_leg(res):
    """CPU figures of the same configs: one oracle worker per host core, one point each (two for the
    small orders), 1 BLAS thread per worker, optimal LAPACK workspace -- the same arm bench.py quotes as cpu_baseline."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    plan = [("C2 TS temporal alpha sweep, Ny=64", "ts_temporal_ny96.inp", 64, 1, 4, True, None),
            ("C2 TS temporal alpha sweep, Ny=96", "ts_temporal_ny96.inp", 96, 1, 2, True, None),
            ("C3 crossflow", "ts_temporal_ny96.inp", 128, 1, 1, True, None),
            ("C4 spatial omega sweep, Ny=128 (companion order 1280), ievec=0", "ts_spatial_ny96.inp", 128, 2, 1, False, None),
            ("C4 spatial omega sweep, Ny=128 (companion order 1280), ievec=1", "ts_spatial_ny96.inp", 128, 2, 1, True, None),
            ("C5 neutral-curve", "ts_temporal_ny96.inp", 256, 1, 1, False, 1000.0)]
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_worker, [("ts_temporal_ny96.inp", 16, 1, [(0.3, None)], False)] * cores)     # import + warm-up
        for name, deck, ny, kind, per, vec, re in plan:
            base = 0.2 if kind == 1 else 0.08
            jobs = [(deck, ny, kind, [(base + 0.01 * (w * per + q), re) for q in range(per)], vec) for w in range(cores)]
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs)
            dt = time.perf_counter() - t0
            for r in res:
                if r["config"].startswith(name):
                    r["cpu_port"] = dict(eigensolves_per_s=round(cores * per / dt, 2), cores=cores, sample=f"{cores * per} points, one worker per core, {dt:.1f} s")
                    r["gpu_over_cpu"] = round(r["eigensolves_per_s"] / (cores * per / dt), 1)


def main():
    sb.init(0)
    res = []
    for ny in (64, 96):
        c = case("ts_temporal_ny96.inp", ny, "ts_profile.0")
        a, b = sb.mtemporal_points(0.05, 0.45, 0.4 / 296, 0.0, 0.1, 1.0)
        res.append(run(f"C2 TS temporal alpha sweep, Ny={ny}", 1, c, a, b, True))
    cf = fsc.profile_from_deck(open(os.path.join(G, "cf_thesis_fsc.inp")).read())
    c = case("cf_thesis_temporal_ny96.inp", 128, table=cf["table"])
    a, b = sb.mtemporal_points(-0.5, 0.0, 0.5 / 32, 0.1, 0.6, 0.5 / 32)
    res.append(run("C3 crossflow temporal (alpha, beta) 32 x 32 grid on the generated FSC profile, Ny=128", 1, c, a, b, True, reps=1))
    c = case("ts_spatial_ny96.inp", 128, "ts_profile.0")
    om = np.linspace(0.02, 0.14, 128)
    res.append(run("C4 spatial omega sweep, Ny=128 (companion order 1280), ievec=0", 2, c, om, om * 0, False))
    res.append(run("C4 spatial omega sweep, Ny=128 (companion order 1280), ievec=1", 2, c, om, om * 0, True, reps=1))
    c = case("ts_temporal_ny96.inp", 256, "ts_profile.0")
    Re = np.logspace(2.5, 4, 17)[:, None] * np.ones((1, 17))
    al = np.ones((17, 1)) * np.linspace(0.02, 0.5, 17)[None, :]
    res.append(run("C5 neutral-curve points (alpha, Re) 17 x 17 of the 100 x 100 sweep, Ny=256, eigenvalues only", 1, c, al.ravel(), al.ravel() * 0,
                   False, Re=Re.ravel(), reps=1))
    if "--cpu" in sys.argv:
        cpu_leg(res)
    name, sms, mem = sb.device_info()
    print(json.dumps(dict(device=name, sms=sms, note="one B200, device-resident plan execute (sweep values in HBM), wall clock around execute",
                          configs=res), indent=1))


if __name__ == "__main__":
    main()
