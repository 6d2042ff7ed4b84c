#!/usr/bin/env python
"""bench.py -- full-spectrum eigensolves/sec on the TS temporal alpha-sweep at Ny=128 (BASELINE.json
configs[1]), one process per GPU.

A "step" is one pass of the hot path (assembly -> B0^-1 A0 -> balance -> Hessenberg -> QR ->
sort [-> eigenvectors]) over this rank's shard of the sweep.  `value` is device-resident
throughput (sweep values already in HBM); `e2e` is the same metric through the reference-facing
C-ABI call `stabgpu_temporal_batch` with pinned HOST buffers, H2D/D2H inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--ny 128] [--no-vectors]
  python bench.py --impl reference ...     # the reference's CPU arithmetic (oracle port) on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=296, help="sweep points per GPU per step (weak scaling)")
    ap.add_argument("--ny", type=int, default=128)
    ap.add_argument("--no-vectors", action="store_true", help="eigenvalues only (the reference always computes vectors)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="points in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: a FIXED sweep (BASELINE configs[4], --config c5) sharded over the ranks, gather timed")
    ap.add_argument("--config", default="c2", choices=["c2", "c5"], help="c2: TS alpha sweep Ny=128 (headline); c5: neutral-curve (alpha,Re) points Ny=256, values only")
    ap.add_argument("--total-points", type=int, default=2048, help="points of the fixed sweep in --scaling strong")
    ap.add_argument("--no-context", action="store_true", help="skip the library eig context number and the parity sample")
    return ap.parse_args()


# ---- workload: TStest profile, thesis TS deck, alpha sweep (mtemporal enumeration) ---------------
def workload(ny, npts_total):
    """Returns (case, alpha[npts_total], beta[npts_total]).  alpha in [0.05, 0.45)."""
    import stab_b200 as sb
    deck = open(os.path.join(ROOT, "tests", "golden", "ts_temporal_ny96.inp")).read()
    c = sb.read_deck(deck)
    c.params.ny = ny
    c.load_profile(os.path.join(ROOT, "tests", "golden", "ts_profile.0"))
    a, b = sb.mtemporal_points(0.05, 0.45, 0.4 / npts_total, 0.0, 0.1, 1.0)
    assert a.size == npts_total, (a.size, npts_total)
    return c, a + 0j, b + 0j


# ---- CPU arm: the reference's arithmetic (oracle port: scipy-LAPACK zgesv + zgeev) ---------------
def _cpu_worker(args):
    ny, alphas, as_coded, want_vectors = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    import stab_oracle as so
    deck = open(os.path.join(ROOT, "tests", "golden", "ts_temporal_ny96.inp")).read()
    p = so.read_deck(deck)
    p.ny = ny
    p.finish()
    g = so.prepare(p, open(os.path.join(ROOT, "tests", "golden", "ts_profile.0")).read())
    t0 = time.perf_counter()
    out = []
    for a in alphas:
        p.alpha = complex(a)
        r = so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=want_vectors, as_coded=as_coded)
        out.append(r["omg"][:4])
    return time.perf_counter() - t0, len(out)


def _cpu_threaded_worker(args):
    """How the reference binary itself runs a sweep: ONE process, points in sequence, OpenBLAS using every core."""
    ny, alphas, want_vectors, threads = args
    os.environ["OPENBLAS_NUM_THREADS"] = str(threads)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import stab_oracle as so
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(threads)
    except Exception:
        pass
    deck = open(os.path.join(ROOT, "tests", "golden", "ts_temporal_ny96.inp")).read()
    p = so.read_deck(deck)
    p.ny = ny
    p.finish()
    g = so.prepare(p, open(os.path.join(ROOT, "tests", "golden", "ts_profile.0")).read())
    p.alpha = complex(alphas[0])
    so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=want_vectors, as_coded=True)     # warm-up
    t0 = time.perf_counter()
    for a in alphas:
        p.alpha = complex(a)
        so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=want_vectors, as_coded=True)
    return len(alphas) / (time.perf_counter() - t0)


def cpu_as_stab_runs(ny, want_vectors, npts=4):
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    with mp.get_context("spawn").Pool(1) as pool:
        rate = pool.map(_cpu_threaded_worker, [(ny, list(np.linspace(0.1, 0.4, npts)), want_vectors, cores)])[0]
    return {"value": rate, "unit": "eigensolves/s", "threads": cores,
            "what": f"one process, {npts} points in sequence, multi-threaded OpenBLAS, lwork=2n as coded (how the reference binary runs a sweep)"}


def cpu_arm(ny, alphas, want_vectors, cores=None):
    """One worker per core, one BLAS thread each (the natural CPU parallelisation of a sweep of
    independent points).  Returns (solves/s, cores, wall seconds)."""
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0))
    cores = max(1, min(cores, len(alphas)))
    chunks = [list(alphas[i::cores]) for i in range(cores)]
    ctx = mp.get_context("spawn")
    best = None
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(16, [0.3], False, want_vectors)] * cores)        # import + warm-up
        for as_coded in (False, True):       # optimal workspace vs the reference's lwork=2n: quote the faster
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(ny, ch, as_coded, want_vectors) for ch in chunks])
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, as_coded)
    return len(alphas) / best[0], cores, best[0], best[1]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    sample = args.cpu_sample or min(args.points, max(8 * cores, 64))
    alphas = np.linspace(0.05, 0.45, sample, endpoint=False)
    want_vectors = not args.no_vectors
    times = []
    for it in range(args.warmup + args.steps):
        rate, used, dt, as_coded = cpu_arm(args.ny, alphas, want_vectors)
        if it >= args.warmup:
            times.append(dt)
        if it == 0 and dt > 60:           # keep the whole run bounded
            times = [dt]
            break
    ms = 1e3 * float(np.mean(times))
    val = sample / (ms / 1e3)
    desc = (f"each step = {sample} points spread over the workload's {args.points}-point sweep (bounded sample), "
            f"one worker per core ({used}), 1 BLAS thread each; CPU arm is fixed at this box's host cores for every --gpus N")
    line = {
        "impl": "reference", "metric": "full-spectrum eigensolves/sec at Ny=%d" % args.ny, "value": val,
        "unit": "eigensolves/s", "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": config_dict(args, args.points),
        "sample_points_per_step": sample,
        "cpu_baseline": {"value": val, "unit": "eigensolves/s", "cores": used, "kind": "port", "sample": desc,
                         "lapack": "scipy OpenBLAS zgesv+zgeev('N','%s'), %s workspace" % ("V" if want_vectors else "N", "lwork=2n (as coded)" if as_coded else "optimal")},
        "e2e": {"value": val, "unit": "eigensolves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def config_dict(args, npts):
    return {"workload": "TStest Tollmien-Schlichting temporal alpha-sweep at fixed Re (BASELINE configs[1]; SURVEY C2): "
                        "TStest/profile.0, M=0.3 Re=1000 Pr=1, Yi=1 algebraic map, alpha in [0.05,0.45), beta=0",
            "ny": args.ny, "n": 5 * args.ny, "points_per_gpu_per_step": npts,
            "eigenvectors": not args.no_vectors,
            "l2": "per-step working set (points x 16 n^2 B >= 1.6 GB) exceeds the 126 MB L2; no explicit flush"}


# ---- clocks sampler -------------------------------------------------------------------------------
class Clocks:
    def __init__(self, dev):
        self.dev, self.rows, self.stop, self.proc = dev, [], False, None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"
        base = ["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + q, "--format=csv,noheader,nounits"]
        # one streaming nvidia-smi (a sample every 100 ms) instead of one process per sample: a 3-step timed region of
        # ~0.9 s then holds ~8 samples rather than 1-2
        try:
            self.proc = subprocess.Popen(base + ["-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                f = [x.strip() for x in line.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
                if self.stop:
                    break
            return
        except Exception:
            self.proc = None
        while not self.stop:
            try:
                o = subprocess.run(base, capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.proc is not None:
            try:
                self.proc.terminate()
                self.proc.wait(timeout=3)
            except Exception:
                try:
                    self.proc.kill()
                except Exception:
                    pass
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(float(r[6]) for r in self.rows)}


class DevArray:
    """__cuda_array_interface__ wrapper of a raw device pointer (for torch.as_tensor)."""

    def __init__(self, ptr, nfloat64):
        self.__cuda_array_interface__ = {"shape": (nfloat64,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def fp64_gemm_peak(torch, dev):
    """cuBLAS DGEMM 4096^3 burst TFLOP/s -- MEASURED_PEAKS.json has no FP64 entry."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(2):
        a @ b
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = max(best, 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    return best


def parity_sample(ny, alpha, omg, ev, k=4):
    """After the timed loops: `k` points of the timed sweep (rank 0's shard, from the e2e call's output) against the
    oracle -- the checker, never the thing measured.  max_rel_phys: worst relative distance of a matched eigenvalue over
    the physical window |omega| < 2 (denominator floored at 1e-3); max_resid: worst ||M v - w v|| / (||M||_F ||v||)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import stab_oracle as so
    from helpers import match_spectra
    deck = open(os.path.join(ROOT, "tests", "golden", "ts_temporal_ny96.inp")).read()
    p = so.read_deck(deck)
    p.ny = ny
    p.finish()
    g = so.prepare(p, open(os.path.join(ROOT, "tests", "golden", "ts_profile.0")).read())
    idx = np.linspace(0, len(alpha) - 1, k).round().astype(int)
    worst_rel = worst_res = worst_ts = 0.0
    for j in idx:
        p.alpha = complex(alpha[j])
        r = so.solve_temporal(p, g["vm"], g["deta"], g["d2eta"], want_vectors=False)
        ref = r["omg"]
        _, d = match_spectra(ref, omg[j])
        phys = np.abs(ref) < 2.0
        worst_rel = max(worst_rel, float((d[phys] / np.maximum(np.abs(ref[phys]), 1e-3)).max()))
        jm = int(np.argmax(np.where(phys, ref.imag, -np.inf)))       # the least stable discrete mode: what a stability analysis reads
        worst_ts = max(worst_ts, float(d[jm] / abs(ref[jm])))
        if ev is not None:
            V = ev[j].T                                        # library layout (column, row) -> eigenvectors in columns
            R = r["M"] @ V - V * omg[j][None, :]
            worst_res = max(worst_res, float((np.linalg.norm(R, axis=0) / (np.linalg.norm(r["M"]) * np.linalg.norm(V, axis=0))).max()))
    return {"points": [int(j) for j in idx], "max_rel_phys": worst_rel, "max_rel_least_stable_mode": worst_ts,
            "max_resid": worst_res if ev is not None else None,
            "note": "max_rel_phys includes the ill-conditioned continuous-branch modes, where two LAPACK runs on the same matrix differ "
                    "by ~3e-10 (tests/helpers.py::spectrum_parity is the per-mode, condition-aware gate; profiles/r02_parity.json)",
            "oracle": "oracle/stab_oracle.py (scipy LAPACK zgesv + zgeev, lwork=2n as coded)"}


def library_eig_context(torch, dev, ny, alpha, nmat=4):
    """Context only: the vendor-library route on the same box -- torch.linalg.eig on the device (cuSOLVER
    Xgeev / MAGMA hybrid, whichever torch dispatches) on `nmat` assembled operators of the sweep, eigenvectors on."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import stab_oracle as so
        deck = open(os.path.join(ROOT, "tests", "golden", "ts_temporal_ny96.inp")).read()
        p = so.read_deck(deck)
        p.ny = ny
        p.finish()
        g = so.prepare(p, open(os.path.join(ROOT, "tests", "golden", "ts_profile.0")).read())
        mats = []
        for a in alpha[:nmat]:
            p.alpha = complex(a)
            A0, B0, _ = so.assemble_temporal(p, g["vm"], g["deta"], g["d2eta"])
            mats.append(np.linalg.solve(B0, A0))
        M = torch.as_tensor(np.stack(mats), device=dev)
        torch.linalg.eig(M[:1])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        torch.linalg.eig(M)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {"library_eig": {"value": nmat / dt, "unit": "eigensolves/s", "matrices": nmat,
                                "what": "torch.linalg.eig (complex128, eigenvectors) on cuda:0 on the same operators, wall clock; "
                                        "torch %s dispatches to its CUDA linalg backend (cuSOLVER Xgeev / MAGMA hybrid)" % torch.__version__}}
    except Exception as ex:                                   # context only: never fail the bench line
        return {"library_eig": {"unavailable": str(ex)[:200]}}


def c5_workload(ny, total):
    """BASELINE configs[4]: neutral-curve sweep, 100 x 100 grid Re in logspace(2.5, 4) x alpha in
    linspace(0.02, 0.5), beta = 0, eigenvalues only; `total` of the 10^4 points, spread evenly over the grid in the
    reference's loop order.  Mean flow: TStest/profile.0 with per-point Re overrides (the boundary-layer similarity
    profile does not depend on Re)."""
    import stab_b200 as sb
    deck = open(os.path.join(ROOT, "tests", "golden", "ts_temporal_ny96.inp")).read()
    c = sb.read_deck(deck)
    c.params.ny = ny
    c.load_profile(os.path.join(ROOT, "tests", "golden", "ts_profile.0"))
    Re = np.logspace(2.5, 4.0, 100)
    al = np.linspace(0.02, 0.5, 100)
    RR, AA = np.meshgrid(Re, al, indexing="ij")
    idx = np.linspace(0, RR.size - 1, total).round().astype(int)
    return c, AA.ravel()[idx] + 0j, RR.ravel()[idx].copy()


def run_strong(args, torch, sb, dist, world, rank, local, dev):
    """Strong scaling on BASELINE configs[4]: a FIXED sweep of --total-points (alpha, Re) points at Ny = 256, eigenvalues
    only, sharded contiguously over the ranks (stabgpu_shard_range); a step = H2D of the shard's sweep values + the hot
    path + the full eigenvalue gather (NCCL all-gather of the padded shards, then rank 0's D2H of all of them)."""
    ny = 256 if args.ny == 128 else args.ny
    n = 5 * ny
    total = args.total_points
    case, alpha_all, re_all = c5_workload(ny, total)
    lo, hi = sb.shard_range(total, rank, world)
    mine = hi - lo
    pad = (total + world - 1) // world
    alpha, Re = alpha_all[lo:hi], re_all[lo:hi]
    plan = sb.Plan(1, case.params, case.vm, case.deta, case.d2eta, pad, want_vectors=False)
    if plan.capacity < mine:
        raise SystemExit(f"bench.py: {mine} points do not fit the device workspace (capacity {plan.capacity})")
    stream = torch.cuda.ExternalStream(plan.stream(), device=dev)
    eig_dev = torch.as_tensor(DevArray(plan.eig_dev(), 2 * n * pad), device=dev)
    gathered = torch.empty(world * eig_dev.numel(), dtype=torch.float64, device=dev)
    host = torch.empty(gathered.numel(), dtype=torch.float64, pin_memory=True) if rank == 0 else None
    parts = {"upload": 0.0, "execute": 0.0, "gather": 0.0}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step(acc=None):
        t0 = time.perf_counter()
        plan.upload(alpha, alpha * 0, Re_pt=Re)
        t1 = time.perf_counter()
        plan.execute()
        t2 = time.perf_counter()
        if dist is not None:
            dist.all_gather_into_tensor(gathered, eig_dev)
        else:
            gathered.copy_(eig_dev)
        if rank == 0:
            host.copy_(gathered, non_blocking=False)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if acc is not None:
            acc["upload"] += t1 - t0; acc["execute"] += t2 - t1; acc["gather"] += t3 - t2

    for _ in range(args.warmup):
        step()
    barrier()
    with Clocks(local) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(parts)
        barrier()
        wall = time.perf_counter() - t0
    t = torch.tensor([wall, parts["execute"], -parts["execute"], parts["gather"]], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_max, ex_max, ex_min, ga_max = float(t[0]), float(t[1]), -float(t[2]), float(t[3])
    ms_step = 1e3 * wall_max / args.steps
    info = plan.info()
    nfail = torch.tensor([int(np.count_nonzero(info[:mine]))], device=dev)
    if dist is not None:
        dist.all_reduce(nfail)
    stage = plan.stage_times()
    launches = plan.launch_count() * args.steps
    if rank == 0:
        emit({
            "metric": "full-spectrum eigensolves/sec at Ny=%d" % ny, "value": total / (ms_step * 1e-3), "unit": "eigensolves/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
            "config": {"workload": "neutral-curve sweep (BASELINE configs[4]; SURVEY C5): %d of the 10^4 (alpha, Re) grid points, "
                                   "Re in logspace(2.5,4) x alpha in linspace(0.02,0.5), beta=0, TStest/profile.0, per-point Re" % total,
                       "ny": ny, "n": n, "total_points": total, "points_per_gpu_per_step": pad, "eigenvectors": False,
                       "l2": "per-step working set (points x 16 n^2 B = %.1f GB per GPU) exceeds the 126 MB L2; no explicit flush" % (mine * 16 * n * n / 1e9)},
            "clocks": clk.summary(), "gpu_launches": int(launches),
            "e2e": {"value": total / (ms_step * 1e-3), "unit": "eigensolves/s", "h2d_bytes_per_step": int(mine * 40),
                    "d2h_bytes_per_step": int(16 * n * pad * world),
                    "call": "plan upload (H2D sweep values) + execute + NCCL all-gather + rank-0 D2H of every eigenvalue, all inside the timed step"},
            "strong": {"execute_ms_per_step_max_rank": 1e3 * ex_max / args.steps, "execute_ms_per_step_min_rank": 1e3 * ex_min / args.steps,
                       "gather_ms_per_step_max_rank": 1e3 * ga_max / args.steps, "gather_bytes": int(16 * n * pad * world),
                       "stages_ms_rank0": stage},
            "failed_points": int(nfail.item()),
        })
    if dist is not None:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL's version banner, library
    chatter written to fd 1 from C) was redirected to stderr by main()."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    args = parse()
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import stab_b200 as sb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the stab hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    sb.init(local)

    if args.scaling == "strong":
        return run_strong(args, torch, sb, dist, world, rank, local, dev)

    want_vectors = not args.no_vectors
    P = args.points
    case, alpha_all, beta_all = workload(args.ny, P * world)
    lo, hi = sb.shard_range(P * world, rank, world)
    alpha, beta = alpha_all[lo:hi], beta_all[lo:hi]
    n = 5 * args.ny
    prm = case.params

    # ---- device-resident arm -------------------------------------------------------------------
    plan = sb.Plan(1, prm, case.vm, case.deta, case.d2eta, P, want_vectors=want_vectors)
    if plan.capacity < P:
        raise SystemExit(f"bench.py: {P} points do not fit the device workspace (capacity {plan.capacity})")
    plan.upload(alpha, beta)
    stream = torch.cuda.ExternalStream(plan.stream(), device=dev)
    eig_dev = torch.as_tensor(DevArray(plan.eig_dev(), 2 * n * P), device=dev)
    gathered = torch.empty(world * eig_dev.numel(), dtype=torch.float64, device=dev) if world > 1 else None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        plan.execute()
        if dist is not None:                        # the final result gather of the sweep
            dist.all_gather_into_tensor(gathered, eig_dev)

    for _ in range(args.warmup):
        step()
    stage_acc = {}
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with Clocks(local) as clk:
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            step()
            for k, v in plan.stage_times().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    # the gather (N > 1) runs on torch's stream after execute() has synchronised: wall covers it, events cover the kernels
    ms_total = max(dev_ms, 1e3 * wall) if world > 1 else dev_ms
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = P * world / (ms_step * 1e-3)
    launches = plan.launch_count() * args.steps
    stage_ms = {k: v / args.steps for k, v in stage_acc.items()}
    # one extra (untimed) execute with an event after every Hessenberg kernel: per-class breakdown
    plan.profile_hessenberg(True)
    plan.execute()
    hess_bd = plan.profile_hessenberg(False)
    evec_bd = plan.profile_eigvec() if want_vectors else {}
    ilohi = plan.ilohi()
    plan_info = np.zeros(P, dtype=np.int32)
    sb.lib().stabgpu_plan_download(plan._h, None, None, plan_info.ctypes.data)
    n_fail = int(np.count_nonzero(plan_info))

    # ---- e2e arm: the C-ABI batch call with pinned host buffers ----------------------------------
    omg_h = torch.empty((P, n), dtype=torch.complex128, pin_memory=True).numpy()
    ev_h = torch.empty((P, n, n), dtype=torch.complex128, pin_memory=True).numpy() if want_vectors else None
    info_h = torch.empty((P,), dtype=torch.int32, pin_memory=True).numpy()
    al_h = torch.empty((P,), dtype=torch.complex128, pin_memory=True).numpy(); al_h[:] = alpha
    be_h = torch.empty((P,), dtype=torch.complex128, pin_memory=True).numpy(); be_h[:] = beta
    plan.destroy()
    del eig_dev

    def e2e_step():
        sb.temporal_batch(prm, case.vm, case.deta, case.d2eta, al_h, be_h, want_vectors=want_vectors, out=(omg_h, ev_h, info_h))

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = P * world * args.steps / float(t.item())
    # the drop-in caller's arrays are pageable (a Fortran allocate): same call, numpy-allocated destination, served by the
    # library's pinned staging ring (stabgpu_set_host_staging)
    omg_p = np.empty((P, n), dtype=np.complex128)
    ev_p = np.empty((P, n, n), dtype=np.complex128) if want_vectors else None
    info_p = np.zeros(P, dtype=np.int32)
    if ev_p is not None:
        ev_p[:] = 0                                              # touch the pages once (first-touch cost is the caller's, not the call's)

    def e2e_pageable_step():
        sb.temporal_batch(prm, case.vm, case.deta, case.d2eta, np.array(alpha), np.array(beta), want_vectors=want_vectors, out=(omg_p, ev_p, info_p))

    e2e_pageable_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_pageable_step()
    barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_pageable_val = P * world * args.steps / float(t.item())
    same_bits = bool(np.array_equal(omg_p, omg_h) and (ev_p is None or np.array_equal(ev_p, ev_h)))
    ny = args.ny
    h2d = 2 * 16 * P + 8 * (ny * 5 * 3 + 3 * ny + 2 * ny * ny)
    d2h = 16 * n * P + (16 * n * n * P if want_vectors else 0) + 4 * P

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- rooflines: the graded Hessenberg stage, its HBM-bound GEMV kernel and its DMMA kernels -------
    peak = fp64_gemm_peak(torch, dev)
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"
    NB = 32
    hess_flops = gemv_bytes = gemm_flops = 0.0
    for ilo, ihi in ilohi:
        ilo, ihi = int(ilo), int(ihi)
        nh = ihi - ilo + 1
        hess_flops += (40.0 / 3.0) * nh ** 3 if nh > 2 else 0.0      # ZGEHRD on the active block ((40/3) n^3)
        k = ilo
        while k < ihi:
            for c in range(k, min(k + NB, ihi)):
                gemv_bytes += 16.0 * (ihi - k) * (ihi - c)           # y = A(k+1:ihi, c+1:ihi) v, one pass over the block
                gemv_bytes += 16.0 * (c - k) * (ihi - c)             # + t = V(:,0:j)^H v_j in the same launch (one pass over V(c+1:ihi, 0:j))
            nct = n - (k + NB)
            gemm_flops += 8.0 * NB * ((k + 1) * (ihi - k) + (k + 1) * (NB - 1))
            if nct > 0:
                gemm_flops += 8.0 * NB * ((ihi + 1) * max(ihi + 1 - k - NB, 0) + 2 * (ihi - k) * nct)
            k += NB
    hess_ms = stage_ms.get("hessenberg", 0.0)
    achieved = hess_flops / (hess_ms * 1e-3) / 1e12 if hess_ms > 0 else 0.0
    total_stage = sum(stage_ms.values()) or 1.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    # ceiling of a one-stage (ZGEHRD / ZLAHR2) reduction: every column needs y = A(k+1:, c+1:) v over the not yet
    # updated trailing matrix, 16 (ihi-k)(ihi-c) B from HBM per column (296 matrices = 1.9 GB do not fit the 126 MB L2);
    # even with every other kernel free the stage cannot run faster than those bytes at the measured HBM bandwidth
    cap_tf = hess_flops / (gemv_bytes / (hbm_peak * 1e9)) / 1e12 if gemv_bytes > 0 else None
    roofline = {"kernel": "Hessenberg stage (k_hb_panel_step + k_hb_gemv + k_gemm_pipe<*>): the graded stage of the north star",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "cap_frac": (cap_tf / peak) if (cap_tf and peak) else None, "target_frac": 0.5,
                "cap": "HBM ceiling of the per-column GEMV of a one-stage reduction: (40/3) nh^3 flops / (sum 16 (ihi-k)(ihi-c) B / measured HBM GB/s)",
                "traffic": None, "flops_per_step": hess_flops,
                "peak_source": "measured in this run: cuBLAS DGEMM 4096^3 burst via torch.matmul (MEASURED_PEAKS.json has no FP64 entry)",
                "kernel_share_of_step": hess_ms / total_stage}
    gemv_ms, gemm_ms = hess_bd.get("gemv", 0.0), hess_bd.get("gemm", 0.0)
    gemv_gbs = gemv_bytes / (gemv_ms * 1e-3) / 1e9 if gemv_ms > 0 else 0.0
    gemm_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    roofline_gemv = {"kernel": "k_hb_gemv (dominant kernel of the Hessenberg stage: the GEMV of column j and, in one extra CTA per matrix, the V^H v_j dot products)", "bound": "hbm", "achieved": gemv_gbs,
                     "peak": hbm_peak, "unit": "GB/s", "frac": gemv_gbs / hbm_peak if hbm_peak else None,
                     "algorithmic_bytes_per_step": gemv_bytes, "ms_per_step": gemv_ms, "launches_per_step": 32 * int(np.ceil((n - 1) / 32)),
                     "peak_source": hbm_src,
                     "traffic": (traffic or {}).get("k_hb_gemv"), "timing": "CUDA events after every kernel, separate profiled execute"}
    roofline_gemm = {"kernel": "k_gemm_pipe<*> (persistent cp.async-pipelined DMMA GEMM: rank-32 updates and V^H A products of the Hessenberg stage)", "bound": "tensor", "achieved": gemm_tf, "peak": peak,
                     "unit": "TFLOP/s", "frac": gemm_tf / peak if peak else None, "flops_per_step": gemm_flops, "ms_per_step": gemm_ms,
                     "traffic": (traffic or {}).get("k_gemm_pipe")}
    # the largest single kernel after round 2: register-resident inverse iteration.  Algorithmic work per eigenvalue: the UL
    # elimination carries the column and the right-hand side (2 complex FMAs per row below the pivot, sum_k 16 k = 8 n^2 flops)
    invit_ms = evec_bd.get("invit", 0.0)
    invit_flops = 8.0 * float(n) ** 3 * P
    invit_tf = invit_flops / (invit_ms * 1e-3) / 1e12 if invit_ms > 0 else 0.0
    roofline_invit = {"kernel": "k_invit<20,1> (inverse iteration in panel / bulk form: one warp per eigenvalue, pivot chain of a staged 8-column block on the boundary slot, recorded steps applied to the rows above as a DFMA stream; columns staged by TMA bulk copies)",
                      "bound": "fp64 vector (DFMA; 33.7 TFLOP/s measured on this pool in round 1, the DGEMM figure is used as the denominator)",
                      "achieved": invit_tf, "peak": peak, "unit": "TFLOP/s", "frac": invit_tf / peak if peak else None,
                      "flops_per_step": invit_flops, "ms_per_step": invit_ms}
    asm_ms = stage_ms.get("assemble", 0.0)
    asm_gbs = 16.0 * n * n * P / (asm_ms * 1e-3) / 1e9 if asm_ms > 0 else 0.0
    dominant = max(stage_ms, key=stage_ms.get)

    line = {
        "metric": "full-spectrum eigensolves/sec at Ny=%d" % ny, "value": value, "unit": "eigensolves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": config_dict(args, P),
        "clocks": clk.summary(),
        "e2e": {"value": e2e_val, "unit": "eigensolves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": "stabgpu_temporal_batch, pinned host buffers (direct DMA under the eigenvector stage)",
                "pageable": {"value": e2e_pageable_val, "unit": "eigensolves/s",
                             "call": "the same call with pageable (numpy) destination arrays: vectors staged through the library's "
                                     "pinned ring (4 x 64 MB, host cores / devices copy threads, 2..8)", "bit_identical_to_pinned": same_bits}},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_gemv": roofline_gemv,
        "roofline_gemm": roofline_gemm,
        "roofline_invit": roofline_invit,
        "stages_ms_per_step": stage_ms,
        "dominant_stage": (dominant + " (shifted QR: latency / FP64-vector bound, no clean roofline -- time share only, SURVEY 8d)") if dominant == "qr" else dominant,
        "qr": {"ms_per_step": stage_ms.get("qr", 0.0), "share_of_step": stage_ms.get("qr", 0.0) / total_stage,
               "note": "shifted QR with aggressive early deflation: latency bound (profiles/r02_qr_cycles.txt, r02_ncu_hqr.txt: 100 MB DRAM per matrix), no clean roofline"},
        "hessenberg_breakdown_ms": hess_bd,
        "eigvec_breakdown_ms": evec_bd,
        "assembly_roofline": {"bound": "hbm", "achieved": asm_gbs, "peak": hbm_peak, "unit": "GB/s",
                              "frac": asm_gbs / hbm_peak if hbm_peak else None},
        "failed_points": n_fail,
    }
    if not args.no_context:
        line["parity_sample"] = parity_sample(ny, alpha, omg_h, ev_h)
        line["context"] = library_eig_context(torch, dev, ny, alpha)
    if not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        sample = args.cpu_sample or min(P, max(16 * cores, 64))     # ~10-15 s of CPU work on the box's host cores
        rate, used, dt, as_coded = cpu_arm(ny, np.linspace(0.05, 0.45, sample, endpoint=False), want_vectors)
        line["cpu_baseline"] = {"value": rate, "unit": "eigensolves/s", "cores": used, "kind": "port",
                                "sample": f"{sample} points of the same sweep, one worker per core, 1 BLAS thread each, "
                                          f"{'lwork=2n' if as_coded else 'optimal workspace'} (faster of the two), {dt:.1f} s",
                                "as_stab_runs": cpu_as_stab_runs(ny, want_vectors)}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
