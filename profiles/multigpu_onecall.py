#!/usr/bin/env python
"""ONE serial process, ONE stabgpu_temporal_batch call, k GPUs behind it (stabgpu_init_multi): the drop-in case of
mtemporal.f90:29-39.  Prints eigensolves/s through the C ABI (host buffers in and out, eigenvectors on) for k = 1, 2, 4, 8
devices of the box and checks that every k returns the bits of k = 1.
usage: python profiles/multigpu_onecall.py [points_per_gpu] > profiles/r02_multigpu.txt"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import torch
import stab_b200 as sb

per = int(sys.argv[1]) if len(sys.argv) > 1 else 296
ndev = torch.cuda.device_count()
c = sb.read_deck(open(os.path.join(R, "tests", "golden", "ts_temporal_ny96.inp")).read())
c.params.ny = 128
c.load_profile(os.path.join(R, "tests", "golden", "ts_profile.0"))
print(f"# {ndev} x {torch.cuda.get_device_name(0)}; TS temporal alpha sweep, Ny=128 (n=640), eigenvectors on, {per} points per GPU per call")
print("# one process, one stabgpu_temporal_batch call per line (pageable numpy destination arrays: the library's pinned staging ring)")
ref = None
for k in (1, 2, 4, 8):
    if k > ndev:
        break
    got = sb.init_multi(k)
    P = per * k
    a = np.linspace(0.05, 0.45, P, endpoint=False) + 0j
    sb.temporal_batch(c.params, c.vm, c.deta, c.d2eta, a, a * 0, want_vectors=True)          # warm-up: plans, rings
    t0 = time.perf_counter()
    omg, ev, info = sb.temporal_batch(c.params, c.vm, c.deta, c.d2eta, a, a * 0, want_vectors=True)
    dt = time.perf_counter() - t0
    same = ""
    if k == 1:
        ref = (omg.copy(), ev.copy())
    else:
        same = f"  first {per} points bit-identical to the 1-GPU call: {bool(np.array_equal(omg[:per], ref[0]) and np.array_equal(ev[:per], ref[1]))}"
    print(f"devices {got}: {P} points in {dt * 1e3:.0f} ms = {P / dt:.0f} eigensolves/s, failed {int(np.count_nonzero(info))}{same}", flush=True)
    del omg, ev
