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
print("# one process, one stabgpu_temporal_batch call per line; destination arrays pageable (numpy: the library's pinned staging ring + host copy threads)\n# or page-locked once with stabgpu_host_register (direct DMA)")
base = np.linspace(0.05, 0.45, per, endpoint=False) + 0j
n = 5 * c.params.ny
ref = None
for k in (1, 2, 4, 8):
    if k > ndev:
        break
    got = sb.init_multi(k)
    P = per * k
    a = np.tile(base, k)                          # every shard solves the same `per` points: shard r must equal shard 0 bit for bit
    for mode in ("pageable", "registered"):
        omg = np.empty((P, n), dtype=np.complex128)
        ev = np.empty((P, n, n), dtype=np.complex128)
        info = np.zeros(P, dtype=np.int32)
        if mode == "registered":
            sb.host_register(ev); sb.host_register(omg)
        sb.temporal_batch(c.params, c.vm, c.deta, c.d2eta, a, a * 0, want_vectors=True, out=(omg, ev, info))      # warm-up: plans, rings
        t0 = time.perf_counter()
        sb.temporal_batch(c.params, c.vm, c.deta, c.d2eta, a, a * 0, want_vectors=True, out=(omg, ev, info))
        dt = time.perf_counter() - t0
        if ref is None:
            ref = (omg[:per].copy(), ev[:per].copy())
        same = all(np.array_equal(omg[r * per:(r + 1) * per], ref[0]) and np.array_equal(ev[r * per:(r + 1) * per], ref[1]) for r in range(k))
        print(f"devices {got}, {mode:10s} destination: {P} points in {dt * 1e3:.0f} ms = {P / dt:.0f} eigensolves/s, failed {int(np.count_nonzero(info))}, "
              f"every shard bit-identical to the 1-GPU result: {bool(same)}", flush=True)
        if mode == "registered":
            sb.host_unregister(ev); sb.host_unregister(omg)
        del omg, ev
