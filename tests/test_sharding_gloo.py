"""Multi-process host logic of the sharded sweep on CPU: world_size 2, gloo.
Each rank takes stabgpu_shard_range of the mtemporal point list, produces its block of results and
the final gather reassembles the reference's loop order (iver numbering).  The hot path itself
needs a GPU; here the per-point 'result' is a deterministic stand-in so that only the
partitioning / gather plumbing is exercised."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import stab_b200 as sb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = sb.mtemporal_points(0.05, 0.45, 0.4 / 10, 0.0, 0.2, 0.1)      # 10 x 2 points
    npts, n = a.size, 6
    lo, hi = sb.shard_range(npts, rank, world)
    mine = np.stack([(a[lo:hi] + 1j * b[lo:hi]) * (k + 1) for k in range(n)], axis=1)   # (hi-lo, n) stand-in spectra
    # equal-sized blocks for all_gather: pad to the largest shard
    cap = max(sb.shard_range(npts, r, world)[1] - sb.shard_range(npts, r, world)[0] for r in range(world))
    buf = torch.zeros((cap, n), dtype=torch.complex128)
    buf[: hi - lo] = torch.from_numpy(mine)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    full = np.concatenate([parts[r][: sb.shard_range(npts, r, world)[1] - sb.shard_range(npts, r, world)[0]].numpy()
                           for r in range(world)])
    if rank == 0:
        np.save(out, full)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sweep_gather_world2(tmp_path):
    sys.path.insert(0, ROOT)
    import stab_b200 as sb
    out = str(tmp_path / "full.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    full = np.load(out)
    a, b = sb.mtemporal_points(0.05, 0.45, 0.4 / 10, 0.0, 0.2, 0.1)
    ref = np.stack([(a + 1j * b) * (k + 1) for k in range(6)], axis=1)
    assert full.shape == ref.shape == (20, 6)
    assert np.array_equal(full, ref)          # reference loop order (iver) restored
