"""stab_b200 -- B200-native hot path of sscollis/stab behind a C ABI.

Python here is a thin ctypes binding over `libstabgpu.so` (built in-tree by
`__graft_entry__.build()` / `make -C stab_b200/csrc`).  It mirrors the reference's interface for
this path -- `temporal(name, ind)` (temporal.f90:2), `spatial(name, ind)` (spatial.f90:2),
`mtemporal` / `mspatial` sweeps (mtemporal.f90, mspatial.f90) -- with module `stuff` turned into
an explicit `Params` object.  There is no CPU fallback: if the shared library (or a CUDA device)
is missing, every compute call raises.
"""
from .binding import (  # noqa: F401
    LIB_PATH,
    Params,
    StabGpuError,
    Plan,
    chebyd,
    circh,
    device_info,
    edge_properties,
    getmean,
    init,
    init_multi,
    device_count,
    set_host_staging,
    set_qr_deflation,
    host_register,
    host_unregister,
    finalize,
    polish_batch,
    lib,
    mean_gradients,
    mspatial_points,
    mtemporal_points,
    read_profile,
    set_tuning,
    set_hess_mode,
    set_evec_mode,
    set_lu_mode,
    sgengrid,
    shard_range,
    spatial_assemble,
    spatial_batch,
    temporal_assemble,
    temporal_batch,
    temporal_polish,
    write_eig_file,
    zgeev_batch,
    debug_stages,
    debug_spatial_reduce,
)
from .api import Case, read_deck, spatial, temporal, mtemporal, mspatial, mspatial_stations, read_delta, stab  # noqa: F401
from . import post  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
