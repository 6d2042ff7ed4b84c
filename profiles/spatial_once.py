#!/usr/bin/env python
"""One pass of the spatial path (BASELINE configs[3]: omega sweep, Ny=128, companion order 1280) for ncu captures.
usage: python profiles/spatial_once.py [points] [vec]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import stab_b200 as sb

P = int(sys.argv[1]) if len(sys.argv) > 1 else 148
vec = len(sys.argv) > 2 and sys.argv[2] == "vec"
sb.init(0)
c = sb.read_deck(open(os.path.join(R, "tests", "golden", "ts_spatial_ny96.inp")).read())
c.params.ny = 128
c.load_profile(os.path.join(R, "tests", "golden", "ts_profile.0"))
om = np.linspace(0.02, 0.14, P) + 0j
pl = sb.Plan(2, c.params, c.vm, c.deta, c.d2eta, P, want_vectors=vec, h5=c.h5)
pl.upload(om, om * 0)
pl.execute()
print({k: round(v, 1) for k, v in pl.stage_times().items()})
if vec and os.environ.get("STAB_BREAKDOWN"):
    pl.profile_hessenberg(True)
    pl.execute()
    print("hessenberg", {k: round(v, 1) for k, v in pl.profile_hessenberg(False).items()})
    print("eigenvectors", {k: round(v, 1) for k, v in pl.profile_eigvec().items()})
