#!/usr/bin/env python
"""Build tests/golden/ from the reference checkout (run in the dev container only).

1. Copies the reference's own DATA fixtures for the hot path (mean-flow profiles, decks and
   the golden eigenfunction files its CI diffs against) -- no source code.
2. Freezes oracle outputs (eigenvalues of small cases) into `oracle_frozen.npz` so that a
   drift of the oracle itself (e.g. a different LAPACK build) is visible.

Usage: python tests/golden/make_golden.py [/root/reference]
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import stab_oracle as so  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"

COPIES = {
    # reference path                     -> name under tests/golden/
    "TStest/profile.0": "ts_profile.0",                 # == test/profile.0 (md5 identical)
    "FSCtest/profile.0": "fsc_profile.0",
    "CFtest/profile.0": "cf_profile.0",
    "test/input.dat": "ts_spatial_ny32.inp",
    "test/space.1": "ts_spatial_ny32.space.ref",
    "test/spec.inp": "ts_spatial_ny64.inp",
    "TStest/input.dat": "ts_spatial_ny96.inp",
    "thesis/TStest/temporal.inp": "ts_temporal_ny96.inp",
    "thesis/TStest/time.ref": "ts_temporal_ny96.time.ref",
    "thesis/TStest/spatial.inp": "ts_spatial_thesis_ny96.inp",
    "thesis/TStest/space.ref": "ts_spatial_thesis_ny96.space.ref",
    "FSCtest/input.dat": "fsc_spatial_ny64.inp",
    "FSCtest/space.ref": "fsc_spatial_ny64.space.ref",
    "CFtest/input.dat": "cf_spatial_ny96.inp",
    "CFtest/space.ref": "cf_spatial_ny64.space.ref",
    # the thesis crossflow case: decks of the external `fsc` mean-flow solver and the CI's golden eigenfunctions
    "thesis/TStest/blasius.inp": "ts_thesis_fsc.inp",
    "thesis/CFtest/fsc.inp": "cf_thesis_fsc.inp",
    "thesis/CFtest/temporal.inp": "cf_thesis_temporal_ny96.inp",
    "thesis/CFtest/time.ref": "cf_thesis_temporal_ny96.time.ref",
    "thesis/CFtest/spatial.inp": "cf_thesis_spatial_ny96.inp",
    "thesis/CFtest/space.ref": "cf_thesis_spatial_ny96.space.ref",
}


def main():
    for src, dst in COPIES.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
    frozen = {}
    prof = open(os.path.join(HERE, "ts_profile.0")).read()
    # small temporal and spatial cases, eigenvalues only
    for ny in (16, 24, 32):
        p = so.read_deck(open(os.path.join(HERE, "ts_temporal_ny96.inp")).read())
        p.ny = ny
        r = so.run_deck(p, prof, want_vectors=False)
        frozen[f"temporal_ny{ny}_omg"] = r["omg"]
        frozen[f"temporal_ny{ny}_M"] = r["M"]
        p = so.read_deck(open(os.path.join(HERE, "ts_spatial_ny32.inp")).read())
        p.ny = ny
        r = so.run_deck(p, prof, want_vectors=False)
        frozen[f"spatial_ny{ny}_alp"] = r["alp"]
    np.savez_compressed(os.path.join(HERE, "oracle_frozen.npz"), **frozen)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
