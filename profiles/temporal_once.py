"""Three plan executes of the bench workload (Ny=128 TS alpha sweep, eigenvectors on) for profiler runs:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python profiles/temporal_once.py 296
  ncu --set full --clock-control none --import-source on -k regex:k_invit -c 1 -o gpurun_out/invit python profiles/temporal_once.py 148"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import stab_b200 as sb
P = int(sys.argv[1]) if len(sys.argv) > 1 else 148
sb.init(0)
c = sb.read_deck(open(os.path.join(R, "tests", "golden", "ts_temporal_ny96.inp")).read())
c.params.ny = 128
c.load_profile(os.path.join(R, "tests", "golden", "ts_profile.0"))
a = np.linspace(0.05, 0.45, P, endpoint=False) + 0j
pl = sb.Plan(1, c.params, c.vm, c.deta, c.d2eta, P, want_vectors=True)
pl.upload(a, a * 0)
pl.execute(); pl.execute(); pl.execute()
print({k: round(v, 1) for k, v in pl.stage_times().items()})
