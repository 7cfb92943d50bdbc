"""Small evaluation of every model family, meant to run under compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool racecheck python scripts/sanitize.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import relxill_b200 as rx  # noqa: E402
from common import ALL_MODELS, default_grid, sample_params  # noqa: E402
from relxill_b200.tables import synth  # noqa: E402

tdir = synth.generate(synth.default_table_dir("test"), "test")
rx.init(tdir)
e = default_grid(400)
fin = np.exp(-0.5 * ((np.log(0.5 * (e[1:] + e[:-1])) - np.log(6.4)) / 0.03) ** 2) + 1e-3
for zones in (None, 50):
    rx.set_num_zones(zones)
    for m in ALL_MODELS:
        if zones and not m.startswith("relxilllp"):
            continue
        P = sample_params(m, 3, seed=7)
        f, st = rx.batch_eval(m, e, P, fin if m.startswith("relconv") else None, return_status=True)
        # a second call with one parameter changed goes through the retained batch (state cache)
        P[1, 0] *= 1.01
        f2 = rx.batch_eval(m, e, P, fin if m.startswith("relconv") else None)
        print(m, zones, st.tolist(), float(np.nansum(f)), float(np.nansum(f2)))
# a line model on a grid far finer than the profile: the deep queue of k_line fills inside a tile (resume path), wide zones
ef = np.linspace(0.25, 1.45, 6001) * 6.4
Pf = sample_params("relline", 2, seed=71)
Pf[:, 0] = 6.4
ff, stf = rx.batch_eval("relline", ef, Pf, return_status=True)
print("relline fine grid", stf.tolist(), float(np.nansum(ff)))
print("done")
