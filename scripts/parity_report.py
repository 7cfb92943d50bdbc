"""Largest relative error (bins above 1e-6 of the peak) of the CUDA path against the golden vectors produced by the
unmodified reference, per model."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import relxill_b200 as rx  # noqa: E402
from common import ALL_MODELS, NSCO_MODELS, relerr  # noqa: E402
from relxill_b200.tables import synth  # noqa: E402

tdir = synth.generate(synth.default_table_dir("test"), "test")
rx.init(tdir, 0)
g1 = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
g2 = np.load(os.path.join(ROOT, "tests", "golden", "golden_v2_nsco.npz"))
out = {}
for m in ALL_MODELS:
    g = g2 if m in NSCO_MODELS else g1
    fin = g["conv_input"] if m.startswith("relconv") else None
    got = rx.batch_eval(m, g["energy"], g[f"{m}_params"], fin)
    out[m] = max(relerr(a, b) for a, b in zip(got, g[f"{m}_flux"]))
rx.set_num_zones(50)
for m in ("relxilllp", "relxilllpCp"):
    got = rx.batch_eval(m, g1["energy"], g1[f"{m}_z50_params"])
    out[m + " (50 zones)"] = max(relerr(a, b) for a, b in zip(got, g1[f"{m}_z50_flux"]))
print(json.dumps({"max_relative_error_vs_reference_golden": out, "worst": max(out.values())}))
