"""Single-evaluation latency through the XSPEC symbols (what a fit sees), with and without the state cache."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import relxill_b200 as rx  # noqa: E402
from common import default_grid  # noqa: E402
from relxill_b200.tables import synth  # noqa: E402

tdir = synth.generate(synth.default_table_dir("test"), "test")
rx.init(tdir, 0)
e = default_grid(3000)
out = {}
for model, i_rel, i_x in (("relline", 4, None), ("relxill", 3, 9), ("relxilllp", 0, 8), ("relxillCp", 1, 9), ("relxilllpCp", 4, 7)):
    p0 = rx.default_params(model)
    res = {}
    for label, cache, idx in (("all parameters change", False, i_rel), ("xillver parameter changes, cache on", True, i_x),
                              ("nothing changes, cache on", True, None)):
        if label.startswith("xillver") and i_x is None:
            continue
        rx.set_cache(cache)
        rx.lmod(model, e, p0)   # warm: tables, retained batch
        ts = []
        for k in range(12):
            p = p0.copy()
            if idx is not None:
                p[idx] *= 1.0 - 0.004 * (k + 1)
            t0 = time.perf_counter()
            rx.lmod(model, e, p)
            ts.append(time.perf_counter() - t0)
        res[label] = round(1e3 * float(np.median(ts)), 3)
    out[model] = res
rx.set_cache(True)
print(json.dumps({"unit": "ms per lmod call, 3000 bins, median of 12", "latency": out}))
