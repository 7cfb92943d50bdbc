"""BASELINE config 4 on one GPU's shard: relxillCp / relxilllpCp (10 zones), 8192 uniform-random vectors, bench-size
6-D table (5.2 GB, no row sharing between vectors -> the xillver gather really comes from HBM).  Prints kernel times and
the achieved HBM rate of k_xill."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import relxill_b200 as rx  # noqa: E402
from common import default_grid, sample_params  # noqa: E402
from relxill_b200 import _lib  # noqa: E402
from relxill_b200.tables import synth  # noqa: E402

tdir = synth.generate(synth.default_table_dir("bench"), "bench", ("rel", "lp", "rrad", "xillcp"))
rx.init(tdir, 0)
rx.set_cache(False)
e = default_grid(3000)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for model in ("relxillCp", "relxilllpCp"):
    P = sample_params(model, n, seed=99)
    if model == "relxilllpCp":
        P[:, 14] = 0
    b = rx.Batch(model, e, P)
    out = torch.zeros((n, 3000), dtype=torch.float64, device="cuda")
    for _ in range(2):
        b.run(out.data_ptr())
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(3):
        b.run(out.data_ptr())
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 3
    _lib.lib().relxill_b200_set_profiling(1)
    b.run(out.data_ptr())
    torch.cuda.synchronize()
    kt = b.kernel_times()
    _lib.lib().relxill_b200_set_profiling(0)
    ab = b.algorithmic_bytes()
    ok = int((b.status() == 0).sum())
    xk = kt.get("k_xill", (0.0, 1))[0]
    print(json.dumps({"model": model, "n": n, "ok": ok, "ms": ms, "spectra_per_s": n / ms * 1e3,
                      "kernels_ms": {k: round(v[0], 3) for k, v in kt.items()},
                      "xillver_distinct_GB": ab["xillver"] / 1e9, "rows_per_vector": ab["distinct_rows"] / n,
                      "k_xill_GBps": ab["xillver"] / (xk * 1e-3) / 1e9 if xk else None}))
