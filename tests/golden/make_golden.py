"""Generates tests/golden/golden_v1.npz (golden_v2_nsco.npz with `nsco`, golden_v3_configs.npz with `configs` as argument) from the UNMODIFIED reference (oracle/_ref) on the synthetic
'test' tables.  Run in the build container (needs /root/reference to have been compiled by
`make -C oracle ref`):   python tests/golden/make_golden.py
The fixture pins the oracle restatement and the CUDA path on machines without the reference."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import MODELS, NSCO_MODELS, config_cases, default_grid, sample_params  # noqa: E402
from oracle.pyref import Ref  # noqa: E402  (process-isolated)
from relxill_b200.tables import synth  # noqa: E402


def table_digest(d, keys=("rel", "lp", "xill", "xillcp", "rrad")):
    h = hashlib.sha256()
    for key in keys:
        with open(os.path.join(d, synth.FILES[key]), "rb") as f:
            while True:
                blk = f.read(1 << 24)
                if not blk:
                    break
                h.update(blk)
    return h.hexdigest()


def main():
    tdir = synth.generate(synth.default_table_dir("test"), "test")
    os.environ.pop("RELXILL_NUM_RZONES", None)
    ref = Ref(tdir)
    e = default_grid(600)  # 600 bins keep the fixture small; the grid is arbitrary for every model
    out = {"energy": e, "table_digest": np.array(table_digest(tdir))}
    conv_in = np.exp(-0.5 * ((np.log(0.5 * (e[1:] + e[:-1])) - np.log(6.4)) / 0.02) ** 2) + 1e-3
    out["conv_input"] = conv_in
    for m in MODELS:
        P = np.vstack([ref.default_params(m)[None, :], sample_params(m, 5, seed=1000 + len(m))])
        if m.startswith("relconv"):
            F = np.stack([ref.eval_conv(m, e, p, conv_in) for p in P])
        else:
            F = ref.eval_batch(m, e, P)
        out[f"{m}_params"] = P
        out[f"{m}_flux"] = F
        print(m, P.shape, F.shape, float(F.sum()))
    # 50-zone variants of the lamp-post models (metric config)
    ref.set_num_zones(50)
    for m in ("relxilllp", "relxilllpCp"):
        P = sample_params(m, 3, seed=77)
        out[f"{m}_z50_params"] = P
        out[f"{m}_z50_flux"] = ref.eval_batch(m, e, P)
    ref.set_num_zones(None)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz"), **out)


def main_nsco():
    """xillverNS / relxillNS / xillverCO / relxillCO on the synthetic xillverNS-2.fits and xillverCO.fits"""
    tdir = synth.generate(synth.default_table_dir("test"), "test")
    os.environ.pop("RELXILL_NUM_RZONES", None)
    ref = Ref(tdir)
    e = default_grid(600)
    out = {"energy": e, "table_digest": np.array(table_digest(tdir, ("rel", "xillns", "xillco")))}
    for m in NSCO_MODELS:
        P = np.vstack([ref.default_params(m)[None, :], sample_params(m, 5, seed=1000 + len(m))])
        F = ref.eval_batch(m, e, P)
        out[f"{m}_params"] = P
        out[f"{m}_flux"] = F
        print(m, P.shape, F.shape, float(F.sum()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v2_nsco.npz"), **out)


def main_configs():
    """golden_v3_configs.npz: the BASELINE.json configurations 2-5 (picked rows of the full batches the GPU tests run),
    evaluated by the unmodified reference."""
    tdir = synth.generate(synth.default_table_dir("test"), "test")
    os.environ.pop("RELXILL_NUM_RZONES", None)
    ref = Ref(tdir)
    e = default_grid(600)
    out = {"energy": e, "table_digest": np.array(table_digest(tdir))}
    cases = config_cases()
    # config 5: returning-radiation sweep over spin x height x inclination (the grid of tests/test_gpu_parity.py)
    base = ref.default_params("relxilllp")
    rows = []
    for a in np.linspace(0.0, 0.998, 8):
        for h in np.geomspace(2.0, 100.0, 8):
            for inc in np.linspace(5.0, 80.0, 4):
                p = base.copy()
                p[0], p[2], p[3], p[12] = h, a, inc, 1
                rows.append(p)
    cases.append(("cfg5_relxilllp", "relxilllp", None, np.array(rows), [0, 37, 101, 200, 255]))
    for key, model, zones, P, pick in cases:
        ref.set_num_zones(zones)
        F = ref.eval_batch(model, e, P[pick])
        out[f"{key}_rows"] = np.array(pick)
        out[f"{key}_params"] = P[pick]
        out[f"{key}_flux"] = F
        print(key, model, zones, F.shape, float(F.sum()))
    ref.set_num_zones(None)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v3_configs.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "configs":
        main_configs()
    elif len(sys.argv) > 1 and sys.argv[1] == "nsco":
        main_nsco()
    else:
        main()
