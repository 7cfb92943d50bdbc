"""Turns the reference's own regression fixtures (test/refdata/refdata_localModels/<model>/*.fits: a MODEL extension with
the ISIS parameter list and a DATA extension with the spectrum the published tables gave) into one small file the tests
can read where /root/reference does not exist.

    python tests/golden/make_refdata_fixture.py [/root/reference] -> tests/golden/refdata_v1.npz

Nothing is evaluated here: the fixtures need the PUBLISHED FITS tables, which are not available offline.  The file
serves (i) the parameter-layout check of tests/test_refdata.py and (ii) the spectrum check with the reference's
criterion (test/refdata/test_refdata_relxill.sl:165-171) on the day real tables are found in RELXILL_TABLE_PATH."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from relxill_b200.tables.fitsmin import read_tables  # noqa: E402


def read_fixture(path):
    t = {name: cols for name, cols in read_tables(path).values()}
    mo, da = t["MODEL"], t["DATA"]
    names = [s for s in mo["name"]]
    model = names[0].split("(")[0]
    short = [s.split(").", 1)[1] for s in names]
    return dict(model=model, names=short, values=np.asarray(mo["value"], float).ravel(), bin_lo=np.asarray(da["bin_lo"], float).ravel(),
                bin_hi=np.asarray(da["bin_hi"], float).ravel(), value=np.asarray(da["value"], float).ravel())


def collect(ref_root):
    files = sorted(glob.glob(os.path.join(ref_root, "test/refdata/refdata_localModels/*/*.fits")))
    fx = [read_fixture(f) for f in files]
    npar = max(len(f["names"]) for f in fx)
    vals = np.full((len(fx), npar), np.nan)
    for i, f in enumerate(fx):
        vals[i, : len(f["values"])] = f["values"]
    lo, hi = fx[0]["bin_lo"], fx[0]["bin_hi"]
    assert all(np.array_equal(f["bin_lo"], lo) and np.array_equal(f["bin_hi"], hi) for f in fx), "fixtures share one grid"
    return dict(files=np.array([os.path.relpath(f, ref_root) for f in files]), models=np.array([f["model"] for f in fx]),
                names=np.array([",".join(f["names"]) for f in fx]), values=vals, bin_lo=lo, bin_hi=hi,
                spectra=np.stack([f["value"] for f in fx]))


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    d = collect(ref)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refdata_v1.npz")
    np.savez_compressed(out, **d)
    print(out, len(d["files"]), "fixtures,", os.path.getsize(out), "bytes; models:", sorted(set(d["models"])))
