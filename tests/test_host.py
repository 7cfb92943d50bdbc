"""CPU tests of the host side: the C-ABI library loads and exports what include/*.h declares, the model
database matches the oracle's, the FITS writer/reader pair round-trips, sharding + gather logic under gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "relxill_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    syms = set(re.findall(r"RELXILL_B200_LMOD\((\w+)\)\s*;", txt))
    syms |= set(re.findall(r"\b(relxill_\w+)\s*\(", txt))
    syms.discard("relxill_b200_batch")
    return syms


def test_library_exports_every_declared_symbol():
    from relxill_b200 import _lib, build
    build.build()
    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for s in declared:
        assert hasattr(L, s), f"{s} declared in include/relxill_b200.h but not exported"
    assert declared == set(_lib.ABI_SYMBOLS)


def test_public_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/relxill_b200.h must compile as C99 (what a cgo / ctypes / Fortran-wrapper caller
    sees) and as C++ (what XSPEC's generated wrapper sees), warnings as errors."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "relxill_b200.h"\nint main(void) { return 0; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)], check=True)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)], check=True)


def test_model_layout_matches_oracle(oracle):
    import relxill_b200 as rx
    for m in rx.PARAM_NAMES:
        assert rx.num_params(m) == oracle.num_params(m) == len(rx.PARAM_NAMES[m])
        np.testing.assert_array_equal(rx.default_params(m), oracle.default_params(m))
    with pytest.raises(rx.ModelNotFound):
        rx.num_params("relxillXX")


def test_model_table_matches_the_reference_lmodel_dat():
    """The drop-in boundary: for every model on the hot path the C symbol XSPEC resolves, the parameter order (with the
    $switch entries) and the defaults are the ones of the reference's lmodel_relxill_public.dat / _devel.dat
    (fixture tests/golden/lmodel_layout.json, made by tests/golden/make_lmodel_layout.py from the reference)."""
    import json
    import relxill_b200 as rx
    from relxill_b200 import _lib
    layout = json.load(open(os.path.join(ROOT, "tests", "golden", "lmodel_layout.json")))
    out_of_scope = {"relxillBB", "relxilllpAlpha"}                 # SURVEY.md §2.2; INTEGRATION.md §1
    assert set(layout) - out_of_scope == set(rx.PARAM_NAMES) == set(_lib.LMOD_SYMBOLS)
    for m, ref in layout.items():
        if m in out_of_scope:
            assert ref["symbol"] not in _lib.ABI_SYMBOLS
            continue
        assert _lib.LMOD_SYMBOLS[m] == ref["symbol"], m
        assert [p["name"] for p in ref["params"]] == rx.PARAM_NAMES[m], m
        np.testing.assert_array_equal(rx.default_params(m), [p["default"] for p in ref["params"]], err_msg=m)
        assert rx.num_params(m) == len(ref["params"])


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import relxill_b200 as rx
    with pytest.raises(rx.ModelEvalFailed):
        rx.batch_eval("relline", rx.default_energy_grid(), rx.default_params("relline"))


def test_fits_roundtrip(tmp_path):
    from relxill_b200.tables.fitsmin import Column, FitsWriter, read_tables
    p = str(tmp_path / "t.fits")
    w = FitsWriter(p)
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    d = np.linspace(0, 1, 3)
    w.add_table("X", [Column("name", "A", ["ab", "c", "def"]), Column("a", "E", a), Column("d", "D", d),
                      Column("n", "J", np.array([1, 2, 3]))])
    w.close()
    t = read_tables(p)
    ext, cols = t[2]
    assert ext == "X" and cols["name"] == ["ab", "c", "def"]
    np.testing.assert_array_equal(cols["a"], a)
    np.testing.assert_array_equal(cols["d"][:, 0], d)
    np.testing.assert_array_equal(cols["n"][:, 0], [1, 2, 3])


def test_synthetic_tables_layout(table_dir):
    from relxill_b200.tables import synth
    from relxill_b200.tables.fitsmin import read_tables
    t = read_tables(os.path.join(table_dir, synth.FILES["rrad"]))
    ext, cols = t[3]
    assert ext == "FRAC01" and cols["frac_g"].shape == (50, 1000)
    np.testing.assert_allclose(cols["frac_g"].reshape(50, 50, 20).sum(axis=2), 1.0, rtol=1e-12)


def test_shard_bounds():
    from relxill_b200.dist import shard_bounds, shard_indices
    for n in (1, 7, 4096, 65536 + 3):
        for w in (1, 2, 3, 8):
            cover = np.concatenate([np.arange(*shard_bounds(n, w, r)) for r in range(w)])
            np.testing.assert_array_equal(cover, np.arange(n))
            inter = np.sort(np.concatenate([shard_indices(n, w, r, interleave=True) for r in range(w)]))
            np.testing.assert_array_equal(inter, np.arange(n))


WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], 'tests'))
import numpy as np, torch, torch.distributed as dist
from relxill_b200.dist import sharded_eval
from relxill_b200.tables import synth
from oracle.pyoracle import Oracle
from common import default_grid, sample_params
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
tdir = synth.generate(synth.default_table_dir('test'), 'test')
orc = Oracle(tdir)
e = default_grid(200)
P = sample_params('relline', 5, seed=3)          # 5 vectors over 2 ranks: ragged shards
def evaluate(p_shard):                            # stand-in for the CUDA evaluation (no GPU in this container)
    return torch.from_numpy(np.stack([orc.eval('relline', e, p) for p in p_shard]) if len(p_shard) else np.zeros((0, 200)))
for inter in (False, True):
    full = sharded_eval(evaluate, P, e.size - 1, interleave=inter)
    want = np.stack([orc.eval('relline', e, p) for p in P])
    assert full.shape == (5, 200) and np.array_equal(full.numpy(), want), (rank, inter)
dist.barrier()
if rank == 0: print('GLOO_OK')
"""


def test_sharded_eval_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0 and "GLOO_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
