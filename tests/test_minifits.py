"""The C FITS reader of the table loader (relxill_b200/csrc/minifits.h) on layouts the synthetic tables do not have:
mixed column types (I, B, K, L next to E / D / J / A: every later column's byte offset depends on their sizes), scaled
columns, a file written by cfitsio itself (the reference's regression fixtures, when the tree is present), and an
unknown TFORM code, which must make every read fail instead of shifting the columns silently."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

HARNESS = r'''
#include <stdio.h>
#include <string.h>
#include "minifits.h"
/* usage: harness file extname col kind(d|i|s) row -> prints the cell (first element) or "ERR <rc>" */
int main(int argc, char **argv) {
  if (argc < 6) return 2;
  mf_file *f = mf_open(argv[1]);
  if (!f) { printf("ERR open\n"); return 0; }
  int ih = mf_find_hdu(f, argv[2]);
  if (!ih) { printf("ERR hdu\n"); return 0; }
  const mf_hdu *h = &f->hdus[ih - 1];
  int ic = mf_find_col(h, argv[3]);
  if (!ic) { printf("ERR col\n"); return 0; }
  long row = atol(argv[5]);
  if (argv[4][0] == 's') {
    char buf[128];
    int rc = mf_read_str(h, ic, row, buf, sizeof(buf));
    if (rc) printf("ERR %d\n", rc); else printf("%s\n", buf);
  } else {
    double v = 0;
    int rc = mf_read(h, ic, row, 1, 1, 'd', &v);
    if (rc) printf("ERR %d\n", rc); else printf("%.17g\n", v);
  }
  mf_close(f);
  return 0;
}
'''


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp("mf")
    src = d / "h.c"
    src.write_text(HARNESS)
    exe = d / "h"
    subprocess.run(["gcc", "-O1", "-I", os.path.join(ROOT, "relxill_b200", "csrc"), str(src), "-o", str(exe), "-lm"], check=True)

    def run(path, ext, col, kind, row):
        return subprocess.run([str(exe), str(path), ext, col, kind, str(row)], capture_output=True, text=True, check=True).stdout.strip()
    return run


def _card(k, v):
    return f"{k:<8}= {v:>20}".ljust(80).encode()


def _write_table(path, extname, cols, rows, extra_cards=()):
    """cols: [(name, tform, struct format)], rows: list of tuples; big-endian, one element per cell (strings padded)."""
    fmt = ">" + "".join(c[2] for c in cols)
    rowbytes = struct.calcsize(fmt)
    hdr0 = _card("SIMPLE", "T") + _card("BITPIX", 8) + _card("NAXIS", 0) + _card("EXTEND", "T") + b"END".ljust(80)
    hdr0 += b" " * ((-len(hdr0)) % 2880)
    cards = [_card("XTENSION", "'BINTABLE'"), _card("BITPIX", 8), _card("NAXIS", 2), _card("NAXIS1", rowbytes), _card("NAXIS2", len(rows)),
             _card("PCOUNT", 0), _card("GCOUNT", 1), _card("TFIELDS", len(cols))]
    for i, (name, tform, _) in enumerate(cols, 1):
        cards += [_card(f"TTYPE{i}", f"'{name:<8}'"), _card(f"TFORM{i}", f"'{tform:<8}'")]
    cards += list(extra_cards) + [_card("EXTNAME", f"'{extname:<8}'")]
    hdr = b"".join(cards) + b"END".ljust(80)
    hdr += b" " * ((-len(hdr)) % 2880)
    data = b"".join(struct.pack(fmt, *r) for r in rows)
    data += b"\0" * ((-len(data)) % 2880)
    with open(path, "wb") as f:
        f.write(hdr0 + hdr + data)


def test_mixed_column_types_and_offsets(harness, tmp_path):
    p = tmp_path / "mixed.fits"
    cols = [("NAME", "12A", "12s"), ("SHORT", "1I", "h"), ("BYTE", "1B", "B"), ("FLAG", "1L", "c"), ("LONG", "1K", "q"),
            ("INT", "1J", "i"), ("FLT", "1E", "f"), ("DBL", "1D", "d")]
    rows = [(b"rel_table   ", -7, 200, b"T", 2 ** 40 + 3, 123456, 1.5, 2.25), (b"xillver-a   ", 9, 1, b"F", -5, -1, -0.5, 1e-300)]
    _write_table(p, "MIXED", cols, rows)
    assert harness(p, "MIXED", "NAME", "s", 2) == "xillver-a"
    assert float(harness(p, "MIXED", "SHORT", "d", 1)) == -7
    assert float(harness(p, "MIXED", "BYTE", "d", 1)) == 200
    assert float(harness(p, "MIXED", "LONG", "d", 1)) == 2 ** 40 + 3
    assert float(harness(p, "MIXED", "INT", "d", 2)) == -1            # behind I, B, L, K: offsets from the real sizes
    assert float(harness(p, "MIXED", "FLT", "d", 1)) == 1.5
    assert float(harness(p, "MIXED", "DBL", "d", 2)) == 1e-300


def test_scaled_column(harness, tmp_path):
    p = tmp_path / "scaled.fits"
    _write_table(p, "SC", [("RAW", "1I", "h")], [(100,), (-3,)], extra_cards=[_card("TSCAL1", 0.5), _card("TZERO1", 10.0)])
    assert float(harness(p, "SC", "RAW", "d", 1)) == 60.0
    assert float(harness(p, "SC", "RAW", "d", 2)) == 8.5


def test_unknown_tform_fails_loudly(harness, tmp_path):
    p = tmp_path / "bad.fits"
    _write_table(p, "BAD", [("A", "1J", "i"), ("WEIRD", "1Z", "i"), ("B", "1D", "d")], [(1, 2, 3.0)])
    assert harness(p, "BAD", "A", "d", 1).startswith("ERR")     # not "1": the whole table is rejected
    assert harness(p, "BAD", "B", "d", 1).startswith("ERR")


def test_row_width_mismatch_fails(harness, tmp_path):
    p = tmp_path / "short.fits"
    _write_table(p, "W", [("A", "1J", "i"), ("B", "1D", "d")], [(1, 2.0)])
    raw = bytearray(open(p, "rb").read())
    i = raw.find(b"NAXIS1  =")
    raw[i:i + 80] = _card("NAXIS1", 16)          # the columns add up to 12
    open(p, "wb").write(bytes(raw))
    assert harness(p, "W", "B", "d", 1).startswith("ERR")


def test_file_written_by_cfitsio(harness):
    f = "/root/reference/test/refdata/refdata_localModels/relxilllp/relxilllp_defparam_refdat_0001.fits"
    if not os.path.exists(f):
        pytest.skip("reference tree not present")
    # MODEL: freeze J | max D | min D | name 34A | value D  (test/refdata/fits_model_struct.sl:19-40)
    assert harness(f, "MODEL", "name", "s", 2) == "relxilllp(1).h"
    assert float(harness(f, "MODEL", "value", "d", 2)) == 6.0
    assert float(harness(f, "MODEL", "max", "d", 2)) == 100.0
    assert abs(float(harness(f, "DATA", "bin_lo", "d", 1)) - 0.1) < 1e-15
    d = np.load(os.path.join(HERE, "golden", "refdata_v1.npz"))
    i = list(d["files"]).index("test/refdata/refdata_localModels/relxilllp/relxilllp_defparam_refdat_0001.fits")
    assert float(harness(f, "DATA", "value", "d", 7)) == d["spectra"][i][6]
