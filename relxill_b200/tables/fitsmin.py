"""Minimal FITS binary-table writer/reader (numpy only).

There is no astropy/cfitsio in this image, so the synthetic relxill tables are
written by hand.  Only what the relxill table layouts need is supported:
an empty primary HDU followed by BINTABLE extensions whose columns are
fixed-width vectors of E (float32), D (float64), J (int32) or A (char).

Layout facts relied on (FITS standard 4.0): 2880-byte blocks, 80-char header
cards, big-endian row-major table data, PCOUNT=0.
"""
from __future__ import annotations

import numpy as np

BLOCK = 2880

_FORM = {"E": (">f4", 4), "D": (">f8", 8), "J": (">i4", 4), "A": ("S1", 1)}


def _card(key: str, value, comment: str = "") -> bytes:
    if isinstance(value, bool):
        v = f"{'T' if value else 'F':>20}"
    elif isinstance(value, (int, np.integer)):
        v = f"{int(value):>20d}"
    elif isinstance(value, float):
        v = f"{value:>20.12E}"
    else:
        s = "'" + str(value).replace("'", "''").ljust(8) + "'"
        v = f"{s:<20}"
    card = f"{key:<8}= {v}"
    if comment:
        card += f" / {comment}"
    return card[:80].ljust(80).encode("ascii")


def _pad(buf: bytes, fill: bytes) -> bytes:
    rem = (-len(buf)) % BLOCK
    return buf + fill * rem


def _header(cards) -> bytes:
    return _pad(b"".join(cards) + b"END".ljust(80), b" ")


class Column:
    """One table column: `data` has shape (nrows,) or (nrows, width)."""

    def __init__(self, name: str, code: str, data):
        self.name = name
        self.code = code
        if code == "A":
            strs = [str(s) for s in data]
            width = max(1, max(len(s) for s in strs))
            arr = np.zeros((len(strs), width), dtype="S1")
            arr[:] = b" "
            for i, s in enumerate(strs):
                arr[i, : len(s)] = np.frombuffer(s.encode("ascii"), dtype="S1")
            self.data = arr
        else:
            arr = np.asarray(data)
            if arr.ndim == 1:
                arr = arr[:, None]
            self.data = np.ascontiguousarray(arr.astype(_FORM[code][0]))
        self.width = self.data.shape[1]
        self.nbytes = self.width * _FORM[code][1]


class FitsWriter:
    """Streams HDUs to disk so multi-GB tables never sit in memory twice."""

    def __init__(self, path: str):
        self.f = open(path, "wb")
        cards = [
            _card("SIMPLE", True),
            _card("BITPIX", 8),
            _card("NAXIS", 0),
            _card("EXTEND", True),
        ]
        self.f.write(_header(cards))

    def _table_header(self, extname, cols, nrows):
        row = sum(c.nbytes for c in cols)
        cards = [
            _card("XTENSION", "BINTABLE"),
            _card("BITPIX", 8),
            _card("NAXIS", 2),
            _card("NAXIS1", row),
            _card("NAXIS2", nrows),
            _card("PCOUNT", 0),
            _card("GCOUNT", 1),
            _card("TFIELDS", len(cols)),
        ]
        for i, c in enumerate(cols, 1):
            cards.append(_card(f"TTYPE{i}", c.name))
            cards.append(_card(f"TFORM{i}", f"{c.width}{c.code}"))
        cards.append(_card("EXTNAME", extname))
        return _header(cards), row

    def add_table(self, extname: str, cols):
        nrows = cols[0].data.shape[0]
        for c in cols:
            assert c.data.shape[0] == nrows, (extname, c.name)
        hdr, row = self._table_header(extname, cols, nrows)
        self.f.write(hdr)
        rec = np.zeros((nrows, row), dtype=np.uint8)
        off = 0
        for c in cols:
            rec[:, off : off + c.nbytes] = c.data.view(np.uint8).reshape(nrows, c.nbytes)
            off += c.nbytes
        self.f.write(rec.tobytes())
        self.f.write(b"\0" * ((-rec.size) % BLOCK))

    def begin_stream_table(self, extname: str, col_specs, nrows: int):
        """col_specs: list of (name, code, width).  Rows are appended with
        `stream_rows` as already-packed big-endian uint8 records."""
        class _S:  # tiny stand-in carrying the fields _table_header reads
            pass
        cols = []
        for name, code, width in col_specs:
            s = _S()
            s.name, s.code, s.width = name, code, width
            s.nbytes = width * _FORM[code][1]
            cols.append(s)
        hdr, row = self._table_header(extname, cols, nrows)
        self.f.write(hdr)
        self._stream_left = nrows * row
        self._stream_row = row

    def stream_rows(self, rec: np.ndarray):
        rec = np.ascontiguousarray(rec)
        b = rec.view(np.uint8).reshape(-1)
        assert b.size % self._stream_row == 0
        self.f.write(b.tobytes())
        self._stream_left -= b.size
        if self._stream_left == 0:
            total = self.f.tell()
            self.f.write(b"\0" * ((-total) % BLOCK))

    def close(self):
        self.f.close()


def read_tables(path: str):
    """Reads every BINTABLE back: {hdu_number(1-based): (extname, {col: array})}.
    Used by the tests to check writer <-> C reader consistency."""
    out = {}
    with open(path, "rb") as f:
        raw = np.memmap(path, dtype=np.uint8, mode="r")
        pos = 0
        hdu = 0
        n = raw.size
        while pos < n:
            hdr = {}
            order = []
            while True:
                blk = bytes(raw[pos : pos + BLOCK])
                pos += BLOCK
                done = False
                for i in range(0, BLOCK, 80):
                    c = blk[i : i + 80].decode("ascii")
                    k = c[:8].strip()
                    if k == "END":
                        done = True
                        break
                    if c[8:10] == "= ":
                        v = c[10:].split(" / ")[0].strip()
                        if v.startswith("'"):
                            v = v.strip("'").strip()
                        hdr[k] = v
                        order.append(k)
                if done:
                    break
            hdu += 1
            naxis = int(hdr.get("NAXIS", 0))
            size = 0
            if naxis > 0:
                size = 1
                for a in range(1, naxis + 1):
                    size *= int(hdr[f"NAXIS{a}"])
            if hdr.get("XTENSION", "") == "BINTABLE":
                nrows = int(hdr["NAXIS2"])
                row = int(hdr["NAXIS1"])
                rec = np.asarray(raw[pos : pos + size]).reshape(nrows, row)
                cols = {}
                off = 0
                for i in range(1, int(hdr["TFIELDS"]) + 1):
                    form = hdr[f"TFORM{i}"]
                    code = form[-1]
                    width = int(form[:-1] or 1)
                    dt, sz = _FORM[code]
                    chunk = np.ascontiguousarray(rec[:, off : off + width * sz])
                    if code == "A":
                        cols[hdr[f"TTYPE{i}"]] = [bytes(r).decode("ascii").strip() for r in chunk]
                    else:
                        cols[hdr[f"TTYPE{i}"]] = chunk.view(dt).reshape(nrows, width).astype(dt[1:])
                    off += width * sz
                out[hdu] = (hdr.get("EXTNAME", ""), cols)
            pos += size + ((-size) % BLOCK)
    return out
