"""ctypes access to oracle/_ref/librelxill_ref.so (the unmodified reference + probes).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by relxill_b200/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "librelxill_ref.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available() -> bool:
    return os.path.exists(REF_SO)


class RefLocal:
    """In-process binding.  The reference is full of process-global state (tables, caches, env reads on
    every call), so: one table directory per process, set before the first call."""

    def __init__(self, table_dir: str, num_zones: int | None = None):
        os.environ["RELXILL_TABLE_PATH"] = table_dir
        if num_zones is not None:
            os.environ["RELXILL_NUM_RZONES"] = str(num_zones)
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.ref_eval_model.argtypes = [C.c_char_p, _dp, C.c_int, _dp, _dp]
        L.ref_eval_model.restype = C.c_int
        L.ref_num_params.argtypes = [C.c_char_p]
        L.ref_default_params.argtypes = [C.c_char_p, _dp]
        L.ref_rel_params.argtypes = [C.c_char_p, _dp, _dp, _ip]
        L.ref_syspar.argtypes = [C.c_char_p, _dp] + [_dp] * 9
        L.ref_relbase.argtypes = [C.c_char_p, _dp, _dp, C.c_int, _dp]
        L.ref_relxill_stages.argtypes = [C.c_char_p, _dp] + [_dp] * 8 + [_ip, _ip, _dp, _dp]
        L.ref_conv_grid.argtypes = [_dp]
        L.ref_rebin.argtypes = [_dp, _dp, C.c_int, _dp, _dp, C.c_int]
        L.ref_fft_conv.argtypes = [_dp, _dp, _dp]
        L.ref_nthcomp.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_double, _dp]
        L.ref_kerr_rms.argtypes = [C.c_double]
        L.ref_kerr_rms.restype = C.c_double

    def set_num_zones(self, n):
        if n is None:
            os.environ.pop("RELXILL_NUM_RZONES", None)
        else:
            os.environ["RELXILL_NUM_RZONES"] = str(n)

    def num_params(self, model):
        return self.lib.ref_num_params(model.encode())

    def default_params(self, model):
        out = np.zeros(64)
        n = self.lib.ref_default_params(model.encode(), out)
        if n < 0:
            raise KeyError(model)
        return out[:n].copy()

    def eval(self, model, energy, par):
        energy = np.ascontiguousarray(energy, np.float64)
        par = np.ascontiguousarray(par, np.float64)
        flux = np.zeros(energy.size - 1)
        rc = self.lib.ref_eval_model(model.encode(), energy, energy.size - 1, par, flux)
        if rc:
            raise RuntimeError(f"reference evaluation of {model} failed (rc={rc})")
        return flux

    def eval_conv(self, model, energy, par, flux_in):
        energy = np.ascontiguousarray(energy, np.float64)
        flux = np.array(flux_in, np.float64)
        rc = self.lib.ref_eval_model(model.encode(), energy, energy.size - 1,
                                     np.ascontiguousarray(par, np.float64), flux)
        if rc:
            raise RuntimeError(f"reference evaluation of {model} failed (rc={rc})")
        return flux

    def eval_batch(self, model, energy, params):
        params = np.atleast_2d(np.asarray(params, np.float64))
        return np.stack([self.eval(model, energy, p) for p in params])

    def rel_params(self, model, par):
        rel = np.zeros(12)
        irel = np.zeros(6, np.int32)
        rc = self.lib.ref_rel_params(model.encode(), np.ascontiguousarray(par, np.float64), rel, irel)
        if rc:
            raise RuntimeError(f"rel_params failed rc={rc}")
        keys = "a incl emis1 emis2 rbr rin rout lineE z height gamma beta".split()
        ikeys = "model_type emis_type limb num_zones return_rad ion_grad_type".split()
        d = dict(zip(keys, rel))
        d.update(dict(zip(ikeys, (int(v) for v in irel))))
        return d

    def syspar(self, model, par, nr=1000, ng=40):
        a = {k: np.zeros(nr) for k in ("re", "gmin", "gmax", "emis", "del_emit", "del_inc")}
        trff = np.zeros(nr * ng * 2)
        cosne = np.zeros(nr * ng * 2)
        frac = np.zeros(5)
        rc = self.lib.ref_syspar(model.encode(), np.ascontiguousarray(par, np.float64), a["re"], a["gmin"],
                                 a["gmax"], a["emis"], a["del_emit"], a["del_inc"], trff, cosne, frac)
        if rc:
            raise RuntimeError(f"syspar failed rc={rc}")
        a["trff"] = trff.reshape(nr, ng, 2)
        a["cosne"] = cosne.reshape(nr, ng, 2)
        a["frac"] = frac
        return a

    def relbase(self, model, par, ener):
        ener = np.ascontiguousarray(ener, np.float64)
        flux = np.zeros(ener.size - 1)
        rc = self.lib.ref_relbase(model.encode(), np.ascontiguousarray(par, np.float64), ener, ener.size - 1, flux)
        if rc:
            raise RuntimeError(f"relbase failed rc={rc}")
        return flux

    def stages(self, model, par, nzmax=50, nemax=6000, nimax=16):
        zone = np.zeros(nzmax + 1)
        zpar = np.zeros(nzmax * 4)
        corr = np.zeros(nzmax * 2)
        normch = np.zeros(nzmax)
        emis2 = np.zeros(1000)
        relflux = np.zeros(nzmax * 4096)
        dist = np.zeros(nzmax * nimax)
        xill = np.zeros(nzmax * nemax)
        nex = np.zeros(1, np.int32)
        ni = np.zeros(1, np.int32)
        conv = np.zeros(4096)
        total = np.zeros(4096)
        nz = self.lib.ref_relxill_stages(model.encode(), np.ascontiguousarray(par, np.float64), zone, zpar, corr,
                                         normch, emis2, relflux, dist, xill, nex, ni, conv, total)
        if nz <= 0:
            raise RuntimeError(f"stages failed rc={nz}")
        nex, ni = int(nex[0]), int(ni[0])
        return dict(
            nz=nz, zone=zone[: nz + 1], lxi=zpar[: nz * 4].reshape(nz, 4)[:, 0], dens=zpar[: nz * 4].reshape(nz, 4)[:, 1],
            ect=zpar[: nz * 4].reshape(nz, 4)[:, 2], eshift=zpar[: nz * 4].reshape(nz, 4)[:, 3],
            corr_flux=corr[: nz * 2].reshape(nz, 2)[:, 0], corr_gshift=corr[: nz * 2].reshape(nz, 2)[:, 1],
            normch=normch[:nz], emis2=emis2, relflux=relflux[: nz * 4096].reshape(nz, 4096),
            dist=dist[: nz * ni].reshape(nz, ni), xill=xill[: nz * nex].reshape(nz, nex), conv=conv, total=total,
        )

    def conv_grid(self):
        e = np.zeros(4097)
        self.lib.ref_conv_grid(e)
        return e

    def rebin(self, ener, ener0, flu0):
        ener = np.ascontiguousarray(ener, np.float64)
        out = np.zeros(ener.size - 1)
        self.lib.ref_rebin(ener, out, ener.size - 1, np.ascontiguousarray(ener0, np.float64),
                           np.ascontiguousarray(flu0, np.float64), len(flu0))
        return out

    def fft_conv(self, fxill, frel):
        out = np.zeros(4096)
        self.lib.ref_fft_conv(np.ascontiguousarray(fxill, np.float64), np.ascontiguousarray(frel, np.float64), out)
        return out

    def nthcomp(self, ener, gamma, kte, z):
        ener = np.ascontiguousarray(ener, np.float64)
        out = np.zeros(ener.size - 1)
        self.lib.ref_nthcomp(ener, ener.size - 1, gamma, kte, z, out)
        return out


def _worker(conn, table_dir, num_zones):
    ref = RefLocal(table_dir, num_zones)
    while True:
        msg = conn.recv()
        if msg is None:
            break
        name, args = msg
        try:
            conn.send(("ok", getattr(ref, name)(*args)))
        except Exception as exc:  # noqa: BLE001
            conn.send(("err", repr(exc)))


class Ref:
    """Process-isolated binding used by the tests.  The reference keeps linked-list / deque caches in
    file-scope globals and is not robust when many different models and zone counts are evaluated in one
    process (it can segfault after a few dozen mixed evaluations), so each model gets its own worker
    process, recycled every `max_calls` calls.  Results do not depend on the caches (they are
    result-transparent), only the crash behaviour does."""

    def __init__(self, table_dir: str, num_zones: int | None = None, max_calls: int = 40):
        self.table_dir, self.num_zones, self.max_calls = table_dir, num_zones, max_calls
        self._workers = {}

    def set_num_zones(self, n):
        if n != self.num_zones:
            self.close()
        self.num_zones = n

    def _call(self, key, name, *args):
        import multiprocessing as mp
        w = self._workers.get(key)
        if w is not None and (w[2] >= self.max_calls or not w[0].is_alive()):
            self._stop(key)
            w = None
        if w is None:
            ctx = mp.get_context("spawn")
            parent, child = ctx.Pipe()
            proc = ctx.Process(target=_worker, args=(child, self.table_dir, self.num_zones), daemon=True)
            proc.start()
            w = [proc, parent, 0]
            self._workers[key] = w
        w[2] += 1
        w[1].send((name, args))
        if not w[1].poll(600):
            self._stop(key)
            raise RuntimeError("reference worker timed out")
        try:
            tag, val = w[1].recv()
        except EOFError:
            self._stop(key)
            raise RuntimeError(f"reference worker died in {name}{args[:1]}")
        if tag == "err":
            raise RuntimeError(val)
        return val

    def _stop(self, key):
        w = self._workers.pop(key, None)
        if w is None:
            return
        try:
            w[1].send(None)
        except Exception:  # noqa: BLE001
            pass
        w[0].join(2)
        if w[0].is_alive():
            w[0].kill()

    def close(self):
        for k in list(self._workers):
            self._stop(k)

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def num_params(self, model): return self._call(model, "num_params", model)
    def default_params(self, model): return self._call(model, "default_params", model)
    def eval(self, model, energy, par): return self._call(model, "eval", model, energy, par)
    def eval_conv(self, model, energy, par, flux_in): return self._call(model, "eval_conv", model, energy, par, flux_in)
    def eval_batch(self, model, energy, params): return self._call(model, "eval_batch", model, energy, params)
    def rel_params(self, model, par): return self._call(model, "rel_params", model, par)
    def syspar(self, model, par): return self._call(model, "syspar", model, par)
    def relbase(self, model, par, ener): return self._call(model, "relbase", model, par, ener)
    def stages(self, model, par): return self._call(model, "stages", model, par)
    def conv_grid(self): return self._call("_util", "conv_grid")
    def rebin(self, ener, ener0, flu0): return self._call("_util", "rebin", ener, ener0, flu0)
    def fft_conv(self, fxill, frel): return self._call("_util", "fft_conv", fxill, frel)
    def nthcomp(self, ener, gamma, kte, z): return self._call("_util", "nthcomp", ener, gamma, kte, z)


def default_grid(n=3000, emin=0.1, emax=1000.0):
    """DefaultSpec grid of the reference (src/XspecSpectrum.h:143-149): log grid, last edge forced."""
    i = np.arange(n + 1, dtype=np.float64)
    e = np.exp(i / float(n) * (np.log(emax) - np.log(emin)) + np.log(emin))
    e[-1] = emax
    return e
