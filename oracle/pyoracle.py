"""ctypes access to oracle/liboracle.so (our CPU restatement of the hot path).

TEST INFRASTRUCTURE ONLY — see oracle/relxill_oracle.h.  Same Python surface as
oracle/pyref.Ref so tests can swap one for the other.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORC_SO = os.path.join(HERE, "liboracle.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build():
    subprocess.run(["make", "-C", HERE, "restatement"], check=True, stdout=subprocess.DEVNULL)


class Oracle:
    def __init__(self, table_dir: str, num_zones: int | None = None):
        if not os.path.exists(ORC_SO):
            build()
        self.lib = C.CDLL(ORC_SO)
        L = self.lib
        L.orc_init.argtypes = [C.c_char_p]
        L.orc_set_num_zones_env.argtypes = [C.c_int]
        L.orc_num_params.argtypes = [C.c_char_p]
        L.orc_default_params.argtypes = [C.c_char_p, _dp]
        L.orc_eval_model.argtypes = [C.c_char_p, _dp, C.c_int, _dp, _dp]
        L.orc_syspar.argtypes = [C.c_char_p, _dp] + [_dp] * 9
        L.orc_relbase.argtypes = [C.c_char_p, _dp, _dp, C.c_int, _dp]
        L.orc_relxill_stages.argtypes = [C.c_char_p, _dp] + [_dp] * 8 + [_ip, _ip, _dp, _dp]
        L.orc_conv_grid.argtypes = [_dp]
        L.orc_gshift_fluxboost.argtypes = [C.c_double] * 3
        L.orc_gshift_fluxboost.restype = C.c_double
        L.orc_lin2d_float.argtypes = [C.c_double, C.c_double, C.c_float, C.c_float, C.c_float, C.c_float]
        L.orc_lin2d_float.restype = C.c_double
        L.orc_rebin.argtypes = [_dp, _dp, C.c_int, _dp, _dp, C.c_int]
        L.orc_fft_conv.argtypes = [_dp, _dp, _dp]
        L.orc_nthcomp.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_double, _dp]
        L.orc_kerr_rms.argtypes = [C.c_double]
        L.orc_kerr_rms.restype = C.c_double
        L.orc_init(table_dir.encode())
        self.set_num_zones(num_zones)

    def set_num_zones(self, n):
        self.lib.orc_set_num_zones_env(int(n) if n else 0)

    def num_params(self, model):
        return self.lib.orc_num_params(model.encode())

    def default_params(self, model):
        out = np.zeros(64)
        n = self.lib.orc_default_params(model.encode(), out)
        if n < 0:
            raise KeyError(model)
        return out[:n].copy()

    def eval(self, model, energy, par):
        energy = np.ascontiguousarray(energy, np.float64)
        flux = np.zeros(energy.size - 1)
        rc = self.lib.orc_eval_model(model.encode(), energy, energy.size - 1, np.ascontiguousarray(par, np.float64), flux)
        if rc:
            raise RuntimeError(f"oracle evaluation of {model} failed (rc={rc})")
        return flux

    def eval_conv(self, model, energy, par, flux_in):
        energy = np.ascontiguousarray(energy, np.float64)
        flux = np.array(flux_in, np.float64)
        rc = self.lib.orc_eval_model(model.encode(), energy, energy.size - 1, np.ascontiguousarray(par, np.float64), flux)
        if rc:
            raise RuntimeError(f"oracle evaluation of {model} failed (rc={rc})")
        return flux

    def eval_batch(self, model, energy, params):
        params = np.atleast_2d(np.asarray(params, np.float64))
        return np.stack([self.eval(model, energy, p) for p in params])

    def syspar(self, model, par, nr=1000, ng=40):
        a = {k: np.zeros(nr) for k in ("re", "gmin", "gmax", "emis", "del_emit", "del_inc")}
        trff = np.zeros(nr * ng * 2)
        cosne = np.zeros(nr * ng * 2)
        frac = np.zeros(5)
        rc = self.lib.orc_syspar(model.encode(), np.ascontiguousarray(par, np.float64), a["re"], a["gmin"], a["gmax"],
                                 a["emis"], a["del_emit"], a["del_inc"], trff, cosne, frac)
        if rc:
            raise RuntimeError(f"syspar failed rc={rc}")
        a["trff"] = trff.reshape(nr, ng, 2)
        a["cosne"] = cosne.reshape(nr, ng, 2)
        a["frac"] = frac
        return a

    def relbase(self, model, par, ener):
        ener = np.ascontiguousarray(ener, np.float64)
        flux = np.zeros(ener.size - 1)
        rc = self.lib.orc_relbase(model.encode(), np.ascontiguousarray(par, np.float64), ener, ener.size - 1, flux)
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
        nz = self.lib.orc_relxill_stages(model.encode(), np.ascontiguousarray(par, np.float64), zone, zpar, corr,
                                         normch, emis2, relflux, dist, xill, nex, ni, conv, total)
        if nz <= 0:
            raise RuntimeError(f"stages failed rc={nz}")
        nex, ni = int(nex[0]), int(ni[0])
        zp = zpar[: nz * 4].reshape(nz, 4)
        cr = corr[: nz * 2].reshape(nz, 2)
        return dict(nz=nz, zone=zone[: nz + 1], lxi=zp[:, 0], dens=zp[:, 1], ect=zp[:, 2], eshift=zp[:, 3],
                    corr_flux=cr[:, 0], corr_gshift=cr[:, 1], normch=normch[:nz], emis2=emis2,
                    relflux=relflux[: nz * 4096].reshape(nz, 4096), dist=dist[: nz * ni].reshape(nz, ni),
                    xill=xill[: nz * nex].reshape(nz, nex), conv=conv, total=total)

    def conv_grid(self):
        e = np.zeros(4097)
        self.lib.orc_conv_grid(e)
        return e

    def rebin(self, ener, ener0, flu0):
        ener = np.ascontiguousarray(ener, np.float64)
        out = np.zeros(ener.size - 1)
        self.lib.orc_rebin(ener, out, ener.size - 1, np.ascontiguousarray(ener0, np.float64),
                           np.ascontiguousarray(flu0, np.float64), len(flu0))
        return out

    def fft_conv(self, fxill, frel):
        out = np.zeros(4096)
        self.lib.orc_fft_conv(np.ascontiguousarray(fxill, np.float64), np.ascontiguousarray(frel, np.float64), out)
        return out

    def nthcomp(self, ener, gamma, kte, z):
        ener = np.ascontiguousarray(ener, np.float64)
        out = np.zeros(ener.size - 1)
        self.lib.orc_nthcomp(ener, ener.size - 1, gamma, kte, z, out)
        return out
