"""Host-side mirror of the reference's local-model interface on top of the C ABI.

The reference's in-process API is `LocalModel(ModelName).set_par(XPar, v).eval_model(spectrum)`
(reference test/speed/speed_test.cpp:30-44, src/LocalModel.h:62-128); XSPEC calls the generated
`lmod*` C functions (src/create_wrapper_xspec.py:153-162).  Both are mirrored here with the same
names, parameter order and failure behaviour, plus the batched entry point.  All numerical work
happens in librelxill_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib

# parameter names in lmodel.dat order (reference src/modelfiles/lmodel_relxill_public.dat)
PARAM_NAMES = {
    "relline": ["lineE", "Index1", "Index2", "Rbr", "a", "Incl", "Rin", "Rout", "z", "limb"],
    "relconv": ["Index1", "Index2", "Rbr", "a", "Incl", "Rin", "Rout", "limb"],
    "relline_lp": ["lineE", "h", "a", "Incl", "Rin", "Rout", "z", "limb", "gamma", "switch_returnrad"],
    "relconv_lp": ["h", "beta", "a", "Incl", "Rin", "Rout", "limb", "gamma", "switch_returnrad"],
    "relxill": ["Index1", "Index2", "Rbr", "a", "Incl", "Rin", "Rout", "z", "gamma", "logxi", "Afe", "Ecut",
                "refl_frac"],
    "relxilllp": ["h", "beta", "a", "Incl", "Rin", "Rout", "z", "gamma", "logxi", "Afe", "Ecut", "refl_frac",
                  "switch_returnrad", "switch_reflfrac_boost"],
    "relxillCp": ["Incl", "a", "Rin", "Rout", "Rbr", "Index1", "Index2", "z", "gamma", "logxi", "logN", "Afe", "kTe",
                  "refl_frac"],
    "xillver": ["gamma", "Afe", "Ecut", "logxi", "z", "Incl", "refl_frac"],
    "xillverCp": ["gamma", "Afe", "kTe", "logxi", "logN", "z", "Incl", "refl_frac"],
    "relxilllpCp": ["Incl", "a", "Rin", "Rout", "h", "beta", "gamma", "logxi", "logN", "Afe", "kTe", "refl_frac", "z",
                    "iongrad_index", "iongrad_type", "switch_returnrad", "switch_reflfrac_boost"],
    # lmodel_relxill_public.dat:131-153 and lmodel_relxill_devel.dat:1-25
    "xillverNS": ["kTbb", "Afe", "logN", "logxi", "z", "Incl", "refl_frac"],
    "relxillNS": ["Index1", "Index2", "Rbr", "a", "Incl", "Rin", "Rout", "z", "kTbb", "logxi", "Afe", "logN",
                  "refl_frac"],
    "xillverCO": ["gamma", "A_CO", "kTbb", "frac_pl_bb", "Ecut", "z", "Incl", "refl_frac"],
    "relxillCO": ["Index1", "Index2", "Rbr", "a", "Incl", "Rin", "Rout", "z", "gamma", "A_CO", "kTbb", "frac_pl_bb",
                  "Ecut", "refl_frac"],
}


class ModelEvalFailed(RuntimeError):
    """Same role as the reference's ModelEvalFailed (src/LocalModel.h:37-52)."""


class ModelNotFound(KeyError):
    """Same role as the reference's ModelNotFound (src/ModelDatabase.h:33-48)."""


def init(table_dir: str | None = None, device: int = -1) -> None:
    rc = _lib.lib().relxill_b200_init(table_dir.encode() if table_dir else None, device)
    if rc != 0:
        raise RuntimeError("relxill_b200 initialisation failed: " + _lib.last_error())


def shutdown() -> None:
    _lib.lib().relxill_b200_shutdown()


def set_num_zones(n: int | None) -> None:
    """RELXILL_NUM_RZONES of the reference (src/relutility.c:506-544); None/0 = defaults."""
    _lib.lib().relxill_b200_set_num_zones(int(n) if n else 0)


def set_cache(on: bool) -> None:
    """Device-resident state cache (re-use of the previous run's intermediates, include/relxill_b200.h); on by default."""
    _lib.lib().relxill_b200_set_cache(1 if on else 0)


def last_eval_reuse() -> dict:
    """Vectors of the last `batch_eval` / `lmod` call that were recomputed / re-used their relativistic half / re-used
    their whole spectrum (summed over the devices)."""
    out = np.zeros(3, np.int64)
    _lib.lib().relxill_b200_last_eval_reuse(out)
    return dict(recomputed=int(out[0]), reused_rel=int(out[1]), reused_all=int(out[2]))


def init_devices(table_dir: str | None = None, n_devices: int = 0) -> int:
    """One engine per device 0..n_devices-1 (0: all visible) in this process; `batch_eval` then shards over them.
    Returns the number of engines."""
    rc = _lib.lib().relxill_b200_init_devices(table_dir.encode() if table_dir else None, int(n_devices))
    if rc != 0:
        raise RuntimeError("relxill_b200 initialisation failed: " + _lib.last_error())
    return int(_lib.lib().relxill_b200_num_devices())


def num_devices() -> int:
    return int(_lib.lib().relxill_b200_num_devices())


def set_sharding(interleave: bool) -> None:
    """Multi-device split of `batch_eval`: contiguous blocks (default) or round-robin rows (parameter-grid sweeps)."""
    _lib.lib().relxill_b200_set_sharding(1 if interleave else 0)


def set_xill_grid(conv_grid: bool) -> None:
    """Where the per-zone xillver spectra are filed: on the convolution grid (default) or on the table grid
    (include/relxill_b200.h); the results agree to rounding."""
    _lib.lib().relxill_b200_set_xill_grid(1 if conv_grid else 0)


def get_xill_grid() -> bool:
    return bool(_lib.lib().relxill_b200_get_xill_grid())


def num_params(model: str) -> int:
    n = _lib.lib().relxill_b200_num_params(model.encode())
    if n < 0:
        raise ModelNotFound(model)
    return n


def default_params(model: str) -> np.ndarray:
    out = np.zeros(32)
    n = _lib.lib().relxill_b200_default_params(model.encode(), out)
    if n < 0:
        raise ModelNotFound(model)
    return out[:n].copy()


def default_energy_grid(n: int = 3000, emin: float = 0.1, emax: float = 1000.0) -> np.ndarray:
    """DefaultSpec grid (reference src/XspecSpectrum.h:121,143-149)."""
    i = np.arange(n + 1, dtype=np.float64)
    e = np.exp(i / float(n) * (np.log(emax) - np.log(emin)) + np.log(emin))
    e[-1] = emax
    return e


def batch_eval(model: str, energy, params, flux_in=None, return_status: bool = False):
    """N parameter vectors on one energy grid -> flux [N, n_flux] (host arrays in, host arrays out).
    For convolution models pass the input spectra as `flux_in` [N, n_flux]."""
    energy = np.ascontiguousarray(energy, np.float64)
    params = np.ascontiguousarray(np.atleast_2d(params), np.float64)
    npar = num_params(model)
    if params.shape[1] != npar:
        raise ValueError(f"{model} takes {npar} parameters, got {params.shape[1]}")
    n, n_flux = params.shape[0], energy.size - 1
    if flux_in is not None:
        flux = np.ascontiguousarray(np.broadcast_to(np.asarray(flux_in, np.float64), (n, n_flux))).copy()
    else:
        flux = np.zeros((n, n_flux))
    status = np.zeros(n, np.int32)
    rc = _lib.lib().relxill_batch_eval(model.encode(), energy, n_flux, params, n, flux, status)
    if rc != 0:
        raise ModelEvalFailed(f"batched evaluation of {model} failed: {_lib.last_error()}")
    return (flux, status) if return_status else flux


class Batch:
    """A prepared batch: parameters interpreted and resident in HBM; `run` only launches kernels."""

    def __init__(self, model: str, energy, params, keep_intermediates: bool = False):
        """keep_intermediates: also store what only `probe` reads (the fine emission-angle tables)."""
        self.model = model
        self._keep = bool(keep_intermediates)
        self.energy = np.ascontiguousarray(energy, np.float64)
        self.params = np.ascontiguousarray(np.atleast_2d(params), np.float64)
        self.n, self.n_flux = self.params.shape[0], self.energy.size - 1
        if self.params.shape[1] != num_params(model):
            raise ValueError("wrong parameter count")
        self._h = _lib.lib().relxill_b200_prepare(model.encode(), self.energy, self.n_flux, self.params, self.n)
        if not self._h:
            raise ModelEvalFailed(f"prepare({model}) failed: {_lib.last_error()}")

    def run(self, d_flux_ptr: int, stream_ptr: int = 0) -> None:
        if self._keep:
            _lib.lib().relxill_b200_keep_intermediates(1)
        try:
            rc = _lib.lib().relxill_b200_run(self._h, C.c_void_p(d_flux_ptr), C.c_void_p(stream_ptr))
        finally:
            if self._keep:
                _lib.lib().relxill_b200_keep_intermediates(0)
        if rc != 0:
            raise ModelEvalFailed(f"run({self.model}) failed: {_lib.last_error()}")

    def update_params(self, params) -> None:
        """New parameter vectors for the same model and batch size; the next run re-uses what they leave valid."""
        params = np.ascontiguousarray(np.atleast_2d(params), np.float64)
        if params.shape != self.params.shape:
            raise ValueError("update_params: shape must stay the same")
        self.params = params
        if _lib.lib().relxill_b200_update_params(self._h, self.params) != 0:
            raise ModelEvalFailed(f"update_params({self.model}) failed: {_lib.last_error()}")

    def update_energy(self, energy) -> None:
        self.energy = np.ascontiguousarray(energy, np.float64)
        self.n_flux = self.energy.size - 1
        if _lib.lib().relxill_b200_update_energy(self._h, self.energy, self.n_flux) != 0:
            raise ModelEvalFailed(f"update_energy({self.model}) failed: {_lib.last_error()}")

    def reuse_counts(self) -> dict:
        out = np.zeros(3, np.int64)
        _lib.lib().relxill_b200_reuse_counts(self._h, out)
        return dict(recomputed=int(out[0]), reused_rel=int(out[1]), reused_all=int(out[2]))

    def status(self) -> np.ndarray:
        st = np.zeros(self.n, np.int32)
        _lib.lib().relxill_b200_batch_status(self._h, st)
        return st

    def launches(self) -> int:
        return int(_lib.lib().relxill_b200_last_launches(self._h))

    def algorithmic_bytes(self) -> dict:
        out = np.zeros(8)
        _lib.lib().relxill_b200_algorithmic_bytes(self._h, out)
        return dict(total=out[0], distinct_rows=out[1], xillver=out[2], xillver_upper_bound=out[3],
                    line_profiles=out[4], zone_spectrum_values=out[5], xillver_union=out[6])

    def kernel_times(self) -> dict:
        names = (C.c_char_p * 16)()
        ms = np.zeros(16)
        cnt = np.zeros(16, np.int64)
        n = _lib.lib().relxill_b200_kernel_times(self._h, names, ms, cnt, 16)
        return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def probe(self, iv: int, what: str, max_len: int = 50 * 4096) -> np.ndarray:
        out = np.zeros(max_len)
        n = _lib.lib().relxill_b200_probe(self._h, iv, what.encode(), out, max_len)
        if n < 0:
            raise RuntimeError(f"probe({what}) failed: {_lib.last_error()}")
        return out[:n].copy()

    def close(self) -> None:
        if self._h:
            _lib.lib().relxill_b200_free_batch(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def lmod(model: str, energy, parameter, flux_in=None) -> np.ndarray:
    """Calls the XSPEC C symbol of `model` (lmodrelxilllp, ...) exactly as XSPEC would."""
    if model not in _lib.LMOD_SYMBOLS:
        raise ModelNotFound(model)
    energy = np.ascontiguousarray(energy, np.float64)
    parameter = np.ascontiguousarray(parameter, np.float64)
    flux = np.zeros(energy.size - 1) if flux_in is None else np.array(flux_in, np.float64)
    getattr(_lib.lib(), _lib.LMOD_SYMBOLS[model])(energy, energy.size - 1, parameter, 1, flux, None, b"")
    return flux


class LocalModel:
    """`LocalModel(name).set_par(par, value).eval_model(energy)` as in the reference (src/LocalModel.h)."""

    def __init__(self, model: str, params: Sequence[float] | None = None):
        if model not in PARAM_NAMES:
            raise ModelNotFound(model)
        self.model = model
        self.names = PARAM_NAMES[model]
        self._lower = [n.lower() for n in self.names]
        self.values = default_params(model) if params is None else np.array(params, np.float64)

    def set_par(self, name: str, value: float) -> "LocalModel":
        try:
            self.values[self._lower.index(name.lower())] = value
        except ValueError:
            raise KeyError(f"parameter not found: {name}")
        return self

    def get_par(self, name: str) -> float:
        return float(self.values[self._lower.index(name.lower())])

    def eval_model(self, energy, flux_in=None) -> np.ndarray:
        flux, status = batch_eval(self.model, energy, self.values[None, :], flux_in, return_status=True)
        if status[0] != 0:
            raise ModelEvalFailed(f"model evaluation failed (status {int(status[0])})")
        return flux[0]
