"""A pool of single-threaded reference processes (oracle/_ref, the unmodified reference): evaluates many parameter
vectors of ONE model in parallel on the host cores.  TEST INFRASTRUCTURE (and bench.py's cpu_baseline / reference arm).

The reference is not re-entrant and keeps process-global caches, so every worker is its own process, bound to one
table directory and one RELXILL_NUM_RZONES value for its whole life."""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

_REF = None


def _init(table_dir, zones, env):
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)                      # the reference prints banners / warnings on stdout
    for k, v in (env or {}).items():
        os.environ[k] = v
    global _REF
    from oracle.pyref import RefLocal
    if zones is None:
        os.environ.pop("RELXILL_NUM_RZONES", None)
    _REF = RefLocal(table_dir, zones)


def _eval(job):
    model, energy, params = job
    t0 = time.perf_counter()
    out = np.stack([_REF.eval(model, energy, p) for p in params]) if len(params) else np.zeros((0, energy.size - 1))
    return out, time.perf_counter() - t0


class RefPool:
    def __init__(self, table_dir, zones=None, procs=None, env=None):
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        self.cores = int(procs or cores)
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_init, initargs=(table_dir, zones, env))

    def eval_rows(self, model, energy, params, chunk=8):
        """flux [len(params), n_flux] of the rows, and the wall time of the parallel section"""
        params = np.atleast_2d(np.asarray(params, np.float64))
        energy = np.ascontiguousarray(energy, np.float64)
        jobs = [(model, energy, params[i:i + chunk]) for i in range(0, len(params), chunk)]
        t0 = time.perf_counter()
        res = self.pool.map(_eval, jobs, chunksize=1)
        dt = time.perf_counter() - t0
        return np.concatenate([r[0] for r in res]), dt

    def close(self):
        self.pool.terminate()
        self.pool.join()


def census(got, want, floor=1e-6, levels=(1e-5, 1e-8, 1e-10)):
    """Outlier census of a batch against the reference: per-bin relative error on the bins above `floor` of each
    spectrum's peak (the north_star metric), the worst bin and how many bins exceed each level."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    peak = np.abs(want).max(axis=1, keepdims=True)
    m = np.abs(want) > floor * peak
    err = np.zeros_like(want)
    np.divide(np.abs(got - want), np.abs(want), out=err, where=m)
    worst = np.unravel_index(int(np.argmax(err)), err.shape)
    out = {"rows": int(want.shape[0]), "bins_checked": int(m.sum()), "max_rel_err": float(err.max()),
           "worst_row": int(worst[0]), "worst_bin": int(worst[1])}
    for lv in levels:
        out[f"n_bins_over_{lv:.0e}"] = int((err > lv).sum())
    return out
