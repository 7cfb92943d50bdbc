#!/usr/bin/env python
"""bench.py — relxilllp spectra/second (batched parameter vectors, 3000-bin grid) on 1..8 B200.

  python bench.py --gpus N --steps K --warmup W              # this repo (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K --warmup W   # the unmodified reference on the host cores

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): relxilllp lamp post with
returning radiation, RELXILL_NUM_RZONES=50 radial zones, a batch of 4096 MCMC-walker parameter vectors per GPU
(Gaussian ball, seed 4321+rank), DefaultSpec 3000-bin log grid 0.1-1000 keV, synthetic tables of the published
layout and size (xillver-a-Ec5: 13x4x15x11x10 spectra x 2999 bins = 1.03 GB).

One "step" = one pass of the hot path over the batch.  `value` is timed with CUDA events on the launching stream,
inputs (interpreted parameter vectors, energy grid, tables) already resident in HBM; L2 is flushed between steps.
`e2e` goes through the C-ABI call relxill_batch_eval with pinned HOST buffers (host-side parameter interpretation,
H2D, kernels, D2H inside the timed region).  Multi-GPU: the batch is sharded one shard per rank (weak scaling,
no data-path collective); the only exchange is the NCCL all-gather of the result spectra, inside the timed step.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "relxilllp spectra/sec (batched params, 3000 bins)"
UNIT = "spectra/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="relxilllp")
    ap.add_argument("--batch", type=int, default=4096, help="parameter vectors per GPU")
    ap.add_argument("--zones", type=int, default=50)
    ap.add_argument("--bins", type=int, default=3000)
    ap.add_argument("--tables", default="bench", choices=["bench", "test"])
    ap.add_argument("--cpu-evals", type=int, default=24, help="reference evaluations per host core in the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records (ion gradient, config 4, config 5)")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": f"{args.model} lamp post + returning radiation, {args.zones} radial zones, "
                    f"{args.batch} MCMC-walker parameter vectors per GPU, {args.bins}-bin log grid 0.1-1000 keV "
                    "(BASELINE.json configs[2])",
        "model": args.model, "batch_per_gpu": args.batch, "global_batch": args.batch * world, "zones": args.zones,
        "bins": args.bins, "tables": f"synthetic '{args.tables}' size",
        "parallelism": f"parameter-vector sharding x{world}; NCCL all-gather of step k's spectra on a side stream under the kernels "
                       "of step k+1 (two output buffers), the last gather inside the timed region",
        "l2": "flushed between timed steps (256 MiB write, inside the timed region); every step also streams ~17 GB of scratch "
              "through the 126 MB L2",
        "state_cache": "off (every step recomputes every vector)",
    }


def make_tables(args, rank_local, cp=False):
    from relxill_b200.tables import synth
    d = synth.default_table_dir(args.tables)
    which = ("rel", "lp", "rrad") + (("xill",) if (not args.model.endswith("Cp") or cp) else ()) + (("xillcp",) if (args.model.endswith("Cp") or cp) else ())
    lock = d + ".lock"
    os.makedirs(os.path.dirname(d), exist_ok=True)
    if rank_local == 0:
        synth.generate(d, args.tables, which)
        open(lock, "w").write("ready")
    else:
        t0 = time.time()
        while not os.path.exists(lock) and time.time() - t0 < 1200:
            time.sleep(0.5)
        synth.generate(d, args.tables, which)  # no-op when stamped
    return d


# ------------------------------------------------------------------------------------------ reference on host cores
def _cpu_worker_init(table_dir, zones):
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)  # the reference prints banners/warnings on stdout
    global _REF, _GRID
    from oracle.pyref import RefLocal
    if zones is None:
        os.environ.pop("RELXILL_NUM_RZONES", None)
    _REF = RefLocal(table_dir, zones)


def _cpu_worker_eval(job):
    model, energy, params, collect = job
    t0 = time.perf_counter()
    out = [_REF.eval(model, energy, p) for p in params]
    return (np.stack(out) if collect else None), time.perf_counter() - t0


class CpuPool:
    """One single-threaded reference process per host core (the reference is not re-entrant)."""

    def __init__(self, table_dir, zones, model, energy, walkers=None):
        import multiprocessing as mp
        from common import walker_ball
        try:
            self.cores = len(os.sched_getaffinity(0))
        except AttributeError:
            self.cores = os.cpu_count() or 1
        self.model, self.energy = model, energy
        self.walkers = walker_ball(model, 4096, seed=4321) if walkers is None else walkers
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_worker_init, initargs=(table_dir, zones))
        # warm-up: tables loaded, xillver rows of the walker ball touched
        self.pool.map(_cpu_worker_eval, [(model, energy, self.walkers[i:i + 1], False) for i in range(self.cores)])
        self.cursor = self.cores

    def step(self, evals_per_core, collect=False):
        jobs, all_idx = [], []
        for _ in range(self.cores):
            idx = [(self.cursor + k) % len(self.walkers) for k in range(evals_per_core)]
            self.cursor += evals_per_core
            all_idx += idx
            jobs.append((self.model, self.energy, self.walkers[idx], collect))
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker_eval, jobs, chunksize=1)
        dt = time.perf_counter() - t0
        if collect:
            return np.array(all_idx), np.concatenate([r[0] for r in res]), self.cores * evals_per_core, dt
        return self.cores * evals_per_core, dt

    def close(self):
        self.pool.terminate()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyref
    from common import default_grid
    if not pyref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/librelxill_ref.so was not built"}))
        return
    tdir = make_tables(args, 0)
    energy = default_grid(args.bins)
    pool = CpuPool(tdir, args.zones, args.model, energy)
    per_core = max(2, min(args.cpu_evals, 8))
    for _ in range(args.warmup):
        pool.step(1)
    n_tot, t_tot = 0, 0.0
    for _ in range(args.steps):
        n, dt = pool.step(per_core)
        n_tot += n
        t_tot += dt
    pool.close()
    val = n_tot / t_tot
    sample = f"{per_core} evaluations per core per step x {args.steps} steps of the same walker batch"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": pool.cores, "kind": "reference", "sample": sample,
                         "build": "unmodified reference sources, gcc/g++ -O2, cfitsio/FFTW3 shims"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (2 ms period, so that even
    a 100 ms region gets tens of samples); `nvidia-smi -lms` is the fallback when NVML cannot be opened."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        import threading
        self.sm, self.mx, self.reasons = [], None, set()
        self.p = self.thread = None
        self._stop = threading.Event()
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            try:
                import torch
                h = nv.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(device).uuid))
            except Exception:  # noqa: BLE001
                h = nv.nvmlDeviceGetHandleByIndex(device)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                     ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                     ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                     ("hw_power_brake_slowdown", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown))

            def poll():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for nm, bit in names:
                            if r & bit:
                                self.reasons.add(nm)
                    except Exception:  # noqa: BLE001
                        pass
                    self._stop.wait(0.002)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.how = "nvml, 2 ms period"
        except Exception:  # noqa: BLE001
            self.how = "nvidia-smi -lms 20"
            self.path = tempfile.mktemp(suffix=".csv")
            self.f = open(self.path, "w")
            try:
                self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                           "-lms", "20", "-i", str(device)], stdout=self.f, stderr=subprocess.DEVNULL)
                time.sleep(0.5)   # let it print its first lines before the timed region starts
            except OSError:
                self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "how": self.how}
        if self.thread is not None:
            self._stop.set()
            self.thread.join(1.0)
            if self.sm:
                out.update(sm_mhz=float(np.median(self.sm)), sm_max_mhz=self.mx, reasons=sorted(self.reasons),
                           samples=len(self.sm))
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(3)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for ln in open(self.path):
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ ours
def _alg_bytes(ab, n, nz, nb):
    """Algorithmic bytes per launch of each kernel family (SURVEY.md §8d's per-vector figures x n)."""
    nex = ab.get("zone_spectrum_values") or 2999   # values per zone spectrum row as filed by k_xill / read by k_conv
    return {
        # table rows that any vector of the launch reads, once (vectors share rows through L2; the per-vector count is in
        # `table_rows_delivered`) + zone spectra out
        "k_xill": (ab.get("xillver_union") or ab["xillver"]) + n * nz * nex * 8.0,
        "k_line": n * (100 * 40 * 2 * 8.0 + 1000 * 6 * 8.0) + ab["line_profiles"],  # (a,mu0)-interpolated trff rows + radius scalars in, profiles out
        "k_conv": ab["line_profiles"] + n * (nz * nex * 8.0 + nb * 8.0),      # profiles + zone spectra in, spectrum out
        "k_fine": n * (2 * 4 * 40 * 16.0 * 100 + 1000 * 10 * 8.0),           # 4 corners x 100 radii x 40 g* float4 in, angle-distribution parts out
        "k_dist": n * (1000 * 10 * 8.0),
        "k_syspar": 2 * n * (4 * 3 * 100 * 4.0 + 2 * 2 * 3 * 100 * 4.0 + (3 * 2500 + 2 * 50000) * 8.0 + 7 * 1000 * 8.0),
        "k_zone": n * (4 * 1000 * 8.0),
        "k_nth": n * ((nz + 1) * 4 * 8.0 + 900 * 8.0),                       # per solve kTe in, 3 numbers out; the source's solution out
        "k_prim_nth": n * (2 * 4096 * 8.0 + nb * 8.0),
    }


class Workload:
    """One model + parameter shard on this rank: device-resident timing, end-to-end timing through the C ABI,
    per-kernel times.  All ranks call every method (the timings are reduced with MAX over the ranks)."""

    def __init__(self, ctx, model, params, zones, energy):
        self.ctx, self.model, self.params, self.zones, self.energy = ctx, model, np.ascontiguousarray(params), zones, energy
        rx, torch = ctx["rx"], ctx["torch"]
        rx.set_num_zones(zones)
        self.n, self.nb = len(params), energy.size - 1
        self.batch = rx.Batch(model, energy, self.params)
        self.outs = [torch.zeros((self.n, self.nb), dtype=torch.float64, device="cuda") for _ in range(2)]

    def _max_over_ranks(self, x):
        torch, dist = self.ctx["torch"], self.ctx["dist"]
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if self.ctx["world"] > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self):
        if self.ctx["world"] > 1:
            self.ctx["dist"].barrier()
        self.ctx["torch"].cuda.synchronize()

    def resident(self, steps, warmup, gather=False, sampler_factory=None):
        """K steps of the hot path with the batch resident in HBM.  gather: the NCCL all-gather of step k's spectra runs on a
        side stream under the kernels of step k+1 (two output buffers); the timed region ends when the last gather has
        landed.  Returns (ms total, max over ranks; launches; clocks)."""
        torch, dist, world = self.ctx["torch"], self.ctx["dist"], self.ctx["world"]
        self.ctx["rx"].set_num_zones(self.zones)
        stream = torch.cuda.current_stream()
        gather = gather and world > 1
        if gather:
            comm = self.ctx.setdefault("comm_stream", torch.cuda.Stream())
            gath = [torch.empty((world * self.n, self.nb), dtype=torch.float64, device="cuda") for _ in range(2)]
            c_done = [torch.cuda.Event() for _ in range(2)]
            g_done = [torch.cuda.Event() for _ in range(2)]
        flush = self.ctx["flush"]

        def step(k):
            o = self.outs[k % 2]
            if gather and k >= 2:
                stream.wait_event(g_done[k % 2])     # the gather of step k-2 has read this buffer
            self.batch.run(o.data_ptr(), stream.cuda_stream)
            if gather:
                c_done[k % 2].record(stream)
                comm.wait_event(c_done[k % 2])
                with torch.cuda.stream(comm):
                    dist.all_gather_into_tensor(gath[k % 2], o)
                    g_done[k % 2].record(comm)

        for k in range(warmup):
            step(k)
        if gather:
            stream.wait_stream(comm)
        self.barrier()
        assert (self.batch.status() == 0).all(), "some parameter vectors were rejected"
        sampler = sampler_factory() if sampler_factory else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(stream)
        for k in range(steps):
            flush.fill_(1)               # L2 flush between steps (inside the timed region: 256 MiB write, ~0.05 ms)
            step(k)
        if gather:
            stream.wait_stream(comm)
        e1.record(stream)
        self.barrier()
        ms = self._max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if sampler else None
        self.last_out = self.outs[(steps - 1) % 2]
        return ms, self.batch.launches() * steps, clocks

    def e2e(self, steps, warmup):
        """The same work through relxill_batch_eval with pinned HOST buffers: host-side parameter interpretation, H2D,
        kernels, D2H inside the timed region.  Every rank evaluates its own shard (no collective)."""
        torch, L = self.ctx["torch"], self.ctx["L"]
        self.ctx["rx"].set_num_zones(self.zones)
        h_par = torch.from_numpy(self.params.copy()).pin_memory()
        h_flux = torch.zeros((self.n, self.nb), dtype=torch.float64).pin_memory()
        p_np, f_np = h_par.numpy(), h_flux.numpy()
        st = np.zeros(self.n, np.int32)
        for _ in range(warmup):
            L.relxill_batch_eval(self.model.encode(), self.energy, self.nb, p_np, self.n, f_np, st)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            rc = L.relxill_batch_eval(self.model.encode(), self.energy, self.nb, p_np, self.n, f_np, st)
            assert rc == 0
        torch.cuda.synchronize()
        dt = self._max_over_ranks(time.perf_counter() - t0)
        self.h_flux = f_np
        npar = self.params.shape[1]
        return {"value": self.ctx["world"] * self.n * steps / dt, "unit": UNIT,
                "h2d_bytes_per_step": int(self.n * npar * 8 + (self.nb + 1) * 8),
                "d2h_bytes_per_step": int(self.n * self.nb * 8 + self.n * 4), "steps": steps,
                "api": "relxill_batch_eval (C ABI) with pinned host buffers"}

    def kernel_times(self):
        torch, L = self.ctx["torch"], self.ctx["L"]
        self.ctx["rx"].set_num_zones(self.zones)
        L.relxill_b200_set_profiling(1)
        self.batch.run(self.outs[0].data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        kt = self.batch.kernel_times()
        L.relxill_b200_set_profiling(0)
        return kt

    def close(self):
        self.batch.close()
        self.outs = None


def _roofline(ktimes, alg, peaks, fp64_peak, prof, prof_ok):
    """Dominant kernel of the step against BOTH rooflines; `bound` is the one it sits closer to."""
    tot_k = sum(v[0] for v in ktimes.values()) or 1.0
    dominant = max(ktimes, key=lambda k: ktimes[k][0])
    k_ms, k_cnt = ktimes[dominant]
    t_launch = k_ms / max(k_cnt, 1) * 1e-3
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    hbm_ach = alg.get(dominant, 0.0) / t_launch / 1e9
    hbm = {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "peak_source": hbm_src,
           "algorithmic_bytes_per_launch": alg.get(dominant, 0.0)}
    kk = (prof or {}).get("kernels", {}).get(dominant) if prof_ok else None
    fp64, traffic = None, None
    if kk and kk.get("fp64_flop"):
        # flops per launch: SASS-level thread instruction counts (DADD/DMUL = 1, DFMA = 2) of an ncu capture of the same
        # build, workload and batch (they do not depend on the clock); time: this run's CUDA events
        ach = kk["fp64_flop"] / max(kk.get("launches", 1), 1) * max(k_cnt, 1) / (k_ms * 1e-3) / 1e12
        stale = abs(kk["ms"] - k_ms) > 0.15 * k_ms
        fp64 = {"achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak and fp64_peak > 0 else None,
                "peak_source": "measured in this run (DFMA microkernel, relxill_b200_measure_fp64_peak)",
                "flop_per_launch": kk["fp64_flop"] / max(kk.get("launches", 1), 1),
                "flop_source": "static: " + str(prof.get("source")), "capture_ms": kk["ms"], "capture_matches_run": not stale,
                "fp64_pipe_pct_in_capture": kk.get("fp64_pipe_pct"), "issue_active_pct_in_capture": kk.get("issue_active_pct")}
        traffic = kk["dram_bytes_read"] + kk["dram_bytes_write"]
    bound = "fp64" if (fp64 and fp64["frac"] and fp64["frac"] > hbm["frac"]) else "hbm"
    g = fp64 if bound == "fp64" else hbm
    out = {"kernel": dominant, "bound": bound, "achieved": g["achieved"], "peak": g["peak"], "unit": g["unit"], "frac": g["frac"],
           "traffic": traffic, "traffic_source": ("static: " + str(prof.get("source"))) if traffic is not None else None,
           "share_of_step": k_ms / tot_k, "ms_per_launch": t_launch * 1e3, "hbm": hbm, "fp64": fp64}
    # the whole step against the FP64 roofline (all kernels of the capture)
    if prof_ok and fp64_peak and fp64_peak > 0:
        step_flop = sum(float(x.get("fp64_flop") or 0.0) for x in prof["kernels"].values())
        if step_flop > 0:
            out["step_fp64"] = {"flop_per_step": step_flop, "tflops": step_flop / (tot_k * 1e-3) / 1e12,
                                "frac": step_flop / (tot_k * 1e-3) / 1e12 / fp64_peak}
    return out


def _config5_sweep(rx, n_a=32, n_h=32, n_i=16):
    """BASELINE.json configs[4]: relxilllp with returning radiation on a spin x height x inclination grid."""
    base = rx.default_params("relxilllp")
    names = [x.lower() for x in rx.PARAM_NAMES["relxilllp"]]
    A, H, I = np.meshgrid(np.linspace(0.0, 0.998, n_a), np.geomspace(2.0, 50.0, n_h), np.linspace(5.0, 85.0, n_i), indexing="ij")
    P = np.tile(base, (A.size, 1))
    P[:, names.index("a")] = A.ravel()
    P[:, names.index("h")] = H.ravel()
    P[:, names.index("incl")] = I.ravel()
    P[:, names.index("switch_returnrad")] = 1
    return P


def run_ours(args):
    import torch
    import torch.distributed as dist
    import relxill_b200 as rx
    from relxill_b200 import _lib
    from common import default_grid, walker_ball, sample_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    extras = not args.no_extras
    t_wall0 = time.time()
    tdir = make_tables(args, local, cp=extras or args.model.endswith("Cp"))
    rx.init(tdir, local)
    rx.set_cache(False)   # every timed step recomputes every vector; the state cache is measured separately below
    L = _lib.lib()
    energy = default_grid(args.bins)
    n, nb = args.batch, args.bins
    ctx = {"rx": rx, "torch": torch, "dist": dist, "L": L, "world": world, "rank": rank,
           "flush": torch.empty(256 << 20, dtype=torch.uint8, device="cuda")}

    # ---- the metric workload (BASELINE.json configs[2])
    params = walker_ball(args.model, n, seed=4321 + rank)
    main = Workload(ctx, args.model, params, args.zones, energy)
    ms_total, launches, clocks = main.resident(args.steps, args.warmup, gather=True,
                                               sampler_factory=(lambda: ClockSampler(local)) if rank == 0 else None)
    value = world * n * args.steps / (ms_total * 1e-3)
    e2e = main.e2e(max(2, min(args.steps, 5)), max(1, min(2, args.warmup)))
    assert np.allclose(main.h_flux, main.last_out.cpu().numpy(), rtol=1e-12, atol=0), "e2e result differs from the resident run"
    resident_result = main.last_out.clone()
    ktimes = main.kernel_times()
    ab = main.batch.algorithmic_bytes()

    # ---- sub-records: the other configurations of BASELINE.json, each rank its own shard, no collective
    sub = {}
    if extras:
        def sub_record(key, model, P, zones, what):
            w = Workload(ctx, model, P, zones, energy)
            ms, _, _ = w.resident(3, 2)
            e = w.e2e(2, 1)
            kt = w.kernel_times()
            tot = sum(v[0] for v in kt.values()) or 1.0
            dom = max(kt, key=lambda k: kt[k][0])
            rec = {"workload": what, "model": model, "vectors_per_gpu": len(P), "zones": zones,
                   "value": world * len(P) * 3 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / 3, "steps": 3, "warmup": 2,
                   "e2e": e["value"], "dominant_kernel": dom, "dominant_share": kt[dom][0] / tot,
                   "kernels_ms": {k: round(v[0], 3) for k, v in kt.items()}}
            w.close()
            sub[key] = rec

        P = walker_ball("relxilllpCp", n, seed=4321 + rank)   # iongrad_type 1 (power-law gradient), index ~ N(1, 0.2)
        sub_record("iongrad", "relxilllpCp", P, args.zones,
                   f"relxilllpCp, power-law ionisation gradient (iongrad_type 1), {args.zones} zones, {n} MCMC walkers per GPU, "
                   "6-D nthcomp table (the half of BASELINE configs[2] with a zone-dependent xi)")
        P = sample_params("relxillCp", 8192, seed=99 + rank)
        sub_record("cfg4", "relxillCp", P, None,
                   "BASELINE configs[3]: relxillCp, 8192 uniform-random parameter vectors per GPU (65536 over 8 GPUs), default zones, 6-D table")
        P = _config5_sweep(rx)
        sub_record("cfg5", "relxilllp", np.ascontiguousarray(P[rank::world]), None,
                   "BASELINE configs[4]: relxilllp + returning radiation, 32 x 32 x 16 sweep over spin x height x inclination "
                   f"(16384 points, rows interleaved over {world} GPU(s)), default zones")
        rx.set_num_zones(args.zones)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    fp64_peak = float(L.relxill_b200_measure_fp64_peak())
    prof, prof_ok = None, False
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        prof_ok = (prof["config"]["model"] == args.model and prof["config"]["batch"] == args.batch
                   and prof["config"]["zones"] == args.zones and prof["config"]["n_flux"] == nb)
    except Exception:  # noqa: BLE001
        pass
    alg = _alg_bytes(ab, n, args.zones, nb)
    roofline = _roofline(ktimes, alg, peaks, fp64_peak, prof, prof_ok)
    hbm_peak = roofline["hbm"]["peak"]
    xk = ktimes.get("k_xill", (0.0, 1))
    rows_delivered = None
    if xk[0]:
        kx = (prof or {}).get("kernels", {}).get("k_xill") if prof_ok else None
        rows_delivered = {
            "kernel": "k_xill", "what": "table rows delivered per second, distinct rows counted per vector (SURVEY 8d); MCMC walkers share rows "
                                        "through L2/L1, so this is NOT DRAM traffic and is not a roofline fraction",
            "delivered_gbs": alg["k_xill"] / (xk[0] / max(xk[1], 1) * 1e-3) / 1e9,
            "distinct_corner_rows_per_vector": ab["distinct_rows"] / n, "xillver_bytes_distinct": ab["xillver"],
            "dram_gbs_ncu": ((kx["dram_bytes_read"] + kx["dram_bytes_write"]) / (kx["ms"] * 1e-3) / 1e9) if kx else None,
            "dram_frac_of_hbm_peak_ncu": ((kx["dram_bytes_read"] + kx["dram_bytes_write"]) / (kx["ms"] * 1e-3) / 1e9 / hbm_peak) if kx else None,
            "dram_source": ("static: " + str(prof.get("source"))) if kx else None}

    # ---- supplemental: the device-resident state cache on an MCMC-like sequence in which half of the walkers stay where
    #      they were (rejected proposals) from one step to the next.  Not the headline: `value`/`e2e` run with it off.
    state_cache = None
    try:
        rx.set_cache(True)
        batch, out, stream = main.batch, main.outs[0], torch.cuda.current_stream()
        moved = params.copy()
        moved[::2] = walker_ball(args.model, n, seed=991 + rank)[::2]
        seqs = [moved, params, moved, params]
        batch.run(out.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for q in seqs:
            batch.update_params(q)
            batch.run(out.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        dt_c = time.perf_counter() - t0
        state_cache = {"value": n * len(seqs) / dt_c, "unit": UNIT, "reuse_last_step": batch.reuse_counts(),
                       "scenario": "update_params + run per step, every second walker unchanged since the previous step "
                                   "(host interpretation and H2D of the parameters included)"}
    except Exception as exc:  # noqa: BLE001
        state_cache = {"error": str(exc)}
    finally:
        rx.set_cache(False)

    # ---- the reference on the host cores: a bounded sample of the same walkers; its spectra double as the parity check
    #      of the resident result on the bench-size tables
    cpu, parity, config1 = None, None, None
    if not args.no_cpu_baseline:
        from oracle import pyref
        if pyref.available():
            from refpool import census
            pool = CpuPool(tdir, args.zones, args.model, energy, walkers=params)
            idx, want, n_ev, dt_cpu = pool.step(args.cpu_evals, collect=True)
            cpu = {"value": n_ev / dt_cpu, "unit": UNIT, "cores": pool.cores, "kind": "reference",
                   "sample": f"{args.cpu_evals} walker evaluations on each of {pool.cores} cores "
                             f"({n_ev} spectra, {dt_cpu:.1f} s wall), one single-threaded process per core, caches on",
                   "build": "unmodified reference sources, gcc/g++ -O2, cfitsio/FFTW3 shims"}
            got = resident_result[torch.as_tensor(idx, device="cuda")].cpu().numpy()
            c = census(got, want)
            parity = {"against": "oracle/_ref (unmodified reference) on the same walkers and the same bench-size tables",
                      "rows": c["rows"], "bins_checked": c["bins_checked"], "max_rel_err": c["max_rel_err"],
                      "n_bins_over_1e-5": c["n_bins_over_1e-05"], "n_bins_over_1e-8": c["n_bins_over_1e-08"],
                      "tolerance": "1e-5 relative on bins above 1e-6 of the spectrum peak (north_star)"}
            pool.close()
            config1 = _config1_latency(rx, tdir, energy)
        else:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}

    cfg = workload_config(args, world)
    cfg["wall_s"] = round(time.time() - t_wall0, 1)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "kernels_ms": {k: round(v[0], 3) for k, v in ktimes.items()}, "table_rows_delivered": rows_delivered,
        "iongrad": sub.get("iongrad"), "cfg4": sub.get("cfg4"), "cfg5": sub.get("cfg5"), "config1": config1,
        "state_cache": state_cache,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _config1_latency(rx, tdir, energy):
    """BASELINE.json configs[0]: relline, default parameters, single evaluations.  The reference in ONE process (its own
    harness perturbs a parameter per call so that its cache does not answer, test/speed/speed_test.cpp:82-114) next to the
    latency of the drop-in symbol lmodrelline."""
    import multiprocessing as mp
    p0 = rx.default_params("relline")
    names = [x.lower() for x in rx.PARAM_NAMES["relline"]]
    ia = names.index("a")
    P = np.tile(p0, (21, 1))
    P[1:, ia] = 0.998 - 1e-3 * np.arange(1, 21)
    try:
        pool = mp.get_context("spawn").Pool(1, initializer=_cpu_worker_init, initargs=(tdir, None))
        pool.map(_cpu_worker_eval, [("relline", energy, P[:1], False)])          # tables loaded
        _, dt = pool.map(_cpu_worker_eval, [("relline", energy, P[1:], False)])[0]
        pool.terminate()
        cpu_ms = dt / 20 * 1e3
    except Exception as exc:  # noqa: BLE001
        return {"error": str(exc)}
    rx.lmod("relline", energy, p0)
    t0 = time.perf_counter()
    for k in range(20):
        rx.lmod("relline", energy, P[k + 1])
    gpu_ms = (time.perf_counter() - t0) / 20 * 1e3
    return {"workload": "BASELINE configs[0]: relline, default parameters (spin perturbed per call), 3000-bin grid, single evaluations",
            "reference_cpu_ms_per_eval": cpu_ms, "reference_threads": 1, "lmodrelline_gpu_ms_per_call": gpu_ms}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
