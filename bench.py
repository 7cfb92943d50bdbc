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
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": f"{args.model} lamp post + returning radiation, {args.zones} radial zones, "
                    f"{args.batch} MCMC-walker parameter vectors per GPU, {args.bins}-bin log grid 0.1-1000 keV "
                    "(BASELINE.json configs[2])",
        "model": args.model, "batch_per_gpu": args.batch, "global_batch": args.batch * world, "zones": args.zones,
        "bins": args.bins, "tables": f"synthetic '{args.tables}' size",
        "parallelism": f"parameter-vector sharding x{world}, NCCL all-gather of the spectra",
        "l2": "flushed between timed steps (256 MiB write)",
        "state_cache": "off (every step recomputes every vector)",
    }


def make_tables(args, rank_local):
    from relxill_b200.tables import synth
    d = synth.default_table_dir(args.tables)
    which = ("rel", "lp", "rrad", "xill") if not args.model.endswith("Cp") else ("rel", "lp", "rrad", "xillcp")
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
    _REF = RefLocal(table_dir, zones)


def _cpu_worker_eval(job):
    model, energy, params = job
    t0 = time.perf_counter()
    for p in params:
        _REF.eval(model, energy, p)
    return time.perf_counter() - t0


class CpuPool:
    """One single-threaded reference process per host core (the reference is not re-entrant)."""

    def __init__(self, table_dir, zones, model, energy):
        import multiprocessing as mp
        from common import walker_ball
        try:
            self.cores = len(os.sched_getaffinity(0))
        except AttributeError:
            self.cores = os.cpu_count() or 1
        self.model, self.energy = model, energy
        self.walkers = walker_ball(model, 4096, seed=4321)
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_worker_init, initargs=(table_dir, zones))
        # warm-up: tables loaded, xillver rows of the walker ball touched
        self.pool.map(_cpu_worker_eval, [(model, energy, self.walkers[i:i + 1]) for i in range(self.cores)])
        self.cursor = self.cores

    def step(self, evals_per_core):
        jobs = []
        for _ in range(self.cores):
            idx = [(self.cursor + k) % len(self.walkers) for k in range(evals_per_core)]
            self.cursor += evals_per_core
            jobs.append((self.model, self.energy, self.walkers[idx]))
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker_eval, jobs, chunksize=1)
        dt = time.perf_counter() - t0
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
def run_ours(args):
    import torch
    import torch.distributed as dist
    import relxill_b200 as rx
    from common import default_grid, walker_ball

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tdir = make_tables(args, local)
    rx.init(tdir, local)
    rx.set_num_zones(args.zones)
    rx.set_cache(False)   # every timed step recomputes every vector; the state cache is measured separately below
    energy = default_grid(args.bins)
    n, nb = args.batch, args.bins
    params = walker_ball(args.model, n, seed=4321 + rank)
    npar = params.shape[1]

    batch = rx.Batch(args.model, energy, params)
    out = torch.zeros((n, nb), dtype=torch.float64, device="cuda")
    gathered = torch.empty((world * n, nb), dtype=torch.float64, device="cuda") if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def step():
        batch.run(out.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    assert (batch.status() == 0).all(), "some walkers were rejected"
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for e0, e1 in ev:
        flush.fill_(1)           # L2 flush, outside the timed events
        e0.record(stream)
        step()
        e1.record(stream)
    barrier()
    ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    launches = batch.launches() * args.steps
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * n * args.steps / (ms_total * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers
    h_par = torch.from_numpy(params.copy()).pin_memory()
    h_flux = torch.zeros((n, nb), dtype=torch.float64).pin_memory()
    e_np, p_np, f_np = energy, h_par.numpy(), h_flux.numpy()
    st = np.zeros(n, np.int32)
    from relxill_b200 import _lib
    L = _lib.lib()
    for _ in range(max(1, min(2, args.warmup))):
        L.relxill_batch_eval(args.model.encode(), e_np, nb, p_np, n, f_np, st)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        rc = L.relxill_batch_eval(args.model.encode(), e_np, nb, p_np, n, f_np, st)
        assert rc == 0
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * n * e2e_steps / float(t.item())
    assert np.allclose(f_np, out.cpu().numpy(), rtol=1e-12, atol=0), "e2e result differs from the resident run"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel device times (one extra, untimed-for-the-metric step with events around every launch)
    L.relxill_b200_set_profiling(1)
    batch.run(out.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    ktimes = batch.kernel_times()
    L.relxill_b200_set_profiling(0)
    ab = batch.algorithmic_bytes()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    tot_k = sum(v[0] for v in ktimes.values()) or 1.0
    dominant = max(ktimes, key=lambda k: ktimes[k][0])
    nz = args.zones
    nex = ab.get("zone_spectrum_values") or 2999   # values per zone spectrum row as filed by k_xill / read by k_conv
    alg = {  # algorithmic bytes per launch of each kernel family
        "k_xill": ab["xillver"] + n * nz * nex * 8.0,                       # distinct table rows + zone spectra out
        "k_line": n * (1000 * 40 * 2 * 8.0 + 1000 * 5 * 8.0) + ab["line_profiles"],  # fine trff + radius scalars in, profiles out
        "k_conv": ab["line_profiles"] + n * (nz * nex * 8.0 + nb * 8.0),     # profiles + zone spectra in, spectrum out
        "k_fine": n * (2 * 4 * 40 * 16.0 * 100 + 1000 * 40 * 4 * 8.0),      # 4 corners x 100 radii x 40 g* float4 + fine tables out
        "k_dist": n * (1000 * 40 * 4 * 8.0),
        "k_syspar": 2 * n * (4 * 3 * 100 * 4.0 + 2 * 2 * 3 * 100 * 4.0 + (3 * 2500 + 2 * 50000) * 8.0 + 7 * 1000 * 8.0),
        "k_zone": n * (4 * 1000 * 8.0),
        "k_nth": n * (3 * 900 * 64 * 8.0),
        "k_prim_nth": n * (2 * 4096 * 8.0 + nb * 8.0),
    }
    k_ms, k_cnt = ktimes[dominant]
    achieved = alg.get(dominant, 0.0) / (k_ms / max(k_cnt, 1) * 1e-3) / 1e9
    traffic, compute = None, None   # dram bytes / pipe utilisation of this kernel from the committed ncu --set full capture
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        same = (prof["config"]["model"] == args.model and prof["config"]["batch"] == args.batch
                and prof["config"]["zones"] == args.zones and prof["config"]["n_flux"] == nb)
        if same and dominant in prof["kernels"]:
            kk = prof["kernels"][dominant]
            traffic = kk["dram_bytes_read"] + kk["dram_bytes_write"]
            # FP64 side of the roofline: flops counted by ncu (SASS thread instructions DADD/DMUL/DFMA of one launch of
            # this kernel) over the launch time measured live above; peak = SMs x 64 FMA/clk x 2 x max SM clock
            props = torch.cuda.get_device_properties(local)
            fp64_peak = props.multi_processor_count * 128 * float((clocks or {}).get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
            fp64_rate = (kk["fp64_flop"] / (k_ms / max(k_cnt, 1) * 1e-3) / 1e12) if kk.get("fp64_flop") else None
            compute = {"fp64_pipe_pct": kk.get("fp64_pipe_pct"), "issue_active_pct": kk.get("issue_active_pct"),
                       "fp64_tflops": fp64_rate, "fp64_peak_tflops": fp64_peak,
                       "fp64_frac": (fp64_rate / fp64_peak) if fp64_rate else None,
                       "fp64_peak_source": "nominal: SM count x 64 FMA/clk x 2 x max SM clock",
                       "source": prof.get("source")}
            # the whole step: FP64 flops of all its kernels (same capture) over the measured step time
            step_flop = sum(float(x.get("fp64_flop") or 0.0) for x in prof["kernels"].values())
            if step_flop > 0:
                compute["step_fp64_tflops"] = step_flop / (ms_total / args.steps * 1e-3) / 1e12
                compute["step_fp64_frac"] = compute["step_fp64_tflops"] / fp64_peak
    except Exception:  # noqa: BLE001
        pass
    roofline = {"kernel": dominant, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "compute": compute, "peak_source": peak_src,
                "share_of_step": ktimes[dominant][0] / tot_k,
                "algorithmic_bytes_per_launch": alg.get(dominant, 0.0)}
    if roofline["frac"] > 1.0:
        roofline["note"] = ("algorithmic bytes count every vector's distinct table rows (SURVEY 8d); MCMC walkers share them, the rows "
                            "are served from L2/L1 and the figure exceeds the HBM peak: the kernel is not HBM-bound in this workload")
    xk = ktimes.get("k_xill", (0.0, 1))
    hbm_stage = {"kernel": "k_xill", "achieved": alg["k_xill"] / (xk[0] / max(xk[1], 1) * 1e-3) / 1e9 if xk[0] else None,
                 "unit": "GB/s", "peak": peak, "distinct_corner_rows_per_vector": ab["distinct_rows"] / n,
                 "xillver_bytes_distinct": ab["xillver"], "xillver_bytes_upper_bound": ab["xillver_upper_bound"]}
    if hbm_stage["achieved"]:
        hbm_stage["frac"] = hbm_stage["achieved"] / peak
        hbm_stage["note"] = ("distinct rows are counted per vector; walkers share rows through L2/L1, so this is delivered table "
                             "bandwidth, not DRAM traffic (see roofline.traffic / profiles/ncu_traffic.json; scripts/cfg4_probe.py is the "
                             "HBM-bound case)")
        hbm_stage["frac_of_upper_bound_traffic"] = (ab["xillver_upper_bound"] + n * nz * nex * 8.0) / (xk[0] * 1e-3) / 1e9 / peak

    # ---- supplemental: the device-resident state cache on an MCMC-like sequence in which half of the walkers stay where
    #      they were (rejected proposals) from one step to the next.  Not the headline: `value`/`e2e` run with it off.
    state_cache = None
    try:
        rx.set_cache(True)
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

    cpu = None
    if not args.no_cpu_baseline:
        from oracle import pyref
        if pyref.available():
            pool = CpuPool(tdir, args.zones, args.model, energy)
            n_ev, dt_cpu = pool.step(args.cpu_evals)
            pool.close()
            cpu = {"value": n_ev / dt_cpu, "unit": UNIT, "cores": pool.cores, "kind": "reference",
                   "sample": f"{args.cpu_evals} walker evaluations on each of {pool.cores} cores "
                             f"({n_ev} spectra, {dt_cpu:.1f} s wall), one single-threaded process per core, caches on",
                   "build": "unmodified reference sources, gcc/g++ -O2, cfitsio/FFTW3 shims"}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world), "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(n * npar * 8 + (nb + 1) * 8),
                "d2h_bytes_per_step": int(n * nb * 8 + n * 4), "steps": e2e_steps,
                "api": "relxill_batch_eval (C ABI) with pinned host buffers"},
        "gpu_launches": int(launches), "roofline": roofline, "hbm_stage": hbm_stage, "cpu_baseline": cpu,
        "kernels_ms": {k: round(v[0], 3) for k, v in ktimes.items()}, "state_cache": state_cache,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
