"""Writes profiles/ncu_traffic.json (per-kernel DRAM bytes and pipe utilisation per launch) from an
`ncu --set full` capture of `bench.py`, read by bench.py for roofline.traffic.
usage: python scripts/ncu_traffic.py gpurun_out/prof_big3.ncu-rep <tag> [model batch zones n_flux]"""
import csv, json, os, subprocess, sys
rep, tag = sys.argv[1], sys.argv[2]
cfg = dict(model=sys.argv[3] if len(sys.argv) > 3 else "relxilllp", batch=int(sys.argv[4]) if len(sys.argv) > 4 else 4096,
           zones=int(sys.argv[5]) if len(sys.argv) > 5 else 50, n_flux=int(sys.argv[6]) if len(sys.argv) > 6 else 3000)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
def val(r, name):
    i = col[name]
    return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
kernels = {}
for r in rows[2:]:   # launches of the same kernel inside the captured step (k_syspar runs twice) are added up
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    k = kernels.setdefault(name, dict(ms=0.0, dram_bytes_read=0.0, dram_bytes_write=0.0, warp_instructions=0.0, launches=0,
                                      _w=0.0, fp64_pipe_pct=0.0, issue_active_pct=0.0, warps_active_pct=0.0))
    ms = val(r, "gpu__time_duration.sum")
    k["ms"] += ms
    k["launches"] += 1
    k["dram_bytes_read"] += val(r, "dram__bytes_read.sum")
    k["dram_bytes_write"] += val(r, "dram__bytes_write.sum")
    k["warp_instructions"] += val(r, "smsp__inst_executed.sum")
    k["registers"] = int(val(r, "launch__registers_per_thread"))
    for key, metric in (("fp64_pipe_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active")):
        k[key] += ms * val(r, metric)    # time-weighted over the launches
    k["_w"] += ms
for k in kernels.values():
    w = k.pop("_w") or 1.0
    for key in ("fp64_pipe_pct", "issue_active_pct", "warps_active_pct"):
        k[key] /= w
# FP64 floating-point operations per launch, counted from the SASS-level thread instruction counts of the capture
# (DADD / DMUL = 1, DFMA = 2 per active thread; conversions, compares and MUFU seeds are not counted)
import collections, re
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr2, kern, flops = None, None, collections.Counter()
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Kernel Name":
        kern = r[1].replace("rx::", "").replace("void ", "").split("(")[0].split("<")[0] if len(r) > 1 else None
        continue
    if r and r[0] == "Address":
        hdr2 = r
        continue
    if hdr2 is None or len(r) < len(hdr2):
        continue
    try:
        n = int(r[hdr2.index("Thread Instructions Executed")])
    except ValueError:
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[hdr2.index("Source")])
    if m and m.group(1) in ("DADD", "DMUL", "DFMA"):
        flops[kern] += n * (2 if m.group(1) == "DFMA" else 1)
for k in kernels:
    kernels[k]["fp64_flop"] = float(flops.get(k, 0))
    kernels[k]["fp64_tflops"] = flops.get(k, 0) / (kernels[k]["ms"] * 1e-3) / 1e12 if kernels[k]["ms"] else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(dict(source=f"profiles/ncu_{tag}.txt (ncu --set full --clock-control none, one launch per kernel, bench.py --steps 1 --warmup 1)",
               config=cfg, kernels=kernels), open(os.path.join(root, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(kernels, indent=1))
