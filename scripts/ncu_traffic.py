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
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    kernels[name] = dict(
        ms=val(r, "gpu__time_duration.sum"),
        dram_bytes_read=val(r, "dram__bytes_read.sum"), dram_bytes_write=val(r, "dram__bytes_write.sum"),
        fp64_pipe_pct=val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        issue_active_pct=val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        warps_active_pct=val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        registers=int(val(r, "launch__registers_per_thread")), warp_instructions=val(r, "smsp__inst_executed.sum"))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(dict(source=f"profiles/ncu_{tag}.txt (ncu --set full --clock-control none, one launch per kernel, bench.py --steps 1 --warmup 1)",
               config=cfg, kernels=kernels), open(os.path.join(root, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(kernels, indent=1))
