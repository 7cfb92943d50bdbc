"""Stage-by-stage GPU vs oracle comparison, printed as a table (development aid, run by hand on a GPU box:
`python tests/gpu_check.py`; not collected by pytest — the parity tests proper are test_gpu_parity.py).  Lives under
tests/ because it uses the oracle, which only the tests, smoke() and bench.py's CPU legs may touch."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relxill_b200 as rx
from relxill_b200.tables import synth
from oracle.pyoracle import Oracle

T = synth.generate(synth.default_table_dir("test"), "test")
rx.init(T)
nzenv = int(os.environ.get("NZ", "0"))
rx.set_num_zones(nzenv)
o = Oracle(T, nzenv or None)
e = rx.default_energy_grid()
rng = np.random.default_rng(int(os.environ.get("SEED", "7")))
U = rng.uniform

def relerr(a, b, floor=1e-6):
    m = np.abs(b) > floor * np.abs(b).max()
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m])))

def sample(m):
    z = U(0, .3) * (rng.random() < .5)
    if m == 'relline': return [U(1, 8), U(0, 6), U(0, 6), U(2, 100), U(-0.998, 0.998), U(5, 85), -U(1, 5), U(50, 1000), z, rng.integers(0, 3)]
    if m == 'relline_lp': return [U(1, 8), U(1.5, 100), U(-0.998, 0.998), U(5, 85), -U(1, 5), U(50, 1000), z, rng.integers(0, 3), U(1, 3.4), rng.integers(0, 2)]
    if m == 'relxill': return [U(0, 6), U(0, 6), U(2, 100), U(-0.998, 0.998), U(5, 85), -U(1, 5), U(50, 1000), z, U(1, 3.4), U(0, 4.7), U(.5, 10), U(5, 1000), U(-2, 10)]
    if m == 'relxilllp': return [U(1.5, 100), U(0, .5) * (rng.random() < .5), U(-0.998, 0.998), U(5, 85), -U(1, 5), U(50, 1000), z, U(1, 3.4), U(0, 4.7), U(.5, 10), U(5, 1000), U(-2, 10), rng.integers(0, 2), rng.integers(0, 2)]
    if m == 'relxillCp': return [U(5, 85), U(-0.998, 0.998), -U(1, 5), U(50, 1000), U(2, 100), U(0, 6), U(0, 6), z, U(1.2, 3.4), U(0, 4.7), U(15, 20), U(.5, 10), U(1, 400), U(-2, 10)]
    if m == 'relxilllpCp': return [U(5, 85), U(-0.998, 0.998), -U(1, 5), U(50, 1000), U(1.5, 100), U(0, .5) * (rng.random() < .5), U(1.2, 3.4), U(0, 4.7), U(15, 20), U(.5, 10), U(1, 400), U(-2, 10), z, U(0, 3), rng.integers(0, 3), rng.integers(0, 2), rng.integers(0, 2)]

models = os.environ.get("MODELS", "relline,relline_lp,relxill,relxilllp").split(",")
n = int(os.environ.get("N", "8"))
for m in models:
    P = np.array([rx.default_params(m)] + [sample(m) for _ in range(n - 1)], float)
    t = time.time()
    b = rx.Batch(m, e, P, keep_intermediates=True)
    import torch
    out = torch.zeros((n, e.size - 1), dtype=torch.float64, device="cuda")
    b.run(out.data_ptr())
    torch.cuda.synchronize()
    dt = time.time() - t
    flux = out.cpu().numpy()
    st = b.status()
    print(f"== {m}: status {st.tolist()}  ({dt*1e3:.1f} ms incl. first-call setup)")
    for i in range(n):
        fo = o.eval(m, e, P[i])
        line = f"  [{i}] final {relerr(flux[i], fo):.2e}"
        if os.environ.get("STAGES", "1") == "1":
            sp = o.syspar(m, P[i])
            for k in ("re", "gmin", "gmax", "del_emit", "del_inc"):
                line += f" {k} {relerr(b.probe(i, k), sp[k], 0):.1e}"
            line += f" trff {relerr(b.probe(i, 'trff'), sp['trff'].ravel(), 0):.1e} cosne {relerr(b.probe(i, 'cosne'), sp['cosne'].ravel(), 0):.1e}"
            if m.startswith("relxill"):
                sg = o.stages(m, P[i])
                line += f" emis2 {relerr(b.probe(i, 'emis'), sg['emis2'], 0):.1e}"
                for k in ("lxi", "ect", "eshift", "normch", "corr_flux", "corr_gshift"):
                    line += f" {k} {relerr(b.probe(i, k), sg[k], 0):.1e}"
                line += f" relflux {relerr(b.probe(i, 'relflux'), sg['relflux'].ravel()):.1e}"
                line += f" dist {relerr(b.probe(i, 'dist'), sg['dist'].ravel()):.1e}"
                if rx.get_xill_grid():   # zone spectra live on the convolution grid: rebin the oracle's like src/Relxill.cpp:461-463
                    xe, ce = b.probe(i, 'xill_ener'), o.conv_grid()
                    want = np.array([o.rebin(ce, xe, row) for row in sg['xill']])
                    line += f" xillc {relerr(b.probe(i, 'xillc', max_len=50 * 4096), want.ravel()):.1e}"
                else:
                    line += f" xill {relerr(b.probe(i, 'xill'), sg['xill'].ravel()):.1e}"
                line += f" total {relerr(b.probe(i, 'total'), sg['total']):.1e}"
            else:
                line += f" emis {relerr(b.probe(i, 'emis'), sp['emis'], 0):.1e}"
        print(line, flush=True)
    # host-buffer API and the lmod symbol
    f2 = rx.batch_eval(m, e, P)
    f3 = rx.lmod(m, e, P[0])
    print(f"  batch_eval vs run: {np.abs(f2 - flux).max():.1e}; lmod vs run[0]: {np.abs(f3 - flux[0]).max():.1e}")
    b.close()
