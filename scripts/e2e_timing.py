import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["RELXILL_B200_TIMING"] = "1"
import numpy as np, torch
import relxill_b200 as rx
from relxill_b200 import _lib
from relxill_b200.tables import synth
from common import default_grid, walker_ball
T = synth.generate(synth.default_table_dir("bench"), "bench", ("rel", "lp", "rrad", "xill"))
rx.init(T); rx.set_num_zones(50)
e = default_grid(3000); n = 4096
P = walker_ball("relxilllp", n)
hp = torch.from_numpy(P.copy()).pin_memory(); hf = torch.zeros((n, 3000), dtype=torch.float64).pin_memory()
st = np.zeros(n, np.int32); L = _lib.lib()
for it in range(4):
    t = time.perf_counter()
    L.relxill_batch_eval(b"relxilllp", e, 3000, hp.numpy(), n, hf.numpy(), st)
    print("call %d: %.1f ms" % (it, (time.perf_counter() - t) * 1e3), flush=True)
f2 = np.zeros((n, 3000))
t = time.perf_counter(); L.relxill_batch_eval(b"relxilllp", e, 3000, P, n, f2, st); print("pageable: %.1f ms" % ((time.perf_counter() - t) * 1e3))
