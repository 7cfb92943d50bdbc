"""Shared by the tests and tests/golden/make_golden.py: seeded parameter samplers and the parity metric."""
import numpy as np

MODELS = ["relline", "relline_lp", "relconv", "relconv_lp", "relxill", "relxilllp", "relxillCp", "relxilllpCp",
          "xillver", "xillverCp"]
# neutron-star / CO table flavours; their golden vectors live in golden_v2_nsco.npz
NSCO_MODELS = ["xillverNS", "relxillNS", "xillverCO", "relxillCO"]
ALL_MODELS = MODELS + NSCO_MODELS
# north_star tolerance: relative error <= 1e-5 per bin on bins above 1e-6 of the spectrum peak
RTOL = 1e-5
PEAK_FLOOR = 1e-6


def default_grid(n=3000, emin=0.1, emax=1000.0):
    i = np.arange(n + 1, dtype=np.float64)
    e = np.exp(i / float(n) * (np.log(emax) - np.log(emin)) + np.log(emin))
    e[-1] = emax
    return e


def relerr(a, b, floor=PEAK_FLOOR):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    peak = np.abs(b).max()
    if peak == 0:
        return float(np.abs(a).max())
    m = np.abs(b) > floor * peak
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m])))


def sample_params(model, n, seed):
    """n random parameter vectors inside the lmodel.dat soft ranges, row 0 = defaults-like."""
    rng = np.random.default_rng(seed)
    U = rng.uniform
    rows = []
    for _ in range(n):
        z = U(0, .3) * (rng.random() < .5)
        a = U(-0.998, 0.998)
        incl = U(5, 85)
        rin = -U(1, 5)
        rout = U(50, 1000)
        h = U(1.5, 100)
        beta = U(0, .5) * (rng.random() < .5)
        if model == "relline":
            r = [U(1, 8), U(0, 6), U(0, 6), U(2, 100), a, incl, rin, rout, z, rng.integers(0, 3)]
        elif model == "relconv":
            r = [U(0, 6), U(0, 6), U(2, 100), a, incl, rin, rout, rng.integers(0, 3)]
        elif model == "relline_lp":
            r = [U(1, 8), h, a, incl, rin, rout, z, rng.integers(0, 3), U(1, 3.4), rng.integers(0, 2)]
        elif model == "relconv_lp":
            r = [h, beta, a, incl, rin, rout, rng.integers(0, 3), U(1, 3.4), rng.integers(0, 2)]
        elif model == "relxill":
            r = [U(0, 6), U(0, 6), U(2, 100), a, incl, rin, rout, z, U(1, 3.4), U(0, 4.7), U(.5, 10), U(5, 1000), U(-2, 10)]
        elif model == "relxilllp":
            r = [h, beta, a, incl, rin, rout, z, U(1, 3.4), U(0, 4.7), U(.5, 10), U(5, 1000), U(-2, 10),
                 rng.integers(0, 2), rng.integers(0, 2)]
        elif model == "relxillCp":
            r = [incl, a, rin, rout, U(2, 100), U(0, 6), U(0, 6), z, U(1.2, 3.4), U(0, 4.7), U(15, 20), U(.5, 10),
                 U(1, 400), U(-2, 10)]
        elif model == "relxilllpCp":
            r = [incl, a, rin, rout, h, beta, U(1.2, 3.4), U(0, 4.7), U(15, 20), U(.5, 10), U(1, 400), U(-2, 10), z,
                 U(0, 3), rng.integers(0, 3), rng.integers(0, 2), rng.integers(0, 2)]
        elif model == "xillver":
            r = [U(1, 3.4), U(.5, 10), U(5, 1000), U(0, 4.7), z, U(3, 89), U(-2, 5)]
        elif model == "xillverCp":
            r = [U(1.2, 3.4), U(.5, 10), U(1, 400), U(0, 4.7), U(15, 20), z, U(3, 89), U(-2, 5)]
        elif model == "xillverNS":
            r = [U(.5, 10), U(.5, 10), U(15, 19), U(1, 4.7), z, U(3, 89), U(-2, 5)]
        elif model == "relxillNS":
            r = [U(0, 6), U(0, 6), U(2, 100), a, incl, rin, rout, z, U(.5, 10), U(1, 4.7), U(.5, 10), U(15, 19), U(-2, 10)]
        elif model == "xillverCO":
            r = [U(1, 2.8), U(1, 1000), U(.05, .5), U(.01, 1), U(2, 1000), z, U(18.2, 87), U(-2, 5)]
        elif model == "relxillCO":
            r = [U(0, 6), U(0, 6), U(2, 100), a, incl, rin, rout, z, U(1, 2.8), U(1, 1000), U(.05, .5), U(.01, 1),
                 U(2, 1000), U(-2, 10)]
        else:
            raise KeyError(model)
        rows.append([float(x) for x in r])
    return np.array(rows, np.float64)


def walker_ball(model, n, seed=4321):
    """BASELINE config 3: Gaussian ball of MCMC walkers around the defaults, clipped to the hard limits."""
    rng = np.random.default_rng(seed)
    if model == "relxilllp":
        #            h    beta  a     Incl Rin  Rout  z  gamma logxi Afe Ecut  refl rr boost
        c = np.array([6.0, 0.0, 0.9, 30., -1., 400., 0., 2.0, 3.1, 1.0, 300., 1.0, 1, 0])
        s = np.array([0.5, 0.0, 0.03, 3.0, 0., 0., 0., 0.05, 0.1, 0.2, 30., 0.2, 0, 0])
        lo = np.array([2.0, 0, -0.998, 3, -100, 1, 0, 1.0, 0, 0.5, 5, 0, 0, 0])
        hi = np.array([500, .99, 0.998, 87, -1, 1000, 10, 3.4, 4.7, 10, 1000, 10, 1, 1])
    elif model == "relxilllpCp":
        #            Incl a    Rin Rout  h   beta gamma logxi logN Afe kTe refl z idx type rr boost
        c = np.array([30., 0.9, -1., 400., 6.0, 0.0, 2.0, 3.1, 16., 1.0, 60., 1.0, 0., 1.0, 1, 1, 0])
        s = np.array([3.0, 0.03, 0., 0., 0.5, 0.0, 0.05, 0.1, 0.3, 0.2, 6.0, 0.2, 0., 0.2, 0, 0, 0])
        lo = np.array([3, -0.998, -100, 1, 2.0, 0, 1.2, 0, 15, 0.5, 1, 0, 0, 0, 0, 0, 0])
        hi = np.array([87, 0.998, -1, 1000, 500, .99, 3.4, 4.7, 20, 10, 400, 10, 10, 3, 2, 1, 1])
    elif model == "relxill":
        c = np.array([3., 3., 15., 0.9, 30., -1., 400., 0., 2.0, 3.1, 1.0, 300., 1.0])
        s = np.array([0.3, 0.3, 2., 0.03, 3., 0., 0., 0., 0.05, 0.1, 0.2, 30., 0.2])
        lo = np.array([0, 0, 1, -0.998, 3, -100, 1, 0, 1.0, 0, 0.5, 5, 0])
        hi = np.array([10, 10, 1000, 0.998, 87, -1, 1000, 10, 3.4, 4.7, 10, 1000, 10])
    else:
        raise KeyError(model)
    p = c[None, :] + s[None, :] * rng.standard_normal((n, c.size))
    return np.clip(p, lo, hi)


def config_cases():
    """(key, model, zones, full parameter matrix, picked rows) of the BASELINE.json configurations 2-4 (config 5 is the
    returning-radiation grid built in the tests).  tests/golden/make_golden.py evaluates the picked rows with the
    unmodified reference (golden_v3_configs.npz); the GPU tests evaluate the FULL batches and compare those rows."""
    cases = []
    cases.append(("cfg2_relxill", "relxill", None, sample_params("relxill", 1024, seed=1234), [0, 17, 333, 640, 1023]))
    cases.append(("cfg3_relxilllp", "relxilllp", 50, walker_ball("relxilllp", 4096), [0, 1, 1000, 2047, 4095]))
    cases.append(("cfg3_relxilllpCp", "relxilllpCp", 50, walker_ball("relxilllpCp", 4096), [0, 1, 1000, 2047, 4095]))
    cases.append(("cfg4_relxillCp", "relxillCp", None, sample_params("relxillCp", 5000, seed=99), [1, 1666, 2507, 4998]))
    P = sample_params("relxilllpCp", 2500, seed=99)
    P[:, 14] = 0
    cases.append(("cfg4_relxilllpCp", "relxilllpCp", None, P, [1, 833, 1257, 2498]))
    return cases


def edge_params(model):
    """Parameter vectors that exercise the branches the random samplers never reach:
    switch_returnrad in {-1, 2} (src/Rellp.cpp:483-509: -1 = reflection of the returning radiation only is NOT what it
    means — 1 adds the returning emissivity, 2 / -1 replace the direct one), positive Rin / Rout (gravitational
    radii instead of multiples of the ISCO), negative h (multiples of the event horizon) and negative Rbr
    (src/ModelDefinition.cpp:202-226)."""
    import relxill_b200 as rx   # only the default vectors (host-side table in the library)
    base = rx.default_params(model)
    names = [n.lower() for n in rx.PARAM_NAMES[model]]
    rows = []

    def row(**kw):
        p = base.copy()
        for k, v in kw.items():
            p[names.index(k.lower())] = v
        rows.append(p)

    if "switch_returnrad" in names:
        row(switch_returnrad=-1, a=0.9)
        row(switch_returnrad=2, a=0.9)
        row(switch_returnrad=2, a=-0.3)          # negative spin: no correction factors (src/Relxill.cpp:338-341)
        row(switch_returnrad=0, a=0.9)
    row(rin=6.5, rout=350.0, a=0.7)              # positive radii: used as given
    row(rin=1.5, rout=80.0, a=0.5)               # positive Rin below the ISCO (4.23): clamped up to it
    if "h" in names:
        row(h=-2.5, a=0.95)                      # negative height: in units of the event horizon
        row(h=-1.05, a=0.3)                      # below 1.1 r+: clamped
    if "rbr" in names:
        row(rbr=-3.0, index1=5.0, index2=2.0)    # negative break radius: in units of the ISCO
        row(rbr=5000.0, index1=4.0)              # beyond Rout: clamped to Rout
        row(rbr=0.5, index1=4.0)                 # inside Rin: clamped to Rin
    return np.array(rows)
