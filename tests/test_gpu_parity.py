"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> librelxill_b200.so), against the
oracle on the same seeded inputs, against the golden vectors generated from the unmodified reference, and
through size-independent properties at full batch size.

Tolerance (north_star): relative error <= 1e-5 per bin on bins above 1e-6 of the spectrum peak."""
import os

import numpy as np
import pytest

from common import PEAK_FLOOR, RTOL, config_cases, default_grid, edge_params, relerr, sample_params, walker_ball

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
GOLDEN_NSCO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2_nsco.npz")
GOLDEN_CFG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v3_configs.npz")
NSCO_MODELS = ["xillverNS", "relxillNS", "xillverCO", "relxillCO"]
GPU_MODELS = ["relline", "relline_lp", "relconv", "relconv_lp", "relxill", "relxilllp", "relxillCp", "relxilllpCp",
              "xillver", "xillverCp"] + NSCO_MODELS


def _conv_input(e):
    return np.exp(-0.5 * ((np.log(0.5 * (e[1:] + e[:-1])) - np.log(6.4)) / 0.03) ** 2) + 1e-3


@pytest.mark.parametrize("model", GPU_MODELS)
def test_vs_golden_reference_vectors(rx, model):
    g = np.load(GOLDEN_NSCO if model in NSCO_MODELS else GOLDEN)
    e, P, F = g["energy"], g[f"{model}_params"], g[f"{model}_flux"]
    rx.set_num_zones(None)
    fin = g["conv_input"] if model.startswith("relconv") else None
    got, st = rx.batch_eval(model, e, P, fin, return_status=True)
    assert (st == 0).all()
    for a, b in zip(got, F):
        assert relerr(a, b) < RTOL


def test_vs_golden_50_zones(rx):
    g = np.load(GOLDEN)
    rx.set_num_zones(50)
    try:
        got = rx.batch_eval("relxilllp", g["energy"], g["relxilllp_z50_params"])
        for a, b in zip(got, g["relxilllp_z50_flux"]):
            assert relerr(a, b) < RTOL
    finally:
        rx.set_num_zones(None)


@pytest.mark.parametrize("model", GPU_MODELS)
def test_vs_oracle_random(rx, oracle, model):
    e = default_grid(3000)
    P = sample_params(model, 24, seed=2024 + len(model))
    rx.set_num_zones(None)
    oracle.set_num_zones(None)
    fin = _conv_input(e) if model.startswith("relconv") else None
    got, st = rx.batch_eval(model, e, P, fin, return_status=True)
    assert (st == 0).all()
    worst = 0.0
    for a, p in zip(got, P):
        b = oracle.eval_conv(model, e, p, fin) if fin is not None else oracle.eval(model, e, p)
        worst = max(worst, relerr(a, b))
    assert worst < RTOL, worst


def test_metric_config_vs_oracle(rx, oracle):
    """relxilllp, RELXILL_NUM_RZONES=50, MCMC-walker ball (BASELINE config 3), 3000-bin grid."""
    e = default_grid(3000)
    P = walker_ball("relxilllp", 12)
    rx.set_num_zones(50)
    oracle.set_num_zones(50)
    try:
        got = rx.batch_eval("relxilllp", e, P)
        for a, p in zip(got, P):
            assert relerr(a, oracle.eval("relxilllp", e, p)) < RTOL
    finally:
        rx.set_num_zones(None)
        oracle.set_num_zones(None)


def test_ion_gradient_50_zones(rx, oracle):
    """relxilllpCp with iongrad_type 1 (power law) and 2 (alpha disk), RELXILL_NUM_RZONES=50 (BASELINE config 3)."""
    e = default_grid(3000)
    P = walker_ball("relxilllpCp", 8)
    P[:, 14] = [1, 2, 1, 2, 1, 2, 1, 2]
    rx.set_num_zones(50)
    oracle.set_num_zones(50)
    try:
        got, st = rx.batch_eval("relxilllpCp", e, P, return_status=True)
        assert (st == 0).all()
        for a, p in zip(got, P):
            assert relerr(a, oracle.eval("relxilllpCp", e, p)) < RTOL
        g = np.load(GOLDEN)
        got = rx.batch_eval("relxilllpCp", g["energy"], g["relxilllpCp_z50_params"])
        for a, b in zip(got, g["relxilllpCp_z50_flux"]):
            assert relerr(a, b) < RTOL
    finally:
        rx.set_num_zones(None)
        oracle.set_num_zones(None)


def test_constant_density_env(rx, oracle, monkeypatch):
    """RELXILL_CONSTANT_DENSITY=1 (src/relutility.c:372-382, src/IonGradient.cpp:155-158): alpha-disk ionisation
    gradient at constant density; read per call like the reference."""
    e = default_grid(1500)
    P = walker_ball("relxilllpCp", 4)
    P[:, 14] = 2
    base = rx.batch_eval("relxilllpCp", e, P)
    monkeypatch.setenv("RELXILL_CONSTANT_DENSITY", "1")
    got = rx.batch_eval("relxilllpCp", e, P)
    assert max(relerr(a, b) for a, b in zip(got, base)) > 1e-4   # the switch changes the spectrum
    for a, p in zip(got, P):
        assert relerr(a, oracle.eval("relxilllpCp", e, p)) < RTOL
    monkeypatch.delenv("RELXILL_CONSTANT_DENSITY")
    again = rx.batch_eval("relxilllpCp", e, P)
    assert np.array_equal(again, base)


def test_stage_parity(rx, oracle):
    """Intermediates of the pipeline against the oracle's (tight tolerances: same arithmetic)."""
    import torch
    e = default_grid(3000)
    P = sample_params("relxilllp", 4, seed=11)
    P[:, 12] = 1  # returning radiation on: exercises the second system-parameter pass
    P[:, 2] = np.abs(P[:, 2])
    b = rx.Batch("relxilllp", e, P, keep_intermediates=True)
    out = torch.zeros((4, 3000), dtype=torch.float64, device="cuda")
    b.run(out.data_ptr())
    torch.cuda.synchronize()
    for i, p in enumerate(P):
        sp, sg = oracle.syspar("relxilllp", p), oracle.stages("relxilllp", p)
        for k in ("re", "gmin", "gmax", "del_emit", "del_inc"):
            np.testing.assert_allclose(b.probe(i, k), sp[k], rtol=1e-13, err_msg=k)
        np.testing.assert_allclose(b.probe(i, "trff"), sp["trff"].ravel(), rtol=1e-13)
        np.testing.assert_allclose(b.probe(i, "cosne"), sp["cosne"].ravel(), rtol=1e-13)
        np.testing.assert_allclose(b.probe(i, "emis"), sg["emis2"], rtol=1e-12)
        for k in ("lxi", "ect", "eshift", "normch", "corr_flux", "corr_gshift"):
            np.testing.assert_allclose(b.probe(i, k), sg[k], rtol=1e-12, err_msg=k)
        assert relerr(b.probe(i, "relflux"), sg["relflux"]) < 1e-11
        np.testing.assert_allclose(b.probe(i, "dist"), sg["dist"].ravel(), rtol=1e-11)
        if rx.get_xill_grid():   # zone spectra filed on the convolution grid: the oracle's, rebinned like src/Relxill.cpp:461-463
            xe, ce = b.probe(i, "xill_ener"), oracle.conv_grid()
            want = np.array([oracle.rebin(ce, xe, row) for row in sg["xill"]])
            assert relerr(b.probe(i, "xillc", max_len=50 * 4096), want.ravel()) < 1e-11
        else:
            assert relerr(b.probe(i, "xill"), sg["xill"]) < 1e-11
        assert relerr(b.probe(i, "total"), sg["total"]) < RTOL
    b.close()


def test_xill_grid_variants_agree(rx, oracle):
    """The zone spectra filed on the convolution grid (rebin folded into the table-corner refresh of k_xill) and on
    the table grid (rebinned per zone in k_conv) are the same linear map applied in a different order.  The zone spectra
    differ in the last bit; behind the FFT that is 1e-16 of the spectrum's peak on every bin, i.e. up to ~1e-10 relative
    on the faintest bins the parity metric looks at (1e-6 of the peak; measured 1.1e-10, the same floor as against the
    reference).  Both paths match the oracle.  State cache off: every evaluation must run the kernels."""
    e = default_grid(3000)
    assert rx.get_xill_grid()
    rx.set_cache(False)   # the second evaluation must run the kernels, not return the retained spectra
    try:
        for model in ("relxill", "relxilllp", "relxilllpCp", "relxillNS", "relxillCO"):
            P = sample_params(model, 6, seed=5)
            a = rx.batch_eval(model, e, P)
            rx.set_xill_grid(False)
            b = rx.batch_eval(model, e, P)
            rx.set_xill_grid(True)
            assert max(relerr(x, y) for x, y in zip(a, b)) < 1e-8, model
            assert relerr(a[0], oracle.eval(model, e, P[0])) < RTOL
            assert relerr(b[0], oracle.eval(model, e, P[0])) < RTOL
    finally:
        rx.set_xill_grid(True)
        rx.set_cache(True)


def test_xill_any_table_instantiation(rx):
    """k_xill has its row length and inclination count as template constants for the 2999-bin xillver tables and an
    instantiation with run-time strides for any other table; both must give the same spectra, on both zone-spectrum
    grids and for the 5-D and 6-D tables."""
    from relxill_b200 import _lib
    e = default_grid(1200)
    L = _lib.lib()
    rx.set_cache(False)   # every evaluation runs the kernels
    try:
        for conv in (True, False):
            rx.set_xill_grid(conv)
            for model in ("relxilllp", "relxilllpCp", "relxillNS"):
                P = sample_params(model, 5, seed=21)
                a = rx.batch_eval(model, e, P)
                L.relxill_b200_set_xill_generic(1)
                b = rx.batch_eval(model, e, P)
                L.relxill_b200_set_xill_generic(0)
                assert max(relerr(x, y) for x, y in zip(a, b)) < 1e-13, (model, conv)
    finally:
        L.relxill_b200_set_xill_generic(0)
        rx.set_xill_grid(True)
        rx.set_cache(True)


def test_lmod_symbols_match_batch(rx):
    e = default_grid(500)
    for model in GPU_MODELS:
        p = rx.default_params(model)
        fin = _conv_input(e) if model.startswith("relconv") else None
        a = rx.lmod(model, e, p, fin)
        b = rx.batch_eval(model, e, p[None, :], fin)[0]
        np.testing.assert_array_equal(a, b)
        assert np.isfinite(a).all() and a.sum() > 0


def test_local_model_interface(rx, oracle):
    e = default_grid(800)
    lm = rx.LocalModel("relxilllp").set_par("h", 4.0).set_par("a", 0.5).set_par("Incl", 55.0)
    got = lm.eval_model(e)
    p = rx.default_params("relxilllp")
    p[0], p[2], p[3] = 4.0, 0.5, 55.0
    assert relerr(got, oracle.eval("relxilllp", e, p)) < RTOL


def test_invalid_parameters_are_reported_per_vector(rx):
    e = default_grid(300)
    P = np.tile(rx.default_params("relxill"), (4, 1))
    P[1, 3] = 0.9999      # spin above 0.9982 -> rejected by the reference's check_parameter_bounds
    P[2, 4] = 89.5        # inclination outside 3..87 deg
    flux, st = rx.batch_eval("relxill", e, P, return_status=True)
    assert st[0] == 0 and st[3] == 0 and st[1] != 0 and st[2] != 0
    assert (flux[1] == 0).all() and (flux[2] == 0).all() and flux[0].sum() > 0
    np.testing.assert_array_equal(flux[0], flux[3])
    with pytest.raises(rx.ModelEvalFailed):
        rx.LocalModel("relxill", P[1]).eval_model(e)


def test_ragged_and_tiny_grids(rx, oracle):
    rng = np.random.default_rng(5)
    e = np.sort(np.concatenate([[0.05, 2500.0], rng.uniform(0.2, 80.0, 57)]))  # irregular, beyond the conv grid
    p = rx.default_params("relxilllp")
    assert relerr(rx.batch_eval("relxilllp", e, p[None, :])[0], oracle.eval("relxilllp", e, p)) < RTOL
    e1 = np.array([3.0, 7.0])  # a single bin
    np.testing.assert_allclose(rx.batch_eval("relline", e1, rx.default_params("relline")[None, :])[0],
                               oracle.eval("relline", e1, rx.default_params("relline")), rtol=1e-12)


def test_redshift_and_negative_refl_frac(rx, oracle):
    e = default_grid(1000)
    p = rx.default_params("relxilllp")
    p[6] = 0.4          # z
    p[11] = -1.5        # reflection only
    assert relerr(rx.batch_eval("relxilllp", e, p[None, :])[0], oracle.eval("relxilllp", e, p)) < RTOL


# ---------------------------------------------------------------- properties at BASELINE batch sizes
def test_full_batch_properties(rx):
    """4096 walkers x 50 zones x 3000 bins (the metric configuration): order invariance, chunk invariance,
    linearity in refl_frac and flux conservation of the output rebin."""
    e = default_grid(3000)
    n = 4096
    P = walker_ball("relxilllp", n)
    rx.set_num_zones(50)
    try:
        f = rx.batch_eval("relxilllp", e, P)
        assert np.isfinite(f).all() and (f.sum(axis=1) > 0).all()
        # (1) a permuted batch gives the permuted result, bit for bit (vectors are independent)
        perm = np.random.default_rng(1).permutation(n)[:512]
        np.testing.assert_array_equal(rx.batch_eval("relxilllp", e, P[perm]), f[perm])
        # (2) refl_frac enters linearly: f(rf) = |rf| R + primary for rf >= 0 and no boost switch
        sub = P[:256].copy()
        sub[:, 13] = 0
        f1, f2, f3 = (rx.batch_eval("relxilllp", e, np.column_stack([sub[:, :11], np.full(256, rf), sub[:, 12:]]))
                      for rf in (1.0, 2.0, 3.0))
        np.testing.assert_allclose(f3 - f2, f2 - f1, rtol=1e-9, atol=1e-12 * f1.max())
        # (3) the output rebin conserves flux: a 2x coarser grid holds the pairwise sums
        fc = rx.batch_eval("relxilllp", e[::2], P[:256])
        np.testing.assert_allclose(fc, f[:256, 0::2] + f[:256, 1::2], rtol=1e-10)
    finally:
        rx.set_num_zones(None)


def test_relline_normalisation_full_batch(rx):
    """relline is normalised to unit photon flux (reference test/unit/test-relxill.cpp:29-40 checks the same
    integral on the published tables)."""
    e = default_grid(3000, 0.05, 12.0)
    P = sample_params("relline", 1024, seed=9)
    P[:, 0] = 1.0
    P[:, 8] = 0.0
    f = rx.batch_eval("relline", e, P)
    np.testing.assert_allclose(f.sum(axis=1), 1.0, rtol=1e-12)


# ---------------------------------------------------------------- device-resident state cache
def _run_batch(rx, model, e, P, flux_in=None):
    import torch
    b = rx.Batch(model, e, P)
    out = torch.zeros((len(P), e.size - 1), dtype=torch.float64, device="cuda")
    if flux_in is not None:
        out.copy_(torch.from_numpy(np.broadcast_to(flux_in, out.shape).copy()))
    b.run(out.data_ptr())
    torch.cuda.synchronize()
    return b, out


# (model, index of a parameter that only the xillver half reads, index of a relativistic parameter)
_CACHE_CASES = [("relxill", 9, 3), ("relxilllp", 8, 0), ("relxillCp", 9, 1), ("relxilllpCp", 7, 4), ("relxillNS", 9, 3)]


@pytest.mark.parametrize("model,i_xill,i_rel", _CACHE_CASES)
def test_state_cache_is_bit_identical(rx, model, i_xill, i_rel):
    """Re-running a batch after update_params re-uses what is still valid (whole spectrum / relativistic half) and
    must give the bits of a fresh evaluation — the reference's caches are result-transparent too."""
    import torch
    e = default_grid(1500)
    P0 = sample_params(model, 12, seed=5)
    if model == "relxilllp":
        P0[:6, 12] = 0      # first half without returning radiation: their emissivity does not depend on xillver
        P0[6:, 12] = 1
        P0[6:, 2] = 0.5     # positive spin -> correction factors are on
    if model == "relxilllpCp":
        P0[:, 15] = 0
    P1 = P0.copy()
    chg_x, chg_r = [2, 3, 8, 9], [4, 5, 10, 11]
    P1[chg_x, i_xill] *= 1.05
    P1[chg_r, i_rel] *= 0.97
    rx.set_cache(True)
    b, out = _run_batch(rx, model, e, P0)
    assert b.reuse_counts() == dict(recomputed=12, reused_rel=0, reused_all=0)
    b.update_params(P1)
    b.run(out.data_ptr())
    torch.cuda.synchronize()
    got = out.cpu().numpy().copy()
    cnt = b.reuse_counts()
    st = b.status()
    n_all = int(sum(1 for i in (0, 1, 6, 7) if st[i] == 0))
    assert cnt["reused_all"] == n_all, cnt
    if model == "relxilllp":   # rows 8, 9 carry correction factors: a xillver change reaches their emissivity
        assert cnt["reused_rel"] == int(sum(1 for i in (2, 3) if st[i] == 0)), cnt
    else:
        assert cnt["reused_rel"] == int(sum(1 for i in chg_x if st[i] == 0)), cnt
    # a new energy grid (and a redshift): only the final rebin is repeated
    e2 = default_grid(700, 0.3, 200.0)
    P2 = P1.copy()
    P2[:, rx.PARAM_NAMES[model].index("z")] = 0.1
    b.update_energy(e2)
    b.update_params(P2)
    out2 = torch.zeros((12, 700), dtype=torch.float64, device="cuda")
    b.run(out2.data_ptr())
    torch.cuda.synchronize()
    assert b.reuse_counts()["reused_all"] == int((st == 0).sum())
    # fresh evaluations (another batch takes the arena over: nothing to re-use)
    rx.set_cache(False)
    try:
        fresh_b, fresh = _run_batch(rx, model, e, P1)
        assert fresh_b.reuse_counts()["recomputed"] == 12
        np.testing.assert_array_equal(got, fresh.cpu().numpy())
        _, fresh2 = _run_batch(rx, model, e2, P2)
        np.testing.assert_array_equal(out2.cpu().numpy(), fresh2.cpu().numpy())
    finally:
        rx.set_cache(True)
    # the arena now belongs to the last fresh batch: the first batch recomputes everything
    b.run(out2.data_ptr())
    torch.cuda.synchronize()
    assert b.reuse_counts()["recomputed"] == 12
    np.testing.assert_array_equal(out2.cpu().numpy(), fresh2.cpu().numpy())


def test_state_cache_line_and_conv_models(rx):
    import torch
    e = default_grid(1200)
    fin = _conv_input(e)
    rx.set_cache(True)
    try:
        # relconv: the input spectrum changes from call to call, the relativistic kernel does not
        P = sample_params("relconv", 6, seed=3)
        b, out = _run_batch(rx, "relconv", e, P, fin)
        first = out.cpu().numpy().copy()
        out.copy_(torch.from_numpy(np.broadcast_to(2.0 * fin, out.shape).copy()))
        b.run(out.data_ptr())
        torch.cuda.synchronize()
        assert b.reuse_counts()["reused_rel"] == 6
        np.testing.assert_allclose(out.cpu().numpy(), 2.0 * first, rtol=1e-13)
        # relline: unchanged vectors keep their profile, changed ones are recomputed
        P = sample_params("relline", 6, seed=4)
        b, out = _run_batch(rx, "relline", e, P)
        P1 = P.copy()
        P1[3:, 4] *= 0.9
        b.update_params(P1)
        b.run(out.data_ptr())
        torch.cuda.synchronize()
        assert b.reuse_counts()["reused_rel"] == 3
        rx.set_cache(False)
        _, fresh = _run_batch(rx, "relline", e, P1)
        np.testing.assert_array_equal(out.cpu().numpy(), fresh.cpu().numpy())
    finally:
        rx.set_cache(True)


def test_state_cache_behind_the_xspec_symbols(rx):
    """An XSPEC fit calls lmodrelxill again and again with one parameter changed: the retained batch must not change
    any result."""
    e = default_grid(800)
    p = rx.default_params("relxill")
    seq = []
    for k, (i, fac) in enumerate([(0, 1.0), (9, 1.01), (9, 1.02), (3, 0.99), (7, 1.0), (12, 2.0), (12, 2.0)]):
        q = p.copy()
        q[i] *= fac
        if i == 7:
            q[7] = 0.05
        seq.append(q)
        p = q
    rx.set_cache(True)
    cached = [rx.lmod("relxill", e, q) for q in seq]
    rx.set_cache(False)
    try:
        plain = [rx.lmod("relxill", e, q) for q in seq]
    finally:
        rx.set_cache(True)
    for a, b in zip(cached, plain):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("model", ["relxilllp", "relxilllpCp"])
def test_state_cache_survives_an_mcmc_sized_host_batch(rx, model):
    """4096 walkers x 50 zones through the HOST API (relxill_batch_eval cuts the batch into pipelined
    pieces; each piece keeps its slice of the arena), every second walker unchanged on the second call: 2048 whole
    spectra are re-used, and the result carries the bits of a fresh evaluation.  relxilllpCp runs with the ionisation
    gradient (its Kompaneets work arrays are part of the arena: one chunk of 4096)."""
    e = default_grid(3000)
    n = 4096
    P0 = walker_ball(model, n)
    if model == "relxilllpCp":
        P0[:, 14] = 1
    P1 = P0.copy()
    P1[::2] = walker_ball(model, n, seed=991)[::2]
    if model == "relxilllpCp":
        P1[:, 14] = 1
    rx.set_num_zones(50)
    rx.set_cache(True)
    try:
        rx.batch_eval(model, e, P0)
        assert rx.last_eval_reuse() == dict(recomputed=n, reused_rel=0, reused_all=0)
        got = rx.batch_eval(model, e, P1)
        cnt = rx.last_eval_reuse()
        assert cnt["reused_all"] == n // 2 and cnt["recomputed"] + cnt["reused_rel"] == n // 2, cnt
        rx.set_cache(False)
        fresh = rx.batch_eval(model, e, P1)
        np.testing.assert_array_equal(got, fresh)
    finally:
        rx.set_cache(True)
        rx.set_num_zones(None)


# ---------------------------------------------------------------- the other BASELINE.json configurations
@pytest.mark.parametrize("key,model,zones", [("cfg2_relxill", "relxill", None), ("cfg3_relxilllp", "relxilllp", 50),
                                              ("cfg3_relxilllpCp", "relxilllpCp", 50), ("cfg4_relxillCp", "relxillCp", None),
                                              ("cfg4_relxilllpCp", "relxilllpCp", None), ("cfg5_relxilllp", "relxilllp", None)])
def test_baseline_configs_vs_golden_reference_vectors(rx, key, model, zones):
    """BASELINE.json configs 2-5 at their full batch sizes (1024 random relxill vectors; 4096 MCMC walkers with 50 zones;
    the Cp shards streamed through several chunks; the returning-radiation sweep): the rows the fixture holds —
    evaluated by the UNMODIFIED reference, tests/golden/make_golden.py configs — must come out of the full batch
    within the north_star tolerance."""
    g = np.load(GOLDEN_CFG)
    e, rows, want = g["energy"], g[f"{key}_rows"], g[f"{key}_flux"]
    if key == "cfg5_relxilllp":
        P = _config5_grid(rx)
    else:
        P = next(c[3] for c in config_cases() if c[0] == key)
    np.testing.assert_array_equal(P[rows], g[f"{key}_params"])     # the samplers still produce the fixture's vectors
    rx.set_num_zones(zones)
    try:
        f, st = rx.batch_eval(model, e, P, return_status=True)
    finally:
        rx.set_num_zones(None)
    assert (st[rows] == 0).all()
    for i, w in zip(rows, want):
        assert relerr(f[i], w) < RTOL, (key, int(i))


def _config5_grid(rx):
    base = rx.default_params("relxilllp")
    rows = []
    for a in np.linspace(0.0, 0.998, 8):
        for h in np.geomspace(2.0, 100.0, 8):
            for inc in np.linspace(5.0, 80.0, 4):
                p = base.copy()
                p[0], p[2], p[3], p[12] = h, a, inc, 1
                rows.append(p)
    return np.array(rows)


def test_config2_relxill_1024_random(rx, oracle):
    """BASELINE config 2: relxill, 1024 random parameter vectors (seed 1234); a sample against the oracle, the
    whole batch for status / finiteness / batch-order invariance."""
    e = default_grid(3000)
    P = sample_params("relxill", 1024, seed=1234)
    f, st = rx.batch_eval("relxill", e, P, return_status=True)
    assert (st == 0).all() and np.isfinite(f).all()
    for i in (0, 17, 333, 1023):
        assert relerr(f[i], oracle.eval("relxill", e, P[i])) < RTOL
    np.testing.assert_array_equal(rx.batch_eval("relxill", e, P[::-1].copy())[::-1], f)


def test_config4_cp_batch_streams_through_chunks(rx, oracle):
    """BASELINE config 4 (relxillCp / relxilllpCp, uniform-random, one GPU's shard streamed through the scratch arena
    in several chunks): chunking must not change a bit, and sampled rows agree with the oracle."""
    e = default_grid(1000)
    for model, n in (("relxillCp", 5000), ("relxilllpCp", 2500)):
        P = sample_params(model, n, seed=99)
        if model == "relxilllpCp":
            P[:, 14] = 0          # iongrad_type 0, 10 zones (config 4)
        rx.set_cache(False)
        try:
            f, st = rx.batch_eval(model, e, P, return_status=True)
            ok = st == 0
            assert ok.mean() > 0.95 and np.isfinite(f).all()
            pick = [i for i in (1, n // 3, n // 2 + 7, n - 2) if ok[i]]
            for i in pick:
                assert relerr(f[i], oracle.eval(model, e, P[i])) < RTOL, (model, i)
            # the same vectors in small batches (one chunk each)
            sub = slice(n - 300, n)
            np.testing.assert_array_equal(rx.batch_eval(model, e, P[sub]), f[sub])
        finally:
            rx.set_cache(True)


def test_config5_returning_radiation_sweep(rx, oracle):
    """BASELINE config 5: relxilllp with returning radiation on a spin x height x inclination grid."""
    e = default_grid(1500)
    P = _config5_grid(rx)
    f, st = rx.batch_eval("relxilllp", e, P, return_status=True)
    assert (st == 0).all() and np.isfinite(f).all() and (f.sum(axis=1) > 0).all()
    for i in (0, 37, 101, 200, 255):
        assert relerr(f[i], oracle.eval("relxilllp", e, P[i])) < RTOL, i


# ---------------------------------------------------------------- full batches against the unmodified reference
@pytest.fixture(scope="module")
def refpool_factory(table_dir):
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref/librelxill_ref.so not built")
    from refpool import RefPool
    pools = []

    def make(zones):
        pools.append(RefPool(table_dir, zones))
        return pools[-1]
    yield make
    for p in pools:
        p.close()


def _census_report(name, c):
    import json
    print(f"\nPARITY-CENSUS {name}: {json.dumps(c)}")
    out = os.environ.get("RELXILL_B200_CENSUS_OUT")
    if out:
        with open(out, "a") as f:
            f.write(json.dumps({"case": name, **c}) + "\n")


@pytest.mark.parametrize("model", ["relxilllp", "relxilllpCp"])
def test_metric_batch_every_row_vs_reference(rx, refpool_factory, model):
    """BASELINE config 3, all of it: every one of the 4096 MCMC walkers x 50 zones x 3000 bins against the UNMODIFIED
    reference (oracle/_ref in a process pool on the host cores) — relxilllp, and relxilllpCp with the ionisation gradient
    (iongrad_type 1 and 2 alternating).  SURVEY App. C.16 predicts a small rate of single-bin outliers from discrete
    decisions that flip on a last-bit difference (Romberg exit, bin index, g* bracket): the census counts them.  No bin may
    exceed the north_star tolerance."""
    from refpool import census
    e = default_grid(3000)
    P = walker_ball(model, 4096)
    if model == "relxilllpCp":
        P[:, 14] = 1 + (np.arange(4096) % 2)
    rx.set_num_zones(50)
    try:
        got, st = rx.batch_eval(model, e, P, return_status=True)
    finally:
        rx.set_num_zones(None)
    assert (st == 0).all()
    want, dt = refpool_factory(50).eval_rows(model, e, P)
    c = census(got, want)
    c["reference_seconds"] = round(dt, 1)
    _census_report(f"cfg3_{model}_4096x50", c)
    assert c["n_bins_over_1e-05"] == 0, c


def test_config4_shard_vs_reference(rx, refpool_factory):
    """BASELINE config 4 at its per-GPU size: 8192 uniform-random relxillCp vectors and 8192 relxilllpCp vectors
    (iongrad_type 0, 10 zones) — one GPU's shard of the 65536 — with 256 randomly picked rows of each against the
    unmodified reference."""
    from refpool import census
    e = default_grid(3000)
    pool = refpool_factory(None)
    for model in ("relxillCp", "relxilllpCp"):
        P = sample_params(model, 8192, seed=99)
        if model == "relxilllpCp":
            P[:, 14] = 0
        got, st = rx.batch_eval(model, e, P, return_status=True)
        assert np.isfinite(got).all()
        ok = np.flatnonzero(st == 0)
        assert ok.size > 0.95 * 8192
        pick = np.sort(np.random.default_rng(4).choice(ok, 256, replace=False))
        want, _ = pool.eval_rows(model, e, P[pick])
        c = census(got[pick], want)
        _census_report(f"cfg4_{model}_8192_pick256", c)
        assert c["n_bins_over_1e-05"] == 0, (model, c)


def config5_grid_full(rx):
    """BASELINE config 5 as stated: a in linspace(0, 0.998, 32) x h in geomspace(2, 100, 32) x Incl in linspace(5, 80, 16)."""
    base = rx.default_params("relxilllp")
    a, h, inc = np.meshgrid(np.linspace(0.0, 0.998, 32), np.geomspace(2.0, 100.0, 32), np.linspace(5.0, 80.0, 16), indexing="ij")
    P = np.tile(base, (a.size, 1))
    P[:, 2], P[:, 0], P[:, 3], P[:, 12] = a.ravel(), h.ravel(), inc.ravel(), 1
    return P


def test_config5_full_sweep_vs_reference(rx, refpool_factory):
    """BASELINE config 5 at its stated size: the 32 x 32 x 16 = 16384-point returning-radiation sweep, 256 randomly
    picked grid points against the unmodified reference."""
    from refpool import census
    e = default_grid(3000)
    P = config5_grid_full(rx)
    assert P.shape[0] == 16384
    got, st = rx.batch_eval("relxilllp", e, P, return_status=True)
    assert (st == 0).all() and np.isfinite(got).all() and (got.sum(axis=1) > 0).all()
    pick = np.sort(np.random.default_rng(5).choice(16384, 256, replace=False))
    want, _ = refpool_factory(None).eval_rows("relxilllp", e, P[pick])
    c = census(got[pick], want)
    _census_report("cfg5_relxilllp_16384_pick256", c)
    assert c["n_bins_over_1e-05"] == 0, c


# ---------------------------------------------------------------- edge-case parameters and environment switches
@pytest.mark.parametrize("model", ["relxilllp", "relxilllpCp", "relline_lp", "relconv_lp", "relxill", "relline", "relxillCp"])
def test_edge_case_parameters(rx, oracle, model):
    """switch_returnrad in {-1, 2, 0}, positive Rin / Rout, negative h (x r+), negative / out-of-range Rbr
    (src/Rellp.cpp:483-509, src/ModelDefinition.cpp:202-226; tests/common.py: edge_params); the oracle is pinned to the
    reference on the same vectors in tests/test_oracle.py."""
    e = default_grid(1500)
    P = edge_params(model)
    fin = _conv_input(e) if model.startswith("relconv") else None
    rx.set_num_zones(None)
    oracle.set_num_zones(None)
    got, st = rx.batch_eval(model, e, P, fin, return_status=True)
    assert (st == 0).all(), st
    for a, p in zip(got, P):
        b = oracle.eval_conv(model, e, p, fin) if fin is not None else oracle.eval(model, e, p)
        assert relerr(a, b) < RTOL, (model, list(p))


@pytest.mark.parametrize("env,model", [({"RELXILL_RENORMALIZE": "1"}, "relxilllp"), ({"RELXILL_RENORMALIZE": "1"}, "relxillCp"),
                                       ({"RELXILL_RENORMALIZE": "1"}, "relxilllpCp"),
                                       ({"RELLINE_PHYSICAL_NORM": "1"}, "relline"), ({"RELLINE_PHYSICAL_NORM": "1"}, "relconv"),
                                       ({"RELLINE_PHYSICAL_NORM": "1"}, "relxill"),
                                       ({"RELXILL_RETURNRAD_SWITCH": "1"}, "relline_lp"), ({"RELXILL_RETURNRAD_SWITCH": "0"}, "relxilllp"),
                                       ({"RELXILL_NUM_RZONES": "17"}, "relxilllp"), ({"RELXILL_NUM_RZONES": "60"}, "relxilllp")])
def test_environment_switches_are_read_per_call(rx, oracle, monkeypatch, env, model):
    """The reference re-reads its environment switches on every evaluation (src/relutility.c:372-396,506-544,
    src/ModelDefinition.cpp:123-149, src/Relxill.cpp:241-278) — a pyxspec session flips them between calls.  Evaluate,
    set the variable, evaluate again (the retained batch of the first call must not leak its state), unset, evaluate
    again: each result must be the oracle's under the same environment (the oracle is pinned to the reference under
    these variables in tests/test_oracle.py)."""
    e = default_grid(1200)
    P = sample_params(model, 3, seed=77)
    fin = _conv_input(e) if model == "relconv" else None
    rx.set_num_zones(None)
    oracle.set_num_zones(None)

    def both():
        got = rx.batch_eval(model, e, P, fin)
        want = [oracle.eval_conv(model, e, p, fin) if fin is not None else oracle.eval(model, e, p) for p in P]
        assert max(relerr(a, b) for a, b in zip(got, want)) < RTOL, (env, model)
        return got
    base = both()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    switched = both()
    for k in env:
        monkeypatch.delenv(k)
    np.testing.assert_array_equal(both(), base)
    if env in ({"RELXILL_RENORMALIZE": "1"}, {"RELLINE_PHYSICAL_NORM": "1"}, {"RELXILL_NUM_RZONES": "17"}):
        assert max(relerr(a, b) for a, b in zip(switched, base)) > 1e-6      # the switch did something


def test_num_zones_env_between_two_lmod_calls(rx, oracle, monkeypatch):
    """RELXILL_NUM_RZONES flipped between two calls of the XSPEC symbol."""
    e = default_grid(600)
    p = rx.default_params("relxilllp")
    rx.set_num_zones(None)
    oracle.set_num_zones(None)
    a10 = rx.lmod("relxilllp", e, p)
    monkeypatch.setenv("RELXILL_NUM_RZONES", "35")
    a35 = rx.lmod("relxilllp", e, p)
    assert relerr(a35, oracle.eval("relxilllp", e, p)) < RTOL and relerr(a35, a10) > 1e-7
    monkeypatch.delenv("RELXILL_NUM_RZONES")
    np.testing.assert_array_equal(rx.lmod("relxilllp", e, p), a10)


@pytest.mark.parametrize("zones", [2, 3, 8, 9])
def test_few_zones_split_into_runs(rx, oracle, zones):
    """Vectors with <= 8 zones have every zone's radii cut into runs (one CTA each) whose partial line profiles
    k_linemerge adds in order (line.cu: line_parts); 9 zones is the first count that is not split."""
    e = default_grid(900)
    P = sample_params("relxilllp", 6, seed=300 + zones)
    rx.set_num_zones(zones)
    oracle.set_num_zones(zones)
    try:
        f, st = rx.batch_eval("relxilllp", e, P, return_status=True)
        assert (st == 0).all()
        for i in range(len(P)):
            assert relerr(f[i], oracle.eval("relxilllp", e, P[i])) < RTOL, (zones, i)
        # the zone profiles themselves, against the oracle's
        import torch
        b = rx.Batch("relxilllp", e, P[:2], keep_intermediates=True)
        out = torch.zeros((2, e.size - 1), dtype=torch.float64, device="cuda")
        b.run(out.data_ptr())
        torch.cuda.synchronize()
        for i in range(2):
            assert relerr(b.probe(i, "relflux"), oracle.stages("relxilllp", P[i])["relflux"]) < 1e-11, (zones, i)
        b.close()
    finally:
        rx.set_num_zones(None)
        oracle.set_num_zones(None)


def test_mixed_zone_counts_in_one_batch(rx, oracle, monkeypatch):
    """relxilllpCp with RELXILL_NUM_RZONES=5: constant-ionisation vectors get 5 zones (split into runs), vectors with an
    ionisation gradient ignore a value below 10 and get 25 (not split) — both kinds in one launch."""
    e = default_grid(700)
    P = sample_params("relxilllpCp", 8, seed=41)
    names = rx.PARAM_NAMES["relxilllpCp"]
    P[:, names.index("iongrad_type")] = [0, 1, 0, 2, 0, 1, 0, 0]
    rx.set_num_zones(None)
    oracle.set_num_zones(None)
    monkeypatch.setenv("RELXILL_NUM_RZONES", "5")
    try:
        f, st = rx.batch_eval("relxilllpCp", e, P, return_status=True)
        ok = st == 0
        assert ok.sum() >= 6
        for i in np.flatnonzero(ok):
            assert relerr(f[i], oracle.eval("relxilllpCp", e, P[i])) < RTOL, i
        # the same vectors one by one: a spectrum does not depend on its batch
        for i in np.flatnonzero(ok)[:4]:
            np.testing.assert_array_equal(rx.batch_eval("relxilllpCp", e, P[i : i + 1])[0], f[i])
    finally:
        monkeypatch.delenv("RELXILL_NUM_RZONES")


def test_fine_grid_around_the_line_fills_the_deep_queue(rx, oracle):
    """A line model on a grid much finer than the profile's structure: hundreds of Romberg bins per radius, many of them
    past level 2, so the queue of k_line fills inside a tile (the resume path) and zones are wider than any tile."""
    e = np.linspace(0.25, 1.45, 9001) * 6.4
    P = sample_params("relline", 5, seed=71)
    names = rx.PARAM_NAMES["relline"]
    P[:, names.index("lineE")] = 6.4
    P[:, names.index("z")] = [0.0, 0.0, 0.02, 0.0, 0.1]
    P[:, names.index("limb")] = [0, 0, 0, 1, 2]
    f, st = rx.batch_eval("relline", e, P, return_status=True)
    assert (st == 0).all()
    for i in range(len(P)):
        want = oracle.eval("relline", e, P[i])
        assert relerr(f[i], want) < RTOL, i
        assert abs(f[i].sum() / want.sum() - 1) < 1e-9
    # relline_lp through the same path
    P2 = sample_params("relline_lp", 3, seed=72)
    P2[:, rx.PARAM_NAMES["relline_lp"].index("lineE")] = 6.4
    f2 = rx.batch_eval("relline_lp", e, P2)
    for i in range(len(P2)):
        assert relerr(f2[i], oracle.eval("relline_lp", e, P2[i])) < RTOL, i


@pytest.mark.parametrize("model,zones", [("relxilllp", 50), ("relxilllpCp", None), ("relxill", None)])
def test_pipelined_host_call_equals_the_resident_run(rx, model, zones):
    """A host-buffer call above the pipeline threshold runs everything up to the zone spectra on the whole batch and cuts
    only the convolution into pieces (api.cu: split_tail); the result must be the device-resident run's, bit for bit."""
    import torch
    e = default_grid(400)
    n = 3100
    P = walker_ball(model, n, seed=5) if model != "relxill" else sample_params(model, n, seed=5)
    rx.set_cache(False)
    rx.set_num_zones(zones)
    try:
        b = rx.Batch(model, e, P)
        out = torch.zeros((n, e.size - 1), dtype=torch.float64, device="cuda")
        b.run(out.data_ptr())
        torch.cuda.synchronize()
        want = out.cpu().numpy()
        b.close()
        got, st = rx.batch_eval(model, e, P, return_status=True)
        assert (st == 0).mean() > 0.9
        np.testing.assert_array_equal(got, want)
    finally:
        rx.set_cache(True)
        rx.set_num_zones(None)
