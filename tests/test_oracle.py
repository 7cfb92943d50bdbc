"""CPU tests of the oracle: reference known-answer tests, golden vectors generated from the unmodified
reference, and (when oracle/_ref is built) fresh reference runs, whole-model and stage by stage."""
import os

import numpy as np
import pytest

from common import ALL_MODELS, MODELS, NSCO_MODELS, default_grid, relerr, sample_params

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
GOLDEN_NSCO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2_nsco.npz")
GOLDEN_CFG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v3_configs.npz")


# ---------------------------------------------------------------- table-independent KATs of the reference
def test_rebin_spectrum_kat(oracle):
    # reference test/unit/test-stdfunctions.cpp:129-152
    ener0 = np.array([1, 3, 4, 5, 6, 7, 9], float)
    val0 = np.array([1, 1, 1, 2, 1, 1], float)
    ener = np.array([0.5, 2, 4, 5.5, 7.5, 8], float)
    val = oracle.rebin(ener, ener0, val0)
    np.testing.assert_allclose(val, [0.5, 1.5, 2.0, 2.25, 0.25], atol=1e-12)


def test_rebin_commutes_with_the_table_blend(oracle):
    """What the convolution-grid copy of the xillver tables rests on (relxill_b200/csrc/tables.cu, xill.cu): _rebin_spectrum
    (src/relutility.c:549-601) is a fixed linear map, so rebinning a weighted sum of table rows (what the reference
    does per zone, src/Relxill.cpp:461-463) equals the weighted sum of the rebinned rows, to rounding."""
    rng = np.random.default_rng(3)
    src = np.exp(np.linspace(np.log(0.07), np.log(1000.1), 3000))     # a grid like the xillver tables' (2999 bins)
    dst = oracle.conv_grid()
    rows = rng.uniform(0.0, 1.0, (16, src.size - 1)).astype(np.float32).astype(np.float64) * src[:-1] ** -1.5
    w = rng.uniform(0.0, 1.0, 16)
    w /= w.sum()
    blend_then_rebin = oracle.rebin(dst, src, w @ rows)
    rebin_then_blend = w @ np.array([oracle.rebin(dst, src, r) for r in rows])
    m = blend_then_rebin > 0
    assert m.sum() > 2400 and not rebin_then_blend[~m].any()          # only the bins overlapping the table grid are non-zero
    np.testing.assert_allclose(rebin_then_blend[m], blend_then_rebin[m], rtol=1e-13)
    # and the rebin conserves the flux of the overlapping part
    assert abs(blend_then_rebin.sum() / (w @ rows).sum() - 1) < 1e-12


def test_interp_2d_float_kat(oracle):
    # reference test/unit/test-stdfunctions.cpp:192-211
    assert abs(oracle.lib.orc_lin2d_float(0.4, 0.8, 1.0, 2.0, 2.0, 4.0) - 2.52) < 1e-6


def test_gshift_fluxboost_factor_kat(oracle):
    # reference test/unit/tests-returnrad.cpp:479-497 (corrected_gshift_fluxboost_factor)
    f, gamma = oracle.lib.orc_gshift_fluxboost, 2.0
    assert f(1.2, 1.5, gamma) > 1.5 ** gamma and f(1.2, 0.3, gamma) > 0.3 ** gamma
    assert f(1.2, 1.5, gamma) > 1 and 0 < f(1.2, 0.3, gamma) < 1
    assert f(0.9, 1.5, gamma) < 1.5 ** gamma and f(0.9, 0.3, gamma) < 0.3 ** gamma
    assert f(0.9, 1.5, gamma) > 1 and 0 < f(0.9, 0.3, gamma) < 1


def test_energy_grid_shifts(oracle):
    """XspecSpectrum::shift_energy_grid_1keV / _redshift (src/XspecSpectrum.h:61-76; reference test
    test/unit/test-cppspectrum.cpp:57-77), seen from outside: a line at lineE on the grid E is the line at 1 keV on
    E / lineE, and a redshift z moves the grid to E (1 + z)."""
    e = default_grid(800, 0.2, 20.0)
    p = oracle.default_params("relline")
    p[0], p[8] = 6.4, 0.0
    base = oracle.eval("relline", e, p)
    q = p.copy()
    q[0] = 3.2
    np.testing.assert_allclose(oracle.eval("relline", e / 2.0, q), base, rtol=1e-10, atol=1e-14)   # lineE halves with the grid
    q = p.copy()
    q[8] = 0.5
    np.testing.assert_allclose(oracle.eval("relline", e / 1.5, q), base, rtol=1e-10, atol=1e-14)   # observed grid = rest grid / (1 + z)


def test_default_grid_endpoints():
    # reference test/unit/tests-execmodel.cpp:44-61
    e = default_grid(100, 0.5, 10.0)
    assert e[0] == 0.5 and e[-1] == 10.0 and e[1] > e[0] and e.size == 101


def test_kerr_rms(oracle):
    assert abs(oracle.lib.orc_kerr_rms(0.0) - 6.0) < 1e-12
    assert abs(oracle.lib.orc_kerr_rms(0.998) - 1.2369706551751847) < 1e-9
    assert abs(oracle.lib.orc_kerr_rms(-0.998) - 8.994376) < 1e-5


def test_fft_convolution_normalisation(oracle):
    # reference test/unit/test-stdfunctions.cpp:253-298: the convolution keeps the 0.01-1000 keV sum
    e = oracle.conv_grid()
    emid = 0.5 * (e[1:] + e[:-1])
    xill = emid ** -2.0 * np.exp(-emid / 300.0) * np.diff(e)
    xill[(e[:-1] < 0.08) | (e[1:] > 900)] = 0.0
    rel = np.exp(-0.5 * ((np.log(emid) - np.log(0.9)) / 0.1) ** 2)
    band = (e[:-1] >= 0.01) & (e[1:] < 1000.0)
    rel /= rel[band].sum()
    out = oracle.fft_conv(xill, rel)
    assert abs(out[band].sum() - xill[band].sum()) < 1e-8 * xill[band].sum()
    rel2 = rel.copy()
    rel2[1000] = 1000.0
    out2 = oracle.fft_conv(xill, rel2)
    assert abs(out2[band].sum() - xill[band].sum()) > 1e-8 * xill[band].sum()


def test_model_database(oracle):
    counts = dict(relline=10, relconv=8, relline_lp=10, relconv_lp=9, relxill=13, relxilllp=14, relxillCp=14,
                  relxilllpCp=17)
    for m, n in counts.items():
        assert oracle.num_params(m) == n
    assert oracle.default_params("relxilllp")[10] == 300.0


# ---------------------------------------------------------------- golden vectors (from the unmodified reference)
@pytest.mark.parametrize("model", ALL_MODELS)
def test_oracle_vs_golden(oracle, model):
    g = np.load(GOLDEN_NSCO if model in NSCO_MODELS else GOLDEN)
    e = g["energy"]
    P, F = g[f"{model}_params"], g[f"{model}_flux"]
    oracle.set_num_zones(None)
    for p, f in zip(P, F):
        if model.startswith("relconv"):
            got = oracle.eval_conv(model, e, p, g["conv_input"])
        else:
            got = oracle.eval(model, e, p)
        assert relerr(got, f) < 1e-8, (model, p)


@pytest.mark.parametrize("model", ["relxilllp", "relxilllpCp"])
def test_oracle_vs_golden_50_zones(oracle, model):
    g = np.load(GOLDEN)
    oracle.set_num_zones(50)
    try:
        for p, f in zip(g[f"{model}_z50_params"], g[f"{model}_z50_flux"]):
            assert relerr(oracle.eval(model, g["energy"], p), f) < 1e-8
    finally:
        oracle.set_num_zones(None)


@pytest.mark.parametrize("key,model,zones", [("cfg2_relxill", "relxill", None), ("cfg3_relxilllp", "relxilllp", 50),
                                              ("cfg3_relxilllpCp", "relxilllpCp", 50), ("cfg4_relxillCp", "relxillCp", None),
                                              ("cfg4_relxilllpCp", "relxilllpCp", None), ("cfg5_relxilllp", "relxilllp", None)])
def test_oracle_vs_golden_baseline_configs(oracle, key, model, zones):
    """Rows of the BASELINE.json configurations 2-5 (random relxill vectors, 50-zone MCMC walkers, the Cp shards, the
    returning-radiation sweep) evaluated by the unmodified reference: tests/golden/make_golden.py configs."""
    g = np.load(GOLDEN_CFG)
    oracle.set_num_zones(zones)
    try:
        for p, f in zip(g[f"{key}_params"], g[f"{key}_flux"]):
            assert relerr(oracle.eval(model, g["energy"], p), f) < 1e-8, (key, p)
    finally:
        oracle.set_num_zones(None)


def test_golden_tables_match(table_dir):
    import hashlib
    from relxill_b200.tables import synth
    h = hashlib.sha256()
    for key in ("rel", "lp", "xill", "xillcp", "rrad"):
        with open(os.path.join(table_dir, synth.FILES[key]), "rb") as f:
            h.update(f.read())
    assert h.hexdigest() == str(np.load(GOLDEN)["table_digest"]), "synthetic tables changed: regenerate the golden file"
    assert h.hexdigest() == str(np.load(GOLDEN_CFG)["table_digest"]), "synthetic tables changed: regenerate golden_v3_configs.npz"
    h = hashlib.sha256()
    for key in ("rel", "xillns", "xillco"):
        with open(os.path.join(table_dir, synth.FILES[key]), "rb") as f:
            h.update(f.read())
    assert h.hexdigest() == str(np.load(GOLDEN_NSCO)["table_digest"]), "synthetic NS/CO tables changed: regenerate golden_v2_nsco.npz"


# ---------------------------------------------------------------- fresh runs of the unmodified reference
@pytest.mark.parametrize("model", ALL_MODELS)
def test_oracle_vs_reference_random(oracle, ref, model):
    e = default_grid(1200)
    P = sample_params(model, 3, seed=31 + len(model))
    conv_in = np.exp(-0.5 * ((np.log(0.5 * (e[1:] + e[:-1])) - np.log(6.4)) / 0.03) ** 2) + 1e-3
    ref.set_num_zones(None)
    oracle.set_num_zones(None)
    for p in P:
        if model.startswith("relconv"):
            a, b = oracle.eval_conv(model, e, p, conv_in), ref.eval_conv(model, e, p, conv_in)
        else:
            a, b = oracle.eval(model, e, p), ref.eval(model, e, p)
        assert relerr(a, b) < 1e-8, (model, p)


@pytest.mark.parametrize("model", ["relxill", "relxilllp", "relxilllpCp"])
def test_oracle_stages_vs_reference(oracle, ref, model):
    ref.set_num_zones(None)
    oracle.set_num_zones(None)
    for p in sample_params(model, 2, seed=5):
        a, b = oracle.stages(model, p), ref.stages(model, p)
        assert a["nz"] == b["nz"]
        np.testing.assert_array_equal(a["zone"], b["zone"])
        for k in ("lxi", "dens", "ect", "eshift", "normch", "corr_flux", "corr_gshift", "emis2", "dist"):
            np.testing.assert_allclose(a[k], b[k], rtol=1e-11, atol=0, err_msg=k)
        assert relerr(a["relflux"], b["relflux"]) < 1e-11
        assert relerr(a["xill"], b["xill"]) < 1e-11
        assert relerr(a["total"], b["total"]) < 1e-8
        sa, sb = oracle.syspar(model, p), ref.syspar(model, p)
        for k in ("re", "gmin", "gmax", "trff", "cosne", "del_emit", "del_inc"):
            np.testing.assert_allclose(sa[k], sb[k], rtol=1e-13, atol=0, err_msg=k)


def test_nthcomp_vs_reference(oracle, ref):
    e = default_grid(500)
    for gam, kte, z in [(2.0, 60.0, 0.0), (1.4, 5.0, 0.3), (3.2, 350.0, 1.5)]:
        np.testing.assert_allclose(oracle.nthcomp(e, gam, kte, z), ref.nthcomp(e, gam, kte, z), rtol=1e-12)


def test_constant_density_env_vs_reference(oracle, ref, monkeypatch):
    """RELXILL_CONSTANT_DENSITY=1 (src/relutility.c:372-382): alpha-disk gradient at constant density."""
    e = default_grid(800)
    p = sample_params("relxilllpCp", 1, seed=3)[0]
    p[14] = 2
    ref.set_num_zones(None)
    oracle.set_num_zones(None)
    base = oracle.eval("relxilllpCp", e, p)
    ref.close()
    monkeypatch.setenv("RELXILL_CONSTANT_DENSITY", "1")
    try:
        a, b = oracle.eval("relxilllpCp", e, p), ref.eval("relxilllpCp", e, p)
    finally:
        ref.close()                     # workers spawned with the switch set must not outlive this test
    assert relerr(a, b) < 1e-8
    assert relerr(a, base) > 1e-4


# ---------------------------------------------------------------- environment switches and edge-case parameters
def _fresh_ref(table_dir):
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref/librelxill_ref.so not built")
    return pyref.Ref(table_dir)      # workers are spawned on first use and inherit os.environ as it is then


@pytest.mark.parametrize("env,model", [({"RELXILL_RENORMALIZE": "1"}, "relxilllp"), ({"RELXILL_RENORMALIZE": "1"}, "relxillCp"),
                                       ({"RELLINE_PHYSICAL_NORM": "1"}, "relline"), ({"RELLINE_PHYSICAL_NORM": "1"}, "relconv"),
                                       ({"RELLINE_PHYSICAL_NORM": "1"}, "relxill"),
                                       ({"RELXILL_RETURNRAD_SWITCH": "1"}, "relline_lp"), ({"RELXILL_RETURNRAD_SWITCH": "0"}, "relxilllp"),
                                       ({"RELXILL_NUM_RZONES": "17"}, "relxilllp"), ({"RELXILL_NUM_RZONES": "60"}, "relxilllp"),
                                       ({"RELXILL_NUM_RZONES": "7"}, "relxilllpCp")])
def test_env_switches_vs_reference(oracle, table_dir, monkeypatch, env, model):
    """The reference reads RELXILL_RENORMALIZE (src/Relxill.cpp:241-278), RELLINE_PHYSICAL_NORM (src/relutility.c:386-396),
    RELXILL_RETURNRAD_SWITCH (src/ModelDefinition.cpp:123-149) and RELXILL_NUM_RZONES (src/relutility.c:506-544) on every
    evaluation; the oracle must do the same, including the out-of-range zone counts that fall back to the defaults."""
    e = default_grid(500)
    P = sample_params(model, 2, seed=77)
    if model == "relxilllpCp":
        P[:, 14] = [1, 2]     # ionisation gradient: zone counts below 10 are refused
    fin = np.exp(-0.5 * ((np.log(0.5 * (e[1:] + e[:-1])) - np.log(6.4)) / 0.03) ** 2) + 1e-3
    oracle.set_num_zones(None)
    base = [oracle.eval_conv(model, e, p, fin) if model == "relconv" else oracle.eval(model, e, p) for p in P]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ref = _fresh_ref(table_dir)
    try:
        changed = False
        for p, b in zip(P, base):
            got = oracle.eval_conv(model, e, p, fin) if model == "relconv" else oracle.eval(model, e, p)
            want = ref.eval_conv(model, e, p, fin) if model == "relconv" else ref.eval(model, e, p)
            assert relerr(got, want) < 1e-8, (env, model)
            changed |= relerr(got, b) > 1e-6
        expect_change = env not in ({"RELXILL_RETURNRAD_SWITCH": "0"}, {"RELXILL_NUM_RZONES": "60"}, {"RELXILL_NUM_RZONES": "7"})
        if model == "relline_lp":
            expect_change = False     # its lmodel.dat entry carries switch_returnrad, which takes precedence over the environment
        assert changed == expect_change, (env, model)
    finally:
        ref.close()


@pytest.mark.parametrize("model", ["relxilllp", "relxilllpCp", "relline_lp", "relxill", "relline", "relxillCp"])
def test_edge_case_parameters_vs_reference(oracle, ref, model):
    """switch_returnrad -1 / 2, positive Rin / Rout, negative h and Rbr (tests/common.py: edge_params)."""
    from common import edge_params
    e = default_grid(500)
    oracle.set_num_zones(None)
    for p in edge_params(model):
        assert relerr(oracle.eval(model, e, p), ref.eval(model, e, p)) < 1e-8, (model, list(p))
