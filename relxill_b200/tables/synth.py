"""Synthetic relxill tables in the exact FITS layout the reference reads.

The published tables (rel_table_v0.5a.fits, rel_lp_table_v0.5b.fits,
xillver-a-Ec5.fits, xillverCp_v3.4.fits, table_returnRad_v20220301.fits) are
not available offline, so smooth, physically-plausible stand-ins are generated
with fixed formulas (no RNG: the same call always gives the same bytes).  Both
the reference (oracle/_ref) and the CUDA library read these same files.

Layouts follow the reference readers:
  rel table   : reference src/reltable.c:171-311  (HDU number = ia*30+imu+4)
  lp table    : reference src/reltable.c:313-448
  xillver     : reference src/xilltable.c:169-276,513-564
  return rad  : reference src/Relreturn_Table.cpp:289-361

Usage:  python -m relxill_b200.tables.synth OUTDIR [--size test|bench]
"""
from __future__ import annotations

import argparse
import json
import os

import numpy as np

from .fitsmin import Column, FitsWriter

REL_NA, REL_NMU0, REL_NR, REL_NG = 25, 30, 100, 40
LP_NA, LP_NH, LP_NR = 20, 250, 100
RR_NR, RR_NG = 50, 20

SIZES = {
    # name: (xillver 5-D axes, xillverCp 6-D axes, n_ener)
    "test": dict(gam=5, afe=3, lxi=5, ect=5, kte=4, dens=3, gam_cp=4, n_ener=2999),
    "bench": dict(gam=13, afe=4, lxi=15, ect=11, kte=10, dens=6, gam_cp=12, n_ener=2999),
}
# the neutron-star (blackbody-irradiated) and CO flavours: same sizes for "test" and "bench"
SIZES_NSCO = {
    "ns": dict(ktbb=4, afe=3, lxi=4, dens=3),
    "co": dict(gam=3, aco=3, ktbb=3, frac=3, ect=3),
}


def kerr_rms(a):
    a = np.asarray(a, dtype=np.float64)
    sign = np.where(a < 0, -1.0, 1.0)
    z1 = 1.0 + np.cbrt(1.0 - a * a) * (np.cbrt(1.0 + a) + np.cbrt(1.0 - a))
    z2 = np.sqrt(3.0 * a * a + z1 * z1)
    return 3.0 + z2 - sign * np.sqrt((3.0 - z1) * (3.0 + z1 + 2 * z2))


def _spin_nodes(n, amin=-0.998, amax=0.9982):
    """Spin nodes uniform in ISCO radius: linear interpolation of r_ISCO between
    neighbours then stays close to the true ISCO (the reference snaps the
    innermost radius to the ISCO only if it is within 1e-3 relative)."""
    fine = np.linspace(amin, amax, 400001)
    risco = kerr_rms(fine)
    # refine towards a->1 where ISCO(a) bends most: uniform in sqrt(r_isco - 1)
    t = np.sqrt(risco - 1.0)
    targets = np.linspace(t[0], t[-1], n)
    nodes = np.interp(-targets, -t, fine)
    nodes[0], nodes[-1] = amin, amax
    return nodes


# --------------------------------------------------------------------------- rel table
def write_rel_table(path):
    a = _spin_nodes(REL_NA).astype(np.float32)
    mu0 = np.linspace(0.02, 1.0, REL_NMU0).astype(np.float32)
    w = FitsWriter(path)
    w.add_table("a", [Column("a", "E", a)])
    w.add_table("mu0", [Column("mu0", "E", mu0)])
    gst = 5e-3 + (1.0 - 1e-2) / (REL_NG - 1) * np.arange(REL_NG)
    k = np.arange(REL_NR) / (REL_NR - 1.0)
    for ia in range(REL_NA):
        aa = float(a[ia])
        risco = float(kerr_rms(aa))
        r = 1000.0 * (risco / 1000.0) ** k
        r[0], r[-1] = 1000.0, risco
        r32 = r.astype(np.float32)
        rr = r32.astype(np.float64)[:, None]
        for im in range(REL_NMU0):
            mu = float(mu0[im])
            si = np.sqrt(max(1.0 - mu * mu, 0.0))
            grav = np.sqrt(1.0 - 2.0 / (rr[:, 0] + 1.2))
            beta = 0.95 / np.sqrt(rr[:, 0] + 0.5) * (1.0 + 0.1 * aa)
            gmin = grav / (1.0 + 0.9 * beta * si) * 0.98
            gmax = grav / (1.0 - 0.9 * beta * si) * 1.02
            g = gst[None, :]
            bend = 1.0 + 2.0 / rr
            trff1 = (0.3 + 2.0 * g * (1 - g) + 0.5 * g) * (1 + 0.3 * mu) * bend * (1 + 0.1 * aa)
            trff2 = (0.35 + 1.5 * g * (1 - g) + 0.6 * (1 - g) ** 2) * (1 + 0.2 * mu * mu) * bend * (1 - 0.05 * aa)
            cos1 = np.clip(mu * (0.6 + 0.35 * g) + 0.2 / (1 + rr / 10.0) * (1 - mu) + 0.013, 0.05, 0.95)
            cos2 = np.clip(mu * (0.95 - 0.35 * g) + 0.1 / (1 + rr / 20.0) + 0.007, 0.05, 0.95)
            w.add_table(
                f"{ia + 1}_{im + 1}",
                [
                    Column("r", "E", r32),
                    Column("gmin", "E", gmin),
                    Column("gmax", "E", gmax),
                    Column("trff1", "E", trff1),
                    Column("trff2", "E", trff2),
                    Column("cosne1", "E", cos1),
                    Column("cosne2", "E", cos2),
                ],
            )
    w.close()


# --------------------------------------------------------------------------- lamp-post table
def write_lp_table(path):
    a = _spin_nodes(LP_NA).astype(np.float32)
    cols_h = []
    hgrid = np.zeros((LP_NA, LP_NH), np.float32)
    rad = np.zeros((LP_NA, LP_NR), np.float32)
    intens = np.zeros((LP_NA, LP_NH, LP_NR), np.float32)
    dele = np.zeros_like(intens)
    deli = np.zeros_like(intens)
    for ia in range(LP_NA):
        aa = float(a[ia])
        hg = np.geomspace(1.0 + 0.02 * (1 - aa), 520.0 + 10 * aa, LP_NH)
        rg = np.geomspace(float(kerr_rms(aa)), 1000.0, LP_NR)
        rg[-1] = 1000.0
        hgrid[ia] = hg
        rad[ia] = rg
        h = hgrid[ia].astype(np.float64)[:, None]
        r = rad[ia].astype(np.float64)[None, :]
        newton = (1.0 / ((r / h) ** 2 + 1.0)) ** 1.5 / (2 * np.pi * h * h)
        boost = (1.0 + 3.0 / (r * (1 + 0.2 * h))) * (1 + 0.05 * aa)
        intens[ia] = newton * boost
        dele[ia] = np.arctan2(r, h) * (1 - 0.05 * np.exp(-h / 5.0))
        # stored with a sign in the real table; the reader takes fabsf()
        deli[ia] = -(0.1 + 1.3 * (r / (r + h)))
    cols = [Column("a", "E", a), Column("hgrid", "E", hgrid), Column("r", "E", rad)]
    for k in range(LP_NH):
        cols.append(Column(f"h{k + 1}", "E", intens[:, k, :]))
        cols.append(Column(f"del{k + 1}", "E", dele[:, k, :]))
        cols.append(Column(f"del_inc{k + 1}", "E", deli[:, k, :]))
    w = FitsWriter(path)
    w.add_table("I_h", cols)
    w.close()


# --------------------------------------------------------------------------- xillver tables
def _xill_energy(n_ener):
    edges = np.geomspace(0.07, 1000.1, n_ener + 1).astype(np.float32)
    return edges[:-1].copy(), edges[1:].copy()


def _incl_axis():
    # 10 angles, uniform in cos(i): 0.95 ... 0.05 (what get_cosne_bin assumes)
    return np.degrees(np.arccos(0.95 - 0.1 * np.arange(10))).astype(np.float32)


def _gauss(e, e0, sig):
    return np.exp(-0.5 * ((e - e0) / sig) ** 2)


def _xill_spec(emid, de, mu, gam, afe, lxi, ecut, dens=15.0):
    """photons/bin, already in the 'renormalised' units (before x 10^lxi etc.);
    emid,de: (n_ener,), mu: (n_incl,) -> (n_incl, n_ener)"""
    xin = lxi / 4.7
    cont = emid ** (-gam) * np.exp(-emid / ecut)
    normc = 497.0 / np.sum(cont * de * emid * ((emid >= 0.1) & (emid <= 1000)))
    hump = (emid / 25.0) ** 2 / (1.0 + (emid / 25.0) ** 3.5)
    c64 = 6.4 ** (-gam) * np.exp(-6.4 / ecut)
    fe = 8.0 * afe * (1 - 0.7 * xin) * c64 * _gauss(emid, 6.4 + 0.3 * xin, 0.15)
    edge = 1.0 - 0.4 * afe / (1 + afe) * (emid > 7.1) * np.exp(-np.maximum(emid - 7.1, 0) / 5.0)
    soft = (1.2 - xin) * (
        3.0 * 0.65 ** (-gam) * _gauss(emid, 0.65, 0.02)
        + 2.0 * 0.9 ** (-gam) * _gauss(emid, 0.9, 0.03)
        + 1.0 * 1.8 ** (-gam) * _gauss(emid, 1.8, 0.05)
    )
    densx = 0.3 * (dens - 15.0) * np.exp(-emid / 0.5) * emid ** (-gam)
    s = cont * (0.2 + 0.8 * xin + 1.5 * hump) * edge + fe + soft * np.exp(-emid / ecut) + densx
    ang = (0.5 + mu[:, None]) * (1.0 + 0.2 * mu[:, None] * np.log(emid[None, :] + 1.0) / 7.0)
    return normc * s[None, :] * de[None, :] * ang


def _write_param_ext(w, names, vals):
    maxn = max(len(v) for v in vals)
    value = np.zeros((len(names), maxn), np.float32)
    for i, v in enumerate(vals):
        value[i, : len(v)] = v
    n = len(names)
    z = np.zeros(n, np.float32)
    w.add_table(
        "PARAMETERS",
        [
            Column("NAME", "A", [s.ljust(8) for s in names]),
            Column("METHOD", "J", np.zeros(n, np.int32)),
            Column("INITIAL", "E", [v[0] for v in vals]),
            Column("DELTA", "E", z),
            Column("MINIMUM", "E", [v[0] for v in vals]),
            Column("BOTTOM", "E", [v[0] for v in vals]),
            Column("TOP", "E", [v[-1] for v in vals]),
            Column("MAXIMUM", "E", [v[-1] for v in vals]),
            Column("NUMBVALS", "J", np.array([len(v) for v in vals], np.int32)),
            Column("VALUE", "E", value),
        ],
    )


def write_xillver(path, size, nthcomp=False):
    cfg = SIZES[size]
    n_ener = cfg["n_ener"]
    elo, ehi = _xill_energy(n_ener)
    incl = _incl_axis()
    mu = np.cos(np.radians(incl.astype(np.float64)))
    afe = np.geomspace(0.5, 10.0, cfg["afe"]).astype(np.float32)
    lxi = np.linspace(0.0, 4.7, cfg["lxi"]).astype(np.float32)
    if nthcomp:
        gam = np.linspace(1.2, 3.4, cfg["gam_cp"]).astype(np.float32)
        kte = np.geomspace(1.0, 400.0, cfg["kte"]).astype(np.float32)
        dens = np.linspace(15.0, 20.0, cfg["dens"]).astype(np.float32)
        names = ["Gamma", "A_Fe", "logXi", "kTe", "Dens", "Incl"]
        vals = [gam, afe, lxi, kte, dens, incl]
    else:
        gam = np.linspace(1.0, 3.4, cfg["gam"]).astype(np.float32)
        ect = np.geomspace(5.0, 1000.0, cfg["ect"]).astype(np.float32)
        ect[-1] = 1000.0
        names = ["Gamma", "A_Fe", "logXi", "Ecut", "Incl"]
        vals = [gam, afe, lxi, ect, incl]
    w = FitsWriter(path)
    _write_param_ext(w, names, vals)
    w.add_table("ENERGIES", [Column("ENERG_LO", "E", elo), Column("ENERG_HI", "E", ehi)])
    npar = len(names)
    shape = [len(v) for v in vals]
    nrows = int(np.prod(shape))
    w.begin_stream_table("SPECTRA", [("PARAMVAL", "E", npar), ("INTPSPEC", "E", n_ener)], nrows)
    e_lo = elo.astype(np.float64)
    e_hi = ehi.astype(np.float64)
    emid = 0.5 * (e_lo + e_hi)
    de = e_hi - e_lo
    n_incl = len(incl)
    for ig, g in enumerate(gam):
        for ia, af in enumerate(afe):
            for ix, lx in enumerate(lxi):
                if nthcomp:
                    blk = np.zeros((len(kte), len(dens), n_incl, npar + n_ener), ">f4")
                    for ik, kt in enumerate(kte):
                        for idn, dn in enumerate(dens):
                            sp = _xill_spec(emid, de, mu, float(g), float(af), float(lx), 2.5 * float(kt), float(dn))
                            sp = sp * 10.0 ** float(lx) * 10.0 ** (float(dn) - 15.0)
                            blk[ik, idn, :, npar:] = sp
                            blk[ik, idn, :, 0:5] = [g, af, lx, kt, dn]
                            blk[ik, idn, :, 5] = incl
                else:
                    blk = np.zeros((len(ect), n_incl, npar + n_ener), ">f4")
                    for ie, ec in enumerate(ect):
                        sp = _xill_spec(emid, de, mu, float(g), float(af), float(lx), float(ec))
                        sp = sp * 10.0 ** float(lx)
                        blk[ie, :, npar:] = sp
                        blk[ie, :, 0:4] = [g, af, lx, ec]
                        blk[ie, :, 4] = incl
                w.stream_rows(blk.reshape(-1, npar + n_ener))
    w.close()


def _bb_like(emid, ktbb):
    x = np.minimum(emid / ktbb, 600.0)
    return emid ** 2 / (ktbb ** 4 * np.expm1(x))


def write_xillver_nsco(path, flavour, n_ener=2999):
    """xillverNS-2.fits (5-D: kTbb, A_Fe, logXi, Dens, Incl; blackbody irradiation) and xillverCO.fits
    (6-D: Gamma, A_CO, kTbb, Frac, Ecut, Incl), reference src/xilltable.c:682-760, common.h:141-168."""
    cfg = SIZES_NSCO[flavour]
    elo, ehi = _xill_energy(n_ener)
    incl = _incl_axis()
    mu = np.cos(np.radians(incl.astype(np.float64)))
    e_lo, e_hi = elo.astype(np.float64), ehi.astype(np.float64)
    emid, de = 0.5 * (e_lo + e_hi), e_hi - e_lo
    n_incl = len(incl)
    if flavour == "ns":
        ktbb = np.geomspace(0.5, 10.0, cfg["ktbb"]).astype(np.float32)
        afe = np.geomspace(0.5, 10.0, cfg["afe"]).astype(np.float32)
        lxi = np.linspace(1.0, 4.7, cfg["lxi"]).astype(np.float32)
        dens = np.linspace(15.0, 19.0, cfg["dens"]).astype(np.float32)
        names, vals = ["kTbb", "A_Fe", "logXi", "Dens", "Incl"], [ktbb, afe, lxi, dens, incl]
    else:
        gam = np.linspace(1.0, 2.8, cfg["gam"]).astype(np.float32)
        aco = np.geomspace(1.0, 1000.0, cfg["aco"]).astype(np.float32)
        ktbb = np.geomspace(0.05, 0.5, cfg["ktbb"]).astype(np.float32)
        frac = np.geomspace(0.01, 1.0, cfg["frac"]).astype(np.float32)
        ect = np.geomspace(2.0, 1000.0, cfg["ect"]).astype(np.float32)
        names, vals = ["Gamma", "A_CO", "kTbb", "Frac", "Ecut", "Incl"], [gam, aco, ktbb, frac, ect, incl]
    w = FitsWriter(path)
    _write_param_ext(w, names, vals)
    w.add_table("ENERGIES", [Column("ENERG_LO", "E", elo), Column("ENERG_HI", "E", ehi)])
    npar = len(names)
    nrows = int(np.prod([len(v) for v in vals]))
    w.begin_stream_table("SPECTRA", [("PARAMVAL", "E", npar), ("INTPSPEC", "E", n_ener)], nrows)
    band = (emid >= 0.1) & (emid <= 1000)
    if flavour == "ns":
        for kt in ktbb:
            bb = _bb_like(emid, float(kt))
            for af in afe:
                for lx in lxi:
                    blk = np.zeros((len(dens), n_incl, npar + n_ener), ">f4")
                    for idn, dn in enumerate(dens):
                        sp = _xill_spec(emid, de, mu, 2.0, float(af), float(lx), 300.0, float(dn))
                        refl = bb * (0.3 + 0.5 * float(lx) / 4.7) * de * 497.0 / np.sum(bb * de * emid * band)
                        sp = 0.15 * sp + refl[None, :] * (0.5 + mu[:, None])
                        sp = sp * 10.0 ** float(lx) * 10.0 ** (float(dn) - 15.0)
                        blk[idn, :, npar:] = sp
                        blk[idn, :, 0:4] = [kt, af, lx, dn]
                        blk[idn, :, 4] = incl
                    w.stream_rows(blk.reshape(-1, npar + n_ener))
    else:
        for g in gam:
            for ac in aco:
                for kt in ktbb:
                    bb = _bb_like(emid, float(kt))
                    bbn = bb * de * 497.0 / np.sum(bb * de * emid * band)
                    blk = np.zeros((len(frac), len(ect), n_incl, npar + n_ener), ">f4")
                    for ifr, fr in enumerate(frac):
                        for ie, ec in enumerate(ect):
                            sp = _xill_spec(emid, de, mu, float(g), 1.0, 1.0, float(ec))
                            cline = 2.0 * np.log10(float(ac) + 1.0) * 0.28 ** (-float(g)) * _gauss(emid, 0.28, 0.02) \
                                + np.log10(float(ac) + 1.0) * 0.53 ** (-float(g)) * _gauss(emid, 0.53, 0.03)
                            mix = float(fr) * sp + (1.0 - float(fr)) * 0.3 * bbn[None, :] * (0.5 + mu[:, None])
                            sp = mix + 0.05 * cline[None, :] * de[None, :] * (0.5 + mu[:, None])
                            sp = sp * 100.0          # the table is computed for logN = 17 (src/ModelDefinition.cpp:362-363)
                            blk[ifr, ie, :, npar:] = sp
                            blk[ifr, ie, :, 0:5] = [g, ac, kt, fr, ec]
                            blk[ifr, ie, :, 5] = incl
                    w.stream_rows(blk.reshape(-1, npar + n_ener))
    w.close()


# --------------------------------------------------------------------------- returning radiation
RR_SPINS = np.array([-0.5, 0.0, 0.5, 0.8, 0.9, 0.95, 0.99, 0.9982])


def write_rrad_table(path):
    w = FitsWriter(path)
    w.add_table("SPIN", [Column("a", "D", RR_SPINS)])
    for isp, aa in enumerate(RR_SPINS):
        edges = np.geomspace(float(kerr_rms(aa)), 1000.0, RR_NR + 1)
        edges[-1] = 1000.0
        rlo, rhi = edges[:-1], edges[1:]
        rm = 0.5 * (rlo + rhi)
        ri = rm[:, None]  # incident
        re = rm[None, :]  # emitted
        dl = np.abs(np.log(ri / re))
        tf_r = 0.03 * np.exp(-1.3 * dl) * (re / ri) ** 0.5 / (1.0 + re / 60.0) * (1 + 0.3 * aa)
        frac_e = tf_r * 0.8
        g0 = np.sqrt(1 - 2.0 / (re + 1.0)) / np.sqrt(1 - 2.0 / (ri + 1.0))
        gmin = 0.7 * g0 * (1 - 0.1 * np.exp(-dl))
        gmax = 1.4 * g0 * (1 + 0.1 * np.exp(-dl))
        x = (np.arange(RR_NG) + 0.5) / RR_NG
        prof = np.exp(-0.5 * ((x[None, None, :] - 0.45 - 0.1 * np.tanh(np.log(ri / re))[:, :, None]) / 0.22) ** 2) + 0.05
        frac_g = prof / prof.sum(axis=2, keepdims=True)
        f_ret = 0.3 * np.exp(-rm / 8.0) + 0.01
        w.add_table(
            f"FRAC{isp + 1:02d}",
            [
                Column("rlo", "D", rlo),
                Column("rhi", "D", rhi),
                Column("frac_e", "D", frac_e),
                Column("tf_r", "D", tf_r),
                Column("gmin", "D", gmin),
                Column("gmax", "D", gmax),
                Column("frac_g", "D", frac_g.reshape(RR_NR, RR_NR * RR_NG)),
                Column("f_ret", "D", f_ret),
                Column("f_inf", "D", 1.0 - f_ret - 0.05),
                Column("f_bh", "D", np.full(RR_NR, 0.05)),
            ],
        )
    w.close()


FILES = {
    "rel": "rel_table_v0.5a.fits",
    "lp": "rel_lp_table_v0.5b.fits",
    "xill": "xillver-a-Ec5.fits",
    "xillcp": "xillverCp_v3.4.fits",
    "xillns": "xillverNS-2.fits",
    "xillco": "xillverCO.fits",
    "rrad": "table_returnRad_v20220301.fits",
}


def generate(outdir, size="test", which=("rel", "lp", "xill", "xillcp", "rrad", "xillns", "xillco"), force=False):
    """Writes the requested tables into `outdir` (skips files already stamped
    with the same size) and returns the directory."""
    os.makedirs(outdir, exist_ok=True)
    stamp_path = os.path.join(outdir, "synth_stamp.json")
    stamp = {}
    if os.path.exists(stamp_path) and not force:
        with open(stamp_path) as f:
            stamp = json.load(f)
    for key in which:
        path = os.path.join(outdir, FILES[key])
        tag = f"v3:{size}" if key in ("xill", "xillcp") else "v3"
        if stamp.get(key) == tag and os.path.exists(path):
            continue
        tmp = path + ".tmp"
        if key == "rel":
            write_rel_table(tmp)
        elif key == "lp":
            write_lp_table(tmp)
        elif key == "xill":
            write_xillver(tmp, size, nthcomp=False)
        elif key == "xillcp":
            write_xillver(tmp, size, nthcomp=True)
        elif key == "xillns":
            write_xillver_nsco(tmp, "ns")
        elif key == "xillco":
            write_xillver_nsco(tmp, "co")
        elif key == "rrad":
            write_rrad_table(tmp)
        os.replace(tmp, path)
        stamp[key] = tag
        with open(stamp_path, "w") as f:
            json.dump(stamp, f)
    return outdir


def default_table_dir(size="test"):
    """Where tests/bench keep the synthetic tables (outside the repo: they are
    regenerated on every fresh box in a few seconds)."""
    base = os.environ.get("RELXILL_B200_TABLE_CACHE", "/tmp/relxill_b200_tables")
    return os.path.join(base, size)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("outdir")
    ap.add_argument("--size", default="test", choices=sorted(SIZES))
    ap.add_argument("--which", default="rel,lp,xill,xillcp,rrad,xillns,xillco")
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    generate(args.outdir, args.size, tuple(args.which.split(",")), args.force)
    print(args.outdir)
