// tables.cu — read the relxill FITS tables and lay them out in HBM.
//
// File layouts are the reference's (src/reltable.c:171-448, src/xilltable.c:169-276,513-564,
// src/Relreturn_Table.cpp:289-361).  What changes is the residency and layout:
//   * everything is loaded eagerly and stays resident on the device (the reference loads
//     xillver rows lazily per interpolation corner, src/xilltable.c:575-635);
//   * the four transfer-function columns are interleaved as float4 per (a, mu0, r, g*) so one
//     16-byte load fetches trff1/trff2/cosne1/cosne2 of a corner;
//   * xillver rows are padded to a 128-byte multiple for aligned vector loads and are
//     renormalised once at load exactly like renorm_xill_spec (src/xilltable.c:478-485);
//   * linear functionals of the xillver spectra that the returning-radiation correction needs
//     (band energy flux, band photon flux before/after the g=2/3 shift) are reduced to three
//     scalars per table node at load, so the kernels never re-read the spectra for them.
#include "tables.h"
#include "kernels.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "minifits.h"

namespace rx {

#define CUDA_OK(call)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) return std::string("CUDA error: ") + cudaGetErrorString(e_);       \
  } while (0)

Tables::~Tables() {
  for (void *p : allocs_) cudaFree(p);
}

template <class T> const T *Tables::upload(const T *p, size_t n) {
  void *d = nullptr;
  if (cudaMalloc(&d, n * sizeof(T) + 256) != cudaSuccess) { upload_failed_ = true; return nullptr; }
  if (cudaMemcpy(d, p, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {   // every loader checks upload_failed_ before it returns
    cudaFree(d);
    upload_failed_ = true;
    return nullptr;
  }
  allocs_.push_back(d);
  dev_bytes_ += n * sizeof(T);
  return (const T *) d;
}
template <class T> const T *Tables::upload(const std::vector<T> &v) { return upload(v.data(), v.size()); }

static void log_grid(std::vector<double> &e, int n, double emin, double emax) {  // src/relutility.c:399-405
  e.resize(n);
  for (int i = 0; i < n; i++) {
    e[i] = 1.0 * i / (n - 1) * (std::log(emax) - std::log(emin)) + std::log(emin);
    e[i] = std::exp(e[i]);
  }
}

static int lower_index(const double *arr, int n, double val) {
  int klo = 0, khi = n - 1;
  while (khi - klo > 1) {
    const int k = (khi + klo) / 2;
    if (arr[k] > val) khi = k; else klo = k;
  }
  return klo;
}

// Flux-conserving rebin between monotone grids, same bin selection and arithmetic as
// _rebin_spectrum (src/relutility.c:549-601).  Host version (table preprocessing only).
static void rebin_host(const double *ener, double *flu, int nbins, const double *ener0, const double *flu0, int nbins0) {
  int imin = 0, imax = 0;
  for (int ii = 0; ii < nbins; ii++) {
    flu[ii] = 0.0;
    if ((ener0[0] <= ener[ii + 1]) && (ener0[nbins0] >= ener[ii])) {
      while (imin <= nbins0 && ener0[imin] <= ener[ii]) imin++;
      if (imin > 0) imin--;
      while (imax < nbins0 && ener0[imax] <= ener[ii + 1]) imax++;
      if (imax > 0) imax--;
      double elo = ener[ii], ehi = ener[ii + 1];
      if (elo < ener0[imin]) elo = ener0[imin];
      if (ehi > ener0[imax + 1]) ehi = ener0[imax + 1];
      if (imax == imin) {
        flu[ii] = (ehi - elo) / (ener0[imin + 1] - ener0[imin]) * flu0[imin];
      } else {
        const double dmin = (ener0[imin + 1] - elo) / (ener0[imin + 1] - ener0[imin]);
        const double dmax = (ehi - ener0[imax]) / (ener0[imax + 1] - ener0[imax]);
        flu[ii] += flu0[imin] * dmin + flu0[imax] * dmax;
        for (int jj = imin + 1; jj <= imax - 1; jj++) flu[ii] += flu0[jj];
      }
    }
  }
}

void Tables::load_fixed() {
  if (have_fixed_) return;
  log_grid(econv_, NCONV + 1, CONV_EMIN, CONV_EMAX);   // src/Xillspec.h:36-38
  log_grid(ecoarse_, NCOARSE + 1, 0.1, 1000.0);   // src/Xillspec.h:28-32
  std::vector<double> cf(NCONV);
  std::vector<unsigned char> band(NCONV), m1(NCOARSE), m2(NCOARSE);
  for (int i = 0; i < NCONV; i++) {
    cf[i] = 0.5 * (econv_[i] + econv_[i + 1]) / (econv_[i + 1] - econv_[i]);        // src/Relbase.cpp:93-103
    band[i] = (econv_[i] >= 0.01 && econv_[i + 1] < 1000.0) ? 1 : 0;               // src/Relbase.cpp:205
  }
  for (int i = 0; i < NCOARSE; i++) {
    m1[i] = (ecoarse_[i] >= 0.1 && ecoarse_[i] <= 1000.0) ? 1 : 0;                 // src/Xillspec.cpp:199
    m2[i] = (ecoarse_[i] >= 0.1 && ecoarse_[i + 1] <= 1000) ? 1 : 0;               // src/Xillspec.cpp:139
  }
  std::vector<double> gstar(NG), dg(NG);
  const double H = 5e-3;
  for (int i = 0; i < NG; i++) gstar[i] = H + (1.0 - 2 * H) / (NG - 1) * ((float) (i));  // src/Relprofile.cpp:104-106
  for (int i = 0; i < NG; i++)
    dg[i] = (i == 0 || i == NG - 1) ? 0.5 * (gstar[1] - gstar[0]) + H : gstar[1] - gstar[0];
  std::vector<double> tw(2 * NCONV);
  for (int k = 0; k < NCONV; k++) {
    const double ang = -2.0 * M_PI * k / NCONV;
    tw[2 * k] = std::cos(ang);
    tw[2 * k + 1] = std::sin(ang);
  }
  // W = DFT(band): sum over the band of a convolution = sum_k P[k] conj(W[k])  (host radix-2 FFT).  The E_mid/dE
  // weights of the reference (src/Relbase.cpp:93-103) are one constant on this grid — conv_cf_dev holds the
  // largest relative deviation (~1e-13) — so the kernels convolve the plain spectra
  std::vector<double> wr(NCONV), wi(NCONV, 0.0);
  for (int i = 0; i < NCONV; i++) {
    wr[i] = band[i] ? 1.0 : 0.0;
    conv_cf_dev_ = std::max(conv_cf_dev_, std::fabs(cf[i] / cf[NCONV / 2] - 1.0));
  }
  for (int i = 1, j = 0; i < NCONV; i++) {
    int bit = NCONV >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(wr[i], wr[j]); std::swap(wi[i], wi[j]); }
  }
  for (int len = 2; len <= NCONV; len <<= 1) {
    const int half = len >> 1, step = NCONV / len;
    for (int i = 0; i < NCONV; i += len)
      for (int k = 0; k < half; k++) {
        const double c = tw[2 * k * step], s = tw[2 * k * step + 1];
        const double vr = wr[i + k + half] * c - wi[i + k + half] * s, vi = wr[i + k + half] * s + wi[i + k + half] * c;
        const double ur = wr[i + k], ui = wi[i + k];
        wr[i + k] = ur + vr; wi[i + k] = ui + vi;
        wr[i + k + half] = ur - vr; wi[i + k + half] = ui - vi;
      }
  }
  std::vector<double> wpack(2 * (NCONV / 2 + 1));
  for (int k = 0; k <= NCONV / 2; k++) { wpack[2 * k] = wr[k]; wpack[2 * k + 1] = wi[k]; }
  int b0 = NCONV, b1 = -1;
  for (int i = 0; i < NCONV; i++)
    if (band[i]) { b0 = std::min(b0, i); b1 = std::max(b1, i); }
  dt_.econv = upload(econv_);
  dt_.conv_cf = upload(cf);
  dt_.conv_b0 = b0;
  dt_.conv_b1 = b1;
  dt_.conv_i1kev = lower_index(econv_.data(), NCONV + 1, 1.0);                      // src/Relbase.cpp:133-137
  dt_.conv_i3kev = lower_index(econv_.data(), NCONV, 3.0);   // renorm_relxill_spectrum_1keV: binary_search(energy, num_flux_bins, 3.0), src/Relxill.cpp:252-253
  dt_.ecoarse = upload(ecoarse_);
  dt_.coarse_m1 = upload(m1);
  dt_.coarse_m2 = upload(m2);
  dt_.gstar = upload(gstar);
  dt_.d_gstar = upload(dg);
  {
    std::vector<double> gw(NG);
    for (int i = 0; i < NG; i++) gw[i] = dg[i] / std::sqrt(gstar[i] - gstar[i] * gstar[i]);
    dt_.gstar_w = upload(gw);
  }
  dt_.tw = upload(tw);
  dt_.conv_w = upload(wpack);
  have_fixed_ = true;
}

std::string Tables::load_rel() {
  if (have_rel_) return "";
  const std::string path = dir_ + "/rel_table_v0.5a.fits";
  mf_file *f = mf_open(path.c_str());
  if (!f) return "cannot open " + path;
  std::vector<float> a(REL_NA), mu(REL_NMU);
  int h = mf_find_hdu(f, "a");
  int hm = mf_find_hdu(f, "mu0");
  if (!h || !hm || f->nhdu < 3 + REL_NA * REL_NMU) { mf_close(f); return "rel table: unexpected layout in " + path; }
  if (mf_read(&f->hdus[h - 1], mf_find_col(&f->hdus[h - 1], "a"), 1, 1, REL_NA, 'f', a.data()) ||
      mf_read(&f->hdus[hm - 1], mf_find_col(&f->hdus[hm - 1], "mu0"), 1, 1, REL_NMU, 'f', mu.data())) {
    mf_close(f);
    return "rel table: cannot read the spin / inclination axes of " + path;
  }
  const size_t n1 = (size_t) REL_NA * REL_NMU * REL_NRT;
  std::vector<float> r(n1), gmin(n1), gmax(n1), tc(n1 * NG * 4), col((size_t) REL_NRT * NG);
  const char *names[4] = {"trff1", "trff2", "cosne1", "cosne2"};
  for (int ia = 0; ia < REL_NA; ia++)
    for (int im = 0; im < REL_NMU; im++) {
      const mf_hdu *hd = &f->hdus[ia * REL_NMU + im + 4 - 1];  // by HDU number, src/reltable.c:290
      const size_t o1 = ((size_t) ia * REL_NMU + im) * REL_NRT;
      int rc = 0;
      rc |= mf_read(hd, mf_find_col(hd, "r"), 1, 1, REL_NRT, 'f', r.data() + o1);
      rc |= mf_read(hd, mf_find_col(hd, "gmin"), 1, 1, REL_NRT, 'f', gmin.data() + o1);
      rc |= mf_read(hd, mf_find_col(hd, "gmax"), 1, 1, REL_NRT, 'f', gmax.data() + o1);
      for (int q = 0; q < 4; q++) {
        rc |= mf_read(hd, mf_find_col(hd, names[q]), 1, 1, REL_NRT * NG, 'f', col.data());
        for (size_t k = 0; k < (size_t) REL_NRT * NG; k++) tc[(o1 * NG + k) * 4 + q] = col[k];
      }
      if (rc) { mf_close(f); return "rel table: read error in " + path; }
    }
  mf_close(f);
  dt_.rel_a = upload(a);
  dt_.rel_mu0 = upload(mu);
  dt_.rel_r = upload(r);
  dt_.rel_gmin = upload(gmin);
  dt_.rel_gmax = upload(gmax);
  dt_.rel_tc = upload(tc);
  if (upload_failed_ || !dt_.rel_tc) return "out of device memory (rel table)";
  have_rel_ = true;
  return "";
}

std::string Tables::load_lp() {
  if (have_lp_) return "";
  const std::string path = dir_ + "/rel_lp_table_v0.5b.fits";
  mf_file *f = mf_open(path.c_str());
  if (!f) return "cannot open " + path;
  const int h = mf_find_hdu(f, "I_h");
  if (!h) { mf_close(f); return "lp table: no I_h extension in " + path; }
  const mf_hdu *hd = &f->hdus[h - 1];
  std::vector<float> a(LP_NA), hh((size_t) LP_NA * LP_NH), rad((size_t) LP_NA * LP_NRT);
  const size_t n3 = (size_t) LP_NA * LP_NH * LP_NRT;
  std::vector<float> in(n3), de(n3), di(n3);
  int rc = mf_read(hd, mf_find_col(hd, "a"), 1, 1, LP_NA, 'f', a.data());
  const int c_hg = mf_find_col(hd, "hgrid"), c_r = mf_find_col(hd, "r");
  for (int ia = 0; ia < LP_NA && !rc; ia++) {
    rc |= mf_read(hd, c_hg, ia + 1, 1, LP_NH, 'f', hh.data() + (size_t) ia * LP_NH);
    rc |= mf_read(hd, c_r, ia + 1, 1, LP_NRT, 'f', rad.data() + (size_t) ia * LP_NRT);
  }
  for (int ih = 0; ih < LP_NH && !rc; ih++) {
    char nm[32];
    snprintf(nm, sizeof(nm), "h%i", ih + 1);
    const int c1 = mf_find_col(hd, nm);
    snprintf(nm, sizeof(nm), "del%i", ih + 1);
    const int c2 = mf_find_col(hd, nm);
    snprintf(nm, sizeof(nm), "del_inc%i", ih + 1);
    const int c3 = mf_find_col(hd, nm);
    for (int ia = 0; ia < LP_NA; ia++) {
      const size_t o = ((size_t) ia * LP_NH + ih) * LP_NRT;
      rc |= mf_read(hd, c1, ia + 1, 1, LP_NRT, 'f', in.data() + o);
      rc |= mf_read(hd, c2, ia + 1, 1, LP_NRT, 'f', de.data() + o);
      rc |= mf_read(hd, c3, ia + 1, 1, LP_NRT, 'f', di.data() + o);
    }
  }
  mf_close(f);
  if (rc) return "lp table: read error in " + path;
  for (size_t k = 0; k < n3; k++) {  // the sign of the angles is dropped at load, src/reltable.c:370-377
    de[k] = fabsf(de[k]);
    di[k] = fabsf(di[k]);
  }
  dt_.lp_a = upload(a);
  dt_.lp_h = upload(hh);
  dt_.lp_rad = upload(rad);
  dt_.lp_int = upload(in);
  dt_.lp_del = upload(de);
  dt_.lp_dinc = upload(di);
  if (upload_failed_ || !dt_.lp_dinc) return "out of device memory (lp table)";
  have_lp_ = true;
  return "";
}

std::string Tables::load_rrad() {
  if (have_rr_) return "";
  const std::string path = dir_ + "/table_returnRad_v20220301.fits";
  mf_file *f = mf_open(path.c_str());
  if (!f) return "cannot open " + path;
  const int hs = mf_find_hdu(f, "SPIN");
  if (!hs) { mf_close(f); return "returnRad table: no SPIN extension"; }
  const int ns = (int) f->hdus[hs - 1].nrows;
  rr_spin_.resize(ns);
  if (ns < 1 || mf_read(&f->hdus[hs - 1], mf_find_col(&f->hdus[hs - 1], "a"), 1, 1, ns, 'd', rr_spin_.data())) {
    mf_close(f);
    rr_spin_.clear();
    return "returnRad table: cannot read the spin axis";
  }
  const size_t n2 = (size_t) RR_NR * RR_NR;
  std::vector<double> rlo((size_t) ns * RR_NR), rhi((size_t) ns * RR_NR), tf(ns * n2), gmin(ns * n2), gmax(ns * n2),
      fg(ns * n2 * RR_NG), lng(ns * n2 * RR_NG);
  int rc = 0;
  for (int s = 0; s < ns; s++) {
    char nm[32];
    snprintf(nm, sizeof(nm), "FRAC%02i", s + 1);
    const int h = mf_find_hdu(f, nm);
    if (!h) { mf_close(f); return std::string("returnRad table: missing extension ") + nm; }
    const mf_hdu *hd = &f->hdus[h - 1];
    rc |= mf_read(hd, mf_find_col(hd, "rlo"), 1, 1, RR_NR, 'd', rlo.data() + (size_t) s * RR_NR);
    rc |= mf_read(hd, mf_find_col(hd, "rhi"), 1, 1, RR_NR, 'd', rhi.data() + (size_t) s * RR_NR);
    rc |= mf_read(hd, mf_find_col(hd, "tf_r"), 1, 1, n2, 'd', tf.data() + s * n2);
    rc |= mf_read(hd, mf_find_col(hd, "gmin"), 1, 1, n2, 'd', gmin.data() + s * n2);
    rc |= mf_read(hd, mf_find_col(hd, "gmax"), 1, 1, n2, 'd', gmax.data() + s * n2);
    rc |= mf_read(hd, mf_find_col(hd, "frac_g"), 1, 1, n2 * RR_NG, 'd', fg.data() + s * n2 * RR_NG);
  }
  mf_close(f);
  if (rc) return "returnRad table: read error";
  for (size_t q = 0; q < ns * n2; q++)  // g grid of src/Relreturn_Datastruct.cpp:51-59, stored as ln g
    for (int j = 0; j < RR_NG; j++) {
      const double g = ((j + 0.5) / RR_NG) * (gmax[q] - gmin[q]) + gmin[q];
      lng[q * RR_NG + j] = std::log(g);
    }
  dt_.rr_nspin = ns;
  dt_.rr_spin = upload(rr_spin_);
  dt_.rr_rlo = upload(rlo);
  dt_.rr_rhi = upload(rhi);
  dt_.rr_tf = upload(tf);
  dt_.rr_gmin = upload(gmin);
  dt_.rr_gmax = upload(gmax);
  {  // {frac_g, ln g} pairs, g-bin major: the (ring, ring) pairs of neighbouring threads are neighbours in memory
    std::vector<double> fgl((size_t) ns * n2 * RR_NG * 2);
    for (int s = 0; s < ns; s++)
      for (size_t q = 0; q < n2; q++)
        for (int j = 0; j < RR_NG; j++) {
          const size_t o = (((size_t) s * RR_NG + j) * n2 + q) * 2;
          fgl[o] = fg[(s * n2 + q) * RR_NG + j];
          fgl[o + 1] = lng[(s * n2 + q) * RR_NG + j];
        }
    dt_.rr_fgl = upload(fgl);
  }
  if (upload_failed_ || !dt_.rr_fgl) return "out of device memory (returnRad table)";
  have_rr_ = true;
  return "";
}

static int xill_param_id(const char *name) {  // src/common.h:141-161
  if (!strcmp(name, "Gamma")) return 0;
  if (!strcmp(name, "A_Fe")) return 1;
  if (!strcmp(name, "logXi")) return 2;
  if (!strcmp(name, "Ecut") || !strcmp(name, "kTe")) return 3;
  if (!strcmp(name, "Dens")) return 4;
  if (!strcmp(name, "kTbb")) return 5;
  if (!strcmp(name, "A_CO")) return 1;   // PARAM_ACO == PARAM_AFE
  if (!strcmp(name, "Frac")) return 6;
  if (!strcmp(name, "Incl")) return 7;
  return -1;
}

std::string Tables::load_xill(int which) {
  XillHost &xh = xh_[which];
  if (xh.loaded) return "";
  load_fixed();
  static const char *const names[XT_COUNT] = {"/xillver-a-Ec5.fits", "/xillverCp_v3.4.fits", "/xillverNS-2.fits", "/xillverCO.fits"};   // src/common.h:165-168
  const std::string path = dir_ + names[which];
  mf_file *f = mf_open(path.c_str());
  if (!f) return "cannot open " + path;
  const int hp = mf_find_hdu(f, "PARAMETERS"), he = mf_find_hdu(f, "ENERGIES"), hs = mf_find_hdu(f, "SPECTRA");
  if (!hp || !he || !hs) { mf_close(f); return "xillver table: missing extension in " + path; }
  const mf_hdu *P = &f->hdus[hp - 1], *E = &f->hdus[he - 1], *S = &f->hdus[hs - 1];
  xh.npar = (int) P->nrows;
  if (xh.npar != 5 && xh.npar != 6) { mf_close(f); return "xillver table: wrong dimensionality"; }
  // XSPEC atable layout: column 1 NAME (string), 9 NUMBVALS, 10 VALUE (src/xilltable.c:169-238 reads them by number)
  if (P->ncols < 10 || mf_read(P, 9, 1, 1, xh.npar, 'i', xh.nvals)) { mf_close(f); return "xillver table: cannot read NUMBVALS in " + path; }
  long nrows = 1;
  int ax_lxi = -1, ax_dns = -1;
  for (int i = 0; i < xh.npar; i++) {
    char nm[16];
    if (mf_read_str(P, 1, i + 1, nm, sizeof(nm))) { mf_close(f); return "xillver table: cannot read the parameter names in " + path; }
    xh.pindex[i] = xill_param_id(nm);
    if (xh.pindex[i] < 0) { mf_close(f); return std::string("xillver table: unknown parameter ") + nm; }
    if (xh.nvals[i] < 2) { mf_close(f); return std::string("xillver table: fewer than two values on axis ") + nm; }
    xh.vals[i].resize(xh.nvals[i]);
    if (mf_read(P, 10, i + 1, 1, xh.nvals[i], 'f', xh.vals[i].data())) { mf_close(f); return std::string("xillver table: cannot read the values of axis ") + nm; }
    nrows *= xh.nvals[i];
    if (xh.pindex[i] == 2) ax_lxi = i;
    if (xh.pindex[i] == 4) ax_dns = i;
  }
  if (xh.pindex[xh.npar - 1] != 7 || S->nrows != nrows) { mf_close(f); return "xillver table: Incl must be the last axis"; }
  // the kernels assume the reference's standard axis order: [Gamma,] A_Fe, logXi, Ecut|kTe, [Dens,] Incl
  xh.n_incl = xh.nvals[xh.npar - 1];
  if (xh.n_incl > 10) { mf_close(f); return "xillver table: more than 10 inclinations are not supported"; }
  xh.n_ener = (int) E->nrows;
  xh.stride = ((xh.n_ener + 31) / 32) * 32;
  xh.nnodes = nrows / xh.n_incl;
  std::vector<float> elo(xh.n_ener), ehi(xh.n_ener);
  if (xh.n_ener < 2 || mf_read(E, 1, 1, 1, xh.n_ener, 'f', elo.data()) || mf_read(E, 2, 1, 1, xh.n_ener, 'f', ehi.data())) {
    mf_close(f);
    return "xillver table: cannot read the energy grid of " + path;
  }
  xh.ener.resize(xh.n_ener + 1);
  for (int i = 0; i < xh.n_ener; i++) xh.ener[i] = elo[i];                         // src/xilltable.c:1105-1109
  xh.ener[xh.n_ener] = ehi[xh.n_ener - 1];

  const mf_col *sc = &S->cols[1];  // column 2 = INTPSPEC, src/xilltable.c:544
  if (sc->code != 'E' || sc->width != xh.n_ener) { mf_close(f); return "xillver table: INTPSPEC column has the wrong shape"; }
  const int ne = xh.n_ener, ni = xh.n_incl, st = xh.stride;
  std::vector<double> ef(xh.nnodes), p1(xh.nnodes), p2(xh.nnodes);
  std::vector<double> w(ni), avg(ne), ez(ne + 1), fz(ne);
  for (int m = 0; m < ni; m++) {  // angle-average weights, src/Xillspec.cpp:109-126 (degrees passed through *180/pi)
    const double incl_deg = ((double) xh.vals[xh.npar - 1][m]) * 180 / M_PI;
    w[m] = 0.5 * std::cos(incl_deg * M_PI / 180) / ni;
  }
  const double gref = 2. / 3.;
  for (int i = 0; i <= ne; i++) ez[i] = xh.ener[i] / gref;

  // fixed rebin map xillver grid -> convolution grid
  std::vector<int> imin_v(NCONV), imax_v(NCONV);
  std::vector<double> dmin_v(NCONV), dmax_v(NCONV);
  {
    const double *e0 = xh.ener.data(), *e = econv_.data();
    int imin = 0, imax = 0;
    for (int ii = 0; ii < NCONV; ii++) {
      imin_v[ii] = -1; imax_v[ii] = -1; dmin_v[ii] = 0; dmax_v[ii] = 0;
      if ((e0[0] <= e[ii + 1]) && (e0[ne] >= e[ii])) {
        while (imin <= ne && e0[imin] <= e[ii]) imin++;
        if (imin > 0) imin--;
        while (imax < ne && e0[imax] <= e[ii + 1]) imax++;
        if (imax > 0) imax--;
        double elo_ = e[ii], ehi_ = e[ii + 1];
        if (elo_ < e0[imin]) elo_ = e0[imin];
        if (ehi_ > e0[imax + 1]) ehi_ = e0[imax + 1];
        imin_v[ii] = imin; imax_v[ii] = imax;
        if (imax == imin) {
          dmin_v[ii] = (ehi_ - elo_) / (e0[imin + 1] - e0[imin]);
        } else {
          dmin_v[ii] = (e0[imin + 1] - elo_) / (e0[imin + 1] - e0[imin]);
          dmax_v[ii] = (ehi_ - e0[imax]) / (e0[imax + 1] - e0[imax]);
        }
      }
    }
  }
  // the convolution bins that overlap the table grid form one interval [first, last]
  int first = -1, last = -2;
  for (int i = 0; i < NCONV; i++)
    if (imin_v[i] >= 0) { if (first < 0) first = i; last = i; }
  if (first < 0) { mf_close(f); return "xillver table: energy grid does not overlap the convolution grid"; }
  for (int i = first; i <= last; i++)
    if (imin_v[i] < 0) { mf_close(f); return "xillver table: energy grid is not monotone"; }
  xh.xc_first = first;
  xh.xc_n = last - first + 1;
  xh.xc_stride = ((xh.xc_n + 31) / 32) * 32;
  // Convolution-grid copy (relxill models only need it; standalone xillver models interpolate the table grid): every
  // row rebinned exactly like _rebin_spectrum would rebin it (src/relutility.c:549-601), kept in fp64
  const bool want_c = conv_grid_copy_ && (which == XT_STD || which == XT_CP || which == XT_NS || which == XT_CO);
  double *d_datac = nullptr;
  const size_t total_c = (size_t) S->nrows * xh.xc_stride;
  if (want_c) {
    if (cudaMalloc((void **) &d_datac, total_c * sizeof(double)) != cudaSuccess) { mf_close(f); return "out of device memory (xillver table, convolution-grid copy)"; }
    allocs_.push_back(d_datac);
    dev_bytes_ += total_c * sizeof(double);
  }
  float *d_data = nullptr;
  const size_t total = (size_t) nrows * st;
  if (cudaMalloc((void **) &d_data, total * sizeof(float)) != cudaSuccess) { mf_close(f); return "out of device memory (xillver table)"; }
  allocs_.push_back(d_data);
  dev_bytes_ += total * sizeof(float);
  xh.bytes = total * sizeof(float);
  // stream node blocks through a staging buffer
  const long nodes_per_blk = 256;
  std::vector<float> stage((size_t) nodes_per_blk * ni * st, 0.f);
  std::vector<double> stage_c(want_c ? (size_t) nodes_per_blk * ni * xh.xc_stride : 0, 0.0);
  for (long n0 = 0; n0 < xh.nnodes; n0 += nodes_per_blk) {
    const long nb = std::min(nodes_per_blk, xh.nnodes - n0);
    for (long nd = 0; nd < nb; nd++) {
      const long node = n0 + nd;
      long rem = node;
      int idx[6] = {0};
      for (int i = xh.npar - 2; i >= 0; i--) { idx[i] = (int) (rem % xh.nvals[i]); rem /= xh.nvals[i]; }
      const double lxi = (ax_lxi >= 0) ? (double) xh.vals[ax_lxi][idx[ax_lxi]] : 0.0;
      // an axis the table lacks takes the model's fixed value (getDefaultLogxi / getDefaultDensity, src/xilltable.c:566-573):
      // logxi 0; logN 17 for the CO table, else 15 (src/ModelDefinition.cpp:361-363)
      const double dens = (ax_dns >= 0) ? (double) xh.vals[ax_dns][idx[ax_dns]] : (which == XT_CO ? 17.0 : 15.0);
      const double pl = std::pow(10, lxi), pd = std::pow(10, dens - 15);
      const bool do_d = std::fabs(dens - 15) > 1e-6;
      for (int e = 0; e < ne; e++) avg[e] = 0.0;
      for (int m = 0; m < ni; m++) {
        float *dst = stage.data() + ((size_t) nd * ni + m) * st;
        const long row = node * ni + m;
        mf_copy_f32(S->data + (size_t) row * S->rowbytes + sc->offset, dst, ne);
        for (int e = 0; e < ne; e++) {  // renorm_xill_spec: float /= double, twice
          dst[e] /= pl;
          if (do_d) dst[e] /= pd;
          avg[e] += w[m] * (double) dst[e];
        }
      }
      double s_ef = 0.0, s_p1 = 0.0, s_p2 = 0.0;
      rebin_host(ez.data(), fz.data(), ne, xh.ener.data(), avg.data(), ne);       // src/Xillspec.cpp:457-487
      for (int e = 0; e < ne; e++) {
        if (xh.ener[e] >= 0.1 && xh.ener[e + 1] <= 1000) s_ef += avg[e] * 0.5 * (xh.ener[e] + xh.ener[e + 1]);
        if (xh.ener[e] >= 0.15 && xh.ener[e + 1] <= 500.0) { s_p1 += avg[e]; s_p2 += fz[e] * gref; }
      }
      ef[node] = s_ef; p1[node] = s_p1; p2[node] = s_p2;
    }
    if (cudaMemcpy(d_data + (size_t) n0 * ni * st, stage.data(), (size_t) nb * ni * st * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
      mf_close(f);
      return "xillver table: upload failed (" + path + ")";
    }
    if (want_c) {
      for (long r = 0; r < nb * ni; r++) {
        const float *src = stage.data() + (size_t) r * st;
        double *dst = stage_c.data() + (size_t) r * xh.xc_stride;
        for (int c = 0; c < xh.xc_n; c++) {
          const int i0 = imin_v[first + c], i1 = imax_v[first + c];
          double v = (double) src[i0] * dmin_v[first + c] + (double) src[i1] * dmax_v[first + c];   // k_conv's order of operations
          if (i1 - i0 >= 2) {
            v += (double) src[i0 + 1];
            for (int jj = i0 + 2; jj <= i1 - 1; jj++) v += (double) src[jj];
          }
          dst[c] = v;
        }
      }
      if (cudaMemcpy(d_datac + (size_t) n0 * ni * xh.xc_stride, stage_c.data(), (size_t) nb * ni * xh.xc_stride * sizeof(double),
                     cudaMemcpyHostToDevice) != cudaSuccess) {
        mf_close(f);
        return "xillver table: upload of the convolution-grid copy failed (" + path + ")";
      }
    }
  }
  mf_close(f);

  XillDev &xd = dt_.xill[which];
  xd.npar = xh.npar;
  for (int i = 0; i < 6; i++) { xd.nvals[i] = xh.nvals[i]; xd.pindex[i] = xh.pindex[i]; xd.vals[i] = nullptr; }
  for (int i = 0; i < xh.npar; i++) xd.vals[i] = upload(xh.vals[i]);
  xd.n_ener = ne; xd.n_incl = ni; xd.stride = st; xd.nnodes = xh.nnodes;
  xd.data = d_data;
  xd.ener = upload(xh.ener);
  xd.incl = xd.vals[xh.npar - 1];
  xd.node_ef = upload(ef);
  xd.node_p1 = upload(p1);
  xd.node_p2 = upload(p2);
  std::vector<int> ii_v(2 * NCONV);
  std::vector<double> dd_v(2 * NCONV);
  for (int i = 0; i < NCONV; i++) {
    ii_v[2 * i] = std::max(imin_v[i], 0); ii_v[2 * i + 1] = std::max(imax_v[i], 0);   // outside the table grid: bin 0, zero weights
    dd_v[2 * i] = dmin_v[i]; dd_v[2 * i + 1] = dmax_v[i];
  }
  xd.rb_ii = upload(ii_v);
  xd.rb_dd = upload(dd_v);
  if (upload_failed_ || !xd.rb_dd) return "out of device memory (xillver table)";
  xd.xc_first = xh.xc_first; xd.xc_n = xh.xc_n; xd.xc_stride = xh.xc_stride;
  xd.datac = d_datac;
  xh.has_conv_copy = d_datac != nullptr;
  xh.bytes_c = want_c ? total_c * sizeof(double) : 0;
  xh.loaded = true;
  return "";
}

// Multicolour-disk photon spectrum shape used as the nthcomp seed (XSPEC diskbb interpolation formula as in
// reference src/donthcomp.c:49-108, f_mcdint__).  Host only: evaluated once for the fixed kT_bb = 0.05 keV.
static double mcd_value(double et) {
  static const double gc[3] = {.078196667, -1.066202, 1.192418};
  static const double gw[3] = {.5207874, .513457, .4077983};
  static const double gn[3] = {.3728691, .039775528, .037766505};
  static const double res[98] = {
      9.6198382e-4, .0010901181, .0012310012, .0013841352, .0015481583, .0017210036, .0018988943, .002076939,
      .0022484281, .0024049483, .0025366202, .0026316255, .0026774985, .0026613059, .0025708784, .0023962965,
      .002130655, .0017725174, .0013268656, 8.0657672e-4, 2.3337584e-4, -3.6291778e-4, -9.4443569e-4,
      -.0014678875, -.0018873741, -.0021588493, -.0022448371, -.0021198179, -.0017754602, -.0012246034,
      -5.0414167e-4, 3.2507078e-4, .0011811065, .0019673402, .0025827094, .0029342526, .0029517083, .0026012166,
      .0018959062, 9.0128649e-4, -2.6757144e-4, -.0014567885, -.002492855, -.0032079776, -.0034678637,
      -.0031988217, -.0024080969, -.001193624, 2.6134145e-4, .0017117758, .0028906898, .0035614435, .0035711778,
      .0028921374, .0016385898, 4.9857464e-5, -.0015572671, -.0028578151, -.0035924212, -.0036253044,
      -.002975086, -.0018044436, -3.7796664e-4, .0010076215, .0020937327, .0027090854, .0028031667, .0024276576,
      .0017175597, 8.1030795e-4, -1.2592304e-4, -9.4888491e-4, -.0015544816, -.0018831972, -.0019203142,
      -.0016905849, -.0012487737, -6.6789911e-4, -2.7079461e-5, 5.9931935e-4, .0011499748, .0015816521,
      .0018709224, .0020129966, .0020184702, .0019089181, .0017122289, .001458377, .0011760717, 8.9046768e-4,
      6.2190822e-4, 3.8553762e-4, 1.9155022e-4, 4.5837109e-5, -4.9177834e-5, -9.3670762e-5, -8.9622968e-5,
      -4.01538532e-5};
  const double log10e = 0.43429448190325182765;
  const double loget = log10e * std::log(et);
  double pos = (loget - log10e * std::log(.001)) / .06 + 1;
  int j = (int) pos;
  double resfact;
  if (j < 1) resfact = res[0];
  else if (j >= 98) resfact = res[97];
  else { pos -= j; resfact = res[j - 1] * (1. - pos) + res[j] * pos; }
  double gaufact = 1.;
  for (j = 1; j <= 3; ++j) {
    const double z = (loget - gc[j - 1]) / gw[j - 1];
    gaufact += gn[j - 1] * std::exp(-z * z / 2.);
  }
  return std::pow(et / .001, -.66666666666666663) * 193.21556 * (std::pow(et, 1.663753) * .52876731 + 1.) * std::exp(-et) * gaufact
         * (resfact + 1.);
}

// Everything in the Kompaneets set-up that depends only on the photon grid, i.e. on the seed temperature,
// which relxill fixes at kT_bb = 0.05 keV (reference src/relutility.c:625-632): the grid x, the Cooper
// coefficient c2, the Klein-Nishina ratio, x^3 and the disk-blackbody photon production rate
// (src/donthcomp.c:467-585).  Computed once on the host, kept in HBM.
void Tables::load_nthcomp() {
  if (have_nth_) return;
  const double log10e = 0.43429448190325182765;
  const double tempbb = 0.05 / 511.;
  const double delta = .02;
  const double xmin = tempbb * 1e-4;
  const int N = NTH_MAX;
  std::vector<double> x(N + 1), w(N), c2(N), rel(N), x3(N + 1), dph(N, 0.0);
  for (int j = 0; j <= N; j++) x[j] = xmin * std::pow(10., j * delta);
  for (int j = 0; j < N; j++) {
    const double ww = x[j];
    const double w1 = std::sqrt(x[j] * x[j + 1]);
    w[j] = w1;
    c2[j] = std::pow(w1, 4.) / (w1 * 4.6 + 1. + w1 * 1.1 * w1);
    if (ww <= .05) {
      rel[j] = 1 - ww * 2 + ww * 26 * ww / 5;
    } else {
      const double z1 = (ww + 1) / (ww * (ww * ww));
      const double z2 = ww * 2 + 1;
      const double z3 = std::log(z2);
      const double z4 = ww * 2 * (ww + 1) / z2;
      const double z5 = z3 / 2 / ww;
      const double z6 = (ww * 3 + 1) / z2 / z2;
      rel[j] = (z1 * (z4 - z3) + z5 - z6) * .75;
    }
  }
  for (int j = 0; j <= N; j++) x3[j] = std::pow(x[j], 3.);
  int jmaxth = (int) (log10e * std::log(tempbb * 50. / xmin) / delta);
  if (jmaxth > 900) jmaxth = 900;
  {  // disk-blackbody seed: 5-point Gauss integration per photon-grid cell (f_xsdskb__, :142-197)
    static const double gw5[5] = {.236926885, .47862867, .568888888, .47862867, .236926885};
    static const double gx5[5] = {-.906179846, -.53846931, 0., .53846931, .906179846};
    std::vector<double> ear(jmaxth), photar(jmaxth, 0.0);
    for (int j = 1; j <= jmaxth - 1; j++) ear[j - 1] = std::sqrt(x[j - 1] * x[j]) * 511.;
    const double tin = tempbb * 511.;
    const int ne = jmaxth - 2;
    for (int i = 1; i <= ne; i++) {
      const double xn = (ear[i] - ear[i - 1]) / 2.f;
      const double xh = xn + ear[i - 1];
      double ph = 0.f;
      for (int j = 0; j < 5; j++) {
        const double e = xn * gx5[j] + xh;
        const double flux = mcd_value(e / tin) * tin * tin * 1. / 361.;
        ph += gw5[j] * flux;
      }
      photar[i - 1] = ph * xn;
    }
    for (int j = 1; j <= ne; j++) dph[j] = photar[j - 1] * 511. / (ear[j] - ear[j - 1]);
    dph[0] = dph[1];
  }
  dt_.nth_x = upload(x);
  dt_.nth_w = upload(w);
  dt_.nth_c2 = upload(c2);
  dt_.nth_rel = upload(rel);
  dt_.nth_x3 = upload(x3);
  dt_.nth_dphdot = upload(dph);
  {  // per-node products and reciprocals of k_nth's elimination rows
    std::vector<double> rw(N), xd(N), x4(N);
    for (int j = 0; j < N; j++) {
      rw[j] = 1.0 / w[j];
      xd[j] = x[j] * dph[j];
      x4[j] = (x[j] * x[j]) * (x[j] * x[j]);
    }
    dt_.nth_rw = upload(rw);
    dt_.nth_xd = upload(xd);
    dt_.nth_x4 = upload(x4);
  }
  dt_.nth_jnr = (int) (log10e * std::log(.1 / xmin) / delta + 1);
  dt_.nth_jrel = (int) (log10e * std::log(1. / xmin) / delta + 1);
  dt_.nth_jmaxth = jmaxth;
  dt_.nth_xmin = xmin;
  dt_.nth_deltal = delta * std::log(10.);
  // The two band integrals of a solution on the coarse grid (the xillver normalisation, src/Xillspec.cpp:179-205, and the
  // returning-radiation flux, :344-362) are linear in the solution E F_E(x_j): photons per coarse bin are the trapezoid of
  // the interpolated solution at the bin edges (c_donthcomp, src/donthcomp.c:759-786), so both are dot products with
  // weights that depend on the two grids only.  k_nth takes them while it back-substitutes and never files the zone
  // solutions.  (Beyond the last node of a solution the reference returns 0; the solution array is 0 there.)
  {
    std::vector<double> w1(N, 0.0), w2(N, 0.0);
    auto edge = [&](double e, double coef, std::vector<double> &wv) {   // coef * p(e) as weights on the nodes
      const double target = e * 1.0;
      int lo = 0, hi = N;
      while (lo < hi) { const int m = (lo + hi) >> 1; if (x[m] * 511. < target) lo = m + 1; else hi = m; }
      const int j = lo + 1;   // first 1-based index with NOT (x[j-1] * 511 < target)
      if (j > N) return;
      if (j > 1) {
        const int jl = j - 1;
        const double fr = (e / 511. * 1.0 - x[jl - 1]) / (x[jl] - x[jl - 1]);
        wv[jl - 1] += coef * (1.0 - fr);
        wv[jl] += coef * fr;
      } else {
        wv[0] += coef;
      }
    };
    for (int i = 0; i < NCOARSE; i++) {
      const double e0 = ecoarse_[i], e1 = ecoarse_[i + 1];
      const bool m1 = (e0 >= 0.1 && e0 <= 1000.0), m2 = (e0 >= 0.1 && e1 <= 1000);
      const double base = .5 * (e1 - e0) * 0.5 * (e0 + e1);   // photons per bin -> energy per bin
      if (m1) { edge(e1, base / (e1 * e1) * 1e20 * 1.602177e-09, w1); edge(e0, base / (e0 * e0) * 1e20 * 1.602177e-09, w1); }
      if (m2) { edge(e1, base / (e1 * e1), w2); edge(e0, base / (e0 * e0), w2); }
    }
    dt_.nth_w1 = upload(w1);
    dt_.nth_w2 = upload(w2);
    // f_spp__ at 1 keV, z = 0 (src/donthcomp.c:651-681): the bracket on the photon grid
    const double xn = 1.0 / 511.;
    const double xx = 1 / (1 / xn);
    int ih = 2;
    while (ih < N && xx > x[ih - 1]) ++ih;
    dt_.nth_ih1 = ih;
    dt_.nth_xx1 = xx;
  }
  have_nth_ = true;
}

std::string Tables::load(const std::string &dir) {
  dir_ = dir;
  load_fixed();
  if (upload_failed_) return "out of device memory (fixed grids)";
  if (conv_cf_dev_ > 1e-11) return "convolution grid: E_mid/dE is not constant (k_conv relies on it)";
  return "";   // the tables themselves are loaded on first use by the model flavour that needs them
}

std::string Tables::require_xill_only(int xtab) {
  load_fixed();
  std::string err = load_xill(xtab);
  if (!err.empty()) return err;
  if (xtab == XT_CP) load_nthcomp();
  if (upload_failed_) return "out of device memory (tables)";
  return "";
}

std::string Tables::require(bool lp, bool rrad, int xtab) {
  std::string err = load_rel();
  if (!err.empty()) return err;
  if (lp && !(err = load_lp()).empty()) return err;
  if (rrad && !(err = load_rrad()).empty()) return err;
  if (xtab != XT_NONE && !(err = load_xill(xtab)).empty()) return err;
  if (xtab == XT_CP) load_nthcomp();
  if (upload_failed_) return "out of device memory (tables)";
  return "";
}

}  // namespace rx
