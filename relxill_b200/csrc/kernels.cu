// kernels.cu — hand-written sm_100a kernels of the relxill spectrum-evaluation hot path.
//
// One batch chunk = C parameter vectors.  Kernel sequence (see DESIGN.md §4 for the data flow and the measurements):
//   k_syspar   1 CTA / vector   (this file)  (a,mu0) table interpolation of the 100 table radii, fine radial grid,
//                               emissivity (broken power law | lamp post [+ returning radiation, second pass])
//   k_zone     1 CTA / vector   (this file)  per-zone ionisation / density / Ecut, xillver corner slots + weights,
//                               primary-spectrum normalisations, returning-radiation correction factors
//   k_nth      64 solves / CTA  (nthcomp.cu) Kompaneets solutions of the zones and the source (Cp models)
//   k_rows     (13, vector)     (this file)  (a,mu0) half of the transfer-function interpolation, per TABLE radius
//   k_fine     (125, vector)    (this file)  per-radius parts of the emission-angle distribution (radial half of the
//                               interpolation in registers; files the fine arrays only for probes / limb darkening)
//   k_dist     1 CTA / vector   (this file)  emission-angle distribution per zone
//   k_line     one WARP per (vector, zone[, run of radii]) (line.cu)  relline profile: table rows staged with one bulk
//                               copy, bin-stationary tiles, Romberg levels 3+ through a shared-memory queue
//   k_xill     (vector, energy tile) (xill.cu)  corner-stationary xillver blend on the convolution grid, angle-weighted
//   k_conv     1 CTA / vector   (conv.cu)    radix-16 shared-memory FFT convolution per zone, primary spectrum,
//                               rebin to the output grid
//   k_linefinish / k_xillver / k_prim_nth: the line models' normalisation, the standalone xillver models, the nthcomp
//                               primary
// Reference code each kernel replaces is cited at the kernel.  FP64 throughout where the reference
// uses double; float where it uses float (interpolation factors).  No tensor cores: no stage is a
// dense contraction.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdio>

#include "common.h"
#include "devutil.cuh"
#include "kernels.h"

namespace rx {

// ---------------------------------------------------------------------------------- k_syspar
// Replaces interpol_relTable (src/Relprofile.cpp:141-303, the parts that do not need the g* axis),
// get_fine_radial_grid (src/relutility.c:666-677), calc_emis_profile (src/Rellp.cpp:518-564) with
// get_emis_bkn (:421-437), the lamp-post branch (:36-89,:120-285) and the returning-radiation
// emissivity (src/Relreturn_Corona.cpp:39-171,263-321; src/Relreturn_Table.cpp:396-611).
struct SysSmem {
  double rt[REL_NRT], gmin_t[REL_NRT], gmax_t[REL_NRT];
  double lrad[LP_NRT], let[LP_NRT], ldet[LP_NRT], ldit[LP_NRT];
  double re[NR], emis[NR], del_emit[NR];
  double red[256];
  double rad[RR_NR], rlo[RR_NR], rhi[RR_NR], emis_in[RR_NR], emis_ret[RR_NR], cflux[RR_NR], cgsh[RR_NR];
  double prod[RR_NR * RR_NR];
  double scal[8];
  int irad[RR_NR], irad2[RR_NR];
  int ints[8];
};

// corrected_gshift_fluxboost_factor / g (src/Relreturn_Corona.cpp:39-83), with 1/g computed once
__device__ __forceinline__ double gshift_fluxboost_over_g(double xill_gshift_fac, double g, double lng, double gamma) {
  const double g0 = 2. / 3;
  const double ig = 1. / g;
  double corr;
  if (xill_gshift_fac < 1) {
    const double a = (xill_gshift_fac / g0 - 1) / (g0 - 1);
    const double b = 1 - a;
    corr = (g >= 1) ? ig * (ig * a + b) : g * (g * a + b);
  } else {
    const double alin = (xill_gshift_fac - 1) / (g0 - 1);
    const double blin = 1 - alin;
    corr = (g >= 1) ? (ig * alin + blin) : (g * alin + blin);
  }
  double fb = exp(gamma * lng) * corr;   // pow(g, gamma) with ln g tabulated at load
  if (g < 1 && fb > 1) fb = 1;
  if (fb < 0) fb = 0;
  return fb * ig;
}

__global__ void __launch_bounds__(256) k_syspar(const VPar *__restrict__ vps, DevTables T, Scratch S, int pass) {
  extern __shared__ __align__(16) unsigned char smraw[];
  SysSmem &sm = *reinterpret_cast<SysSmem *>(smraw);
  const int v = blockIdx.x, t = threadIdx.x;
  const VPar &vp = vps[v];
  if (S.reuse && S.reuse[v]) return;   // emissivity, radial grid and status of the previous run stand
  if (pass == 1) {
    if (t == 0) S.status[v] = vp.status;
    if (vp.status != ST_OK) return;
  } else {
    if (!vp.do_corr || S.status[v] != ST_OK) return;
  }
  const double a = vp.a, rin = vp.rin, rout = vp.rout;
  double *g_re = S.re + (size_t) v * NR, *g_gmin = S.gmin + (size_t) v * NR, *g_gmax = S.gmax + (size_t) v * NR;
  double *g_emis = S.emis + (size_t) v * NR, *g_de = S.del_emit + (size_t) v * NR, *g_di = S.del_inc + (size_t) v * NR;

  // ---- (a, mu0) bracket: float arithmetic exactly like src/Relprofile.cpp:171-177
  const double mu0 = cos(vp.incl);
  const int ia = bsearch_asc<float>(T.rel_a, REL_NA, (float) a);
  const int im = bsearch_asc<float>(T.rel_mu0, REL_NMU, (float) mu0);
  const float ifac_a = ((float) a - T.rel_a[ia]) / (T.rel_a[ia + 1] - T.rel_a[ia]);
  const float ifac_mu = ((float) mu0 - T.rel_mu0[im]) / (T.rel_mu0[im + 1] - T.rel_mu0[im]);
  const size_t o00 = ((size_t) ia * REL_NMU + im) * REL_NRT, o10 = ((size_t) (ia + 1) * REL_NMU + im) * REL_NRT;
  const size_t o01 = o00 + REL_NRT, o11 = o10 + REL_NRT;

  if (t < REL_NRT) {
    sm.rt[t] = lin1d(ifac_a, T.rel_r[o00 + t], T.rel_r[o10 + t]);
    sm.gmin_t[t] = lin2d_f(ifac_a, ifac_mu, T.rel_gmin[o00 + t], T.rel_gmin[o10 + t], T.rel_gmin[o01 + t], T.rel_gmin[o11 + t]);
    sm.gmax_t[t] = lin2d_f(ifac_a, ifac_mu, T.rel_gmax[o00 + t], T.rel_gmax[o10 + t], T.rel_gmax[o01 + t], T.rel_gmax[o11 + t]);
  }
  __syncthreads();
  if (t == 0) {
    const double rms = vp.rms;
    double last = sm.rt[REL_NRT - 1];
    if ((last > rms) && ((last - rms) / last < 1e-3)) sm.rt[REL_NRT - 1] = rms;
    const int ind_rmin = bsearch_desc(sm.rt, REL_NRT, rin);
    const int ind_rmax = bsearch_desc(sm.rt, REL_NRT, rout);
    if (sm.rt[ind_rmax] < 1000.0 && sm.rt[ind_rmax] * 1.01 > 1000.0) sm.rt[ind_rmax] = 1000.0;
    sm.ints[0] = ind_rmin;
    sm.ints[1] = 0;  // error flag
    if (pass == 1) {  // the (a, mu0) bracket, reused by k_fine
      S.brk_i[2 * v] = ia;
      S.brk_i[2 * v + 1] = im;
      S.brk_f[2 * v] = (double) ifac_a;
      S.brk_f[2 * v + 1] = (double) ifac_mu;
    }
  }
  __syncthreads();
  const int ind_rmin = sm.ints[0];

  // ---- fine radial grid + radial bracket in the table
  const double r1 = 1.0 / sqrt(rout), r2 = 1.0 / sqrt(rin);
  for (int i = t; i < NR; i += 256) {
    double x = ((double) (i)) * (r2 - r1) / (NR - 1) + r1;
    x = 1.0 / x;
    const double re = x * x;
    sm.re[i] = re;
    const int K = count_gt_desc(sm.rt, REL_NRT, re) - 1;
    int it = (K < ind_rmin) ? K : ind_rmin;
    if (it < 0) {
      if (re - 1000.0 <= 1e-6) it = 0; else sm.ints[1] = ST_TABLE_RANGE;
      if (it < 0) it = 0;
    }
    const double fr = (re - sm.rt[it + 1]) / (sm.rt[it] - sm.rt[it + 1]);
    if (fr > 1.0 && it > 0) sm.ints[1] = ST_TABLE_RANGE;
    const double gmn = lin1d(fr, sm.gmin_t[it + 1], sm.gmin_t[it]);
    const double gmx = lin1d(fr, sm.gmax_t[it + 1], sm.gmax_t[it]);
    if (pass == 1) {
      g_re[i] = re;
      g_gmin[i] = gmn;
      g_gmax[i] = gmx;
      S.fr[(size_t) v * NR + i] = fr;
      S.it[(size_t) v * NR + i] = it;
      S.izone[(size_t) v * NR + i] = bsearch_asc<double>(vp.zone, vp.nz + 1, re);
    }
  }
  __syncthreads();
  if (sm.ints[1] != 0) {
    if (t == 0) S.status[v] = sm.ints[1];
    return;
  }
  if (pass == 1 && t <= vp.nz) {  // zfirst[z] = first fine-grid index whose zone is < z (zone z owns [zfirst[z+1], zfirst[z]))
    const int *iz = S.izone + (size_t) v * NR;
    int lo = 0, hi = NR;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (iz[m] >= t) lo = m + 1; else hi = m; }
    S.zfirst[(size_t) v * (NZMAX + 1) + t] = lo;
  }

  // ---- emissivity
  if (vp.emis_type == EMIS_BKN) {
    double part = 0.0;
    for (int i = t; i < NR; i += 256) {
      const double re = sm.re[i];
      double alpha = vp.emis1;
      if (re > vp.rbr) alpha = vp.emis2;
      const double e = pow(re / vp.rbr, -alpha);
      sm.emis[i] = e;
      sm.del_emit[i] = -1.0;
    }
    __syncthreads();
    for (int i = t; i < NR; i += 256) part += sm.emis[i] * (trapez_single(sm.re, i, NR) * 2);
    const double integ = block_sum<256>(part, sm.red);
    for (int i = t; i < NR; i += 256) {
      sm.emis[i] = sm.emis[i] / integ;
      g_di[i] = -1.0;
    }
    __syncthreads();
  } else {
    // lamp post: (a, h) interpolation on the table's 100 radii, src/Rellp.cpp:190-247
    const int la = bsearch_asc<float>(T.lp_a, LP_NA, (float) a);
    const double fa = (double) (((float) a - T.lp_a[la]) / (T.lp_a[la + 1] - T.lp_a[la]));
    const float hf = (float) vp.height;
    const float *h0 = T.lp_h + (size_t) la * LP_NH, *h1 = h0 + LP_NH;
    const int ih0 = bsearch_asc<float>(h0, LP_NH, hf), ih1 = bsearch_asc<float>(h1, LP_NH, hf);
    const double fh0 = (double) ((hf - h0[ih0]) / (h0[ih0 + 1] - h0[ih0]));
    const double fh1 = (double) ((hf - h1[ih1]) / (h1[ih1 + 1] - h1[ih1]));
    if (t < LP_NRT) {
      const size_t q0 = ((size_t) la * LP_NH + ih0) * LP_NRT + t, q1 = ((size_t) (la + 1) * LP_NH + ih1) * LP_NRT + t;
      sm.lrad[t] = lin1d(fa, T.lp_rad[(size_t) la * LP_NRT + t], T.lp_rad[(size_t) (la + 1) * LP_NRT + t]);
      sm.let[t] = (1.0 - fa) * lin1d(fh0, T.lp_int[q0], T.lp_int[q0 + LP_NRT]) + (fa) * lin1d(fh1, T.lp_int[q1], T.lp_int[q1 + LP_NRT]);
      sm.ldet[t] = (1.0 - fa) * lin1d(fh0, T.lp_del[q0], T.lp_del[q0 + LP_NRT]) + (fa) * lin1d(fh1, T.lp_del[q1], T.lp_del[q1 + LP_NRT]);
      sm.ldit[t] = (1.0 - fa) * lin1d(fh0, T.lp_dinc[q0], T.lp_dinc[q0 + LP_NRT]) + (fa) * lin1d(fh1, T.lp_dinc[q1], T.lp_dinc[q1 + LP_NRT]);
    }
    __syncthreads();
    // re-grid onto the fine radii, src/Rellp.cpp:120-175
    const int kk0 = bsearch_asc<double>(sm.lrad, LP_NRT, sm.re[NR - 1]);
    for (int i = t; i < NR; i += 256) {
      const double re = sm.re[i];
      int kk = count_le_asc(sm.lrad, LP_NRT, re) - 1;
      if (kk < kk0) kk = kk0;
      if (kk >= LP_NRT - 1) {
        if (!(re - 1000.0 <= 1e-6)) sm.ints[1] = ST_TABLE_RANGE;
        kk = LP_NRT - 2;
      }
      double f;
      if (sm.ldet[kk] / PI * 180.0 <= 75.0) f = (re - sm.lrad[kk]) / (sm.lrad[kk + 1] - sm.lrad[kk]);
      else f = (log(re) - log(sm.lrad[kk])) / (log(sm.lrad[kk + 1]) - log(sm.lrad[kk]));
      sm.emis[i] = exp(f * log(sm.let[kk + 1]) + (1.0 - f) * log(sm.let[kk]));
      sm.del_emit[i] = lin1d(f, sm.ldet[kk], sm.ldet[kk + 1]);
      g_di[i] = lin1d(f, sm.ldit[kk], sm.ldit[kk + 1]);
    }
    __syncthreads();
    if (t == 0 && pass == 1) {  // photon fate fractions, src/Rellp.cpp:36-89
      double del_ad_max = sm.ldet[LP_NRT - 1];
      double del_bh = sm.del_emit[bsearch_desc(sm.re, NR, rin)];
      double del_ad = sm.del_emit[bsearch_desc(sm.re, NR, rout)];
      if (del_ad_max < PI / 2.0) del_ad_max = PI / 2.0;
      if (vp.beta > 1e-6) {
        del_bh = relat_abberation(del_bh, -1. * vp.beta);
        del_ad = relat_abberation(del_ad, -1. * vp.beta);
      }
      const double f_bh = 0.5 * (1.0 - cos(del_bh));
      const double f_ad = 0.5 * (cos(del_bh) - cos(del_ad));
      const double f_inf_rest = 0.5 * (1.0 + cos(del_ad_max));
      double f_inf = f_inf_rest;
      if (vp.beta > 1e-6) f_inf = 0.5 * (1.0 + cos(relat_abberation(del_ad_max, -1. * vp.beta)));
      double *rf = S.reflfrac + (size_t) v * 8;
      rf[0] = f_ad / f_inf; rf[1] = f_bh; rf[2] = f_ad; rf[3] = f_inf; rf[4] = f_inf_rest;
    }
    for (int i = t; i < NR; i += 256) {  // flux boost source -> disk, src/Relphysics.cpp:301-311
      double boost = pow(gi_potential_lp(sm.re[i], a, vp.height, vp.beta, sm.del_emit[i]), vp.gamma);
      if (vp.beta > 1e-6) {
        const double d = doppler_factor(sm.del_emit[i], vp.beta);
        boost *= d * d;
      }
      sm.emis[i] *= boost;
    }
    __syncthreads();
  }

  // ---- returning radiation.  Pass 2 (with the correction factors) recomputes the whole emissivity, and between the
  // passes only the alpha-disk ionisation gradient reads it (k_zone): every other vector that will get a second pass
  // skips the uncorrected sum here
  if (vp.return_rad != 0 && !(pass == 1 && vp.do_corr && vp.ion_grad_type != ION_ALPHA)) {
    const int is = vp.rr_spin;
    const size_t n2 = (size_t) RR_NR * RR_NR;
    const double *rlo_t = T.rr_rlo + (size_t) is * RR_NR, *rhi_t = T.rr_rhi + (size_t) is * RR_NR;
    const bool have_corr = (pass == 2);
    const double rlo_e = sm.re[NR - 1], rhi_e = sm.re[0];
    if (t == 0) {
      int klo = bsearch_asc<double>(rlo_t, RR_NR, rlo_e);
      int khi = bsearch_asc<double>(rhi_t, RR_NR, rhi_e);
      if (fabs(rhi_e - rhi_t[RR_NR - 1]) < 1e-6) khi = RR_NR - 1; else khi++;
      if (fabs(rlo_e - rlo_t[0]) < 1e-6) klo = 0;
      int nrad = (khi + 1) - klo;
      if (nrad < 2 || nrad > RR_NR) { sm.ints[1] = ST_RRAD; nrad = 2; }
      sm.ints[2] = nrad;
      sm.ints[3] = klo;
    }
    __syncthreads();
    const int nrad = sm.ints[2];
    {
      const int klo = sm.ints[3];
      if (t < nrad) {   // one thread per ring
        const double rlo = (t == 0) ? rlo_e : rlo_t[klo + t], rhi = (t == nrad - 1) ? rhi_e : rhi_t[klo + t];
        sm.irad[t] = klo + t;
        sm.rlo[t] = rlo;
        sm.rhi[t] = rhi;
        const double rad = 0.5 * (rlo + rhi);
        sm.rad[t] = rad;
        sm.emis_in[t] = 0.0;
        if (have_corr) {  // correction factors by zone membership of the ring centre, Relreturn_Corona.cpp:233-247
          int ind = bsearch_asc<double>(vp.zone, vp.nz + 1, rad);
          if (rad < vp.zone[0]) ind = 0;
          else if (rad > vp.zone[vp.nz]) ind = vp.nz;  // (one past the end in the reference as well)
          if (ind >= vp.nz) ind = vp.nz - 1;
          sm.cflux[t] = S.corr_flux[(size_t) v * NZMAX + ind];
          sm.cgsh[t] = S.corr_gshift[(size_t) v * NZMAX + ind];
        }
      } else if (t >= 64 && t < 66) {  // ring-area correction of the two partially covered edge rings
        const int s = t - 64;
        const int idx = s == 0 ? 0 : nrad - 1;
        const double rlo_m = (idx == 0) ? rlo_e : rlo_t[klo + idx], rhi_m = (idx == nrad - 1) ? rhi_e : rhi_t[klo + idx];
        double rlo_tab = rlo_t[klo + idx];
        if (klo + idx == 0 && rlo_tab > vp.rms) rlo_tab = vp.rms;
        const double rhi_tab = rhi_t[klo + idx];
        const double area_table = 0.5 * (rlo_tab + rhi_tab) * (rhi_tab - rlo_tab);
        const double area_model = 0.5 * (rlo_m + rhi_m) * (rhi_m - rlo_m);
        sm.scal[s] = area_model / area_table;
      }
    }
    __syncthreads();
    if (t == 0) {
      if (sm.rad[0] > sm.rad[nrad - 1] || sm.re[nrad - 1] > sm.re[0]) sm.ints[1] = ST_RRAD;
      if (sm.rad[0] < sm.re[NR - 1] || sm.rad[nrad - 1] > sm.re[0]) sm.ints[1] = ST_RRAD;
    }
    // emissivity at the ring centres: inv_rebin_mean (src/relutility.c:636-663).  Its cursor walk finds, for
    // the ring centres in descending order, brackets in strictly ascending fine-grid index; a centre whose
    // bracket does not advance (two centres in one fine bin) and all later ones stay unset (0).
    if (t < nrad) {
      const int c = count_gt_desc(sm.re, NR, sm.rad[t]);
      sm.irad2[t] = (c >= 1 && c <= NR - 1) ? c - 1 : -1;
    }
    __syncthreads();
    if (t < nrad && sm.ints[1] == 0) {
      bool ok = true;
      int prev = -1;
      for (int in = nrad - 1; in >= t && ok; in--) {
        if (sm.irad2[in] < 0 || sm.irad2[in] <= prev) ok = false;
        prev = sm.irad2[in];
      }
      if (ok) {
        const int ii = sm.irad2[t];
        const double f = (sm.rad[t] - sm.re[ii + 1]) / (sm.re[ii] - sm.re[ii + 1]);
        sm.emis_in[t] = lin1d(f, sm.emis[ii + 1], sm.emis[ii]);
      }
    }
    __syncthreads();
    const double *tf_t = T.rr_tf + is * n2, *gmin_t = T.rr_gmin + is * n2, *gmax_t = T.rr_gmax + is * n2;
    const double2 *fgl_t = reinterpret_cast<const double2 *>(T.rr_fgl) + (size_t) is * RR_NG * n2;
    for (int pr = t; pr < nrad * nrad; pr += 256) {
      const int io = pr / nrad, ie = pr - io * nrad;
      const size_t q = (size_t) sm.irad[io] * RR_NR + sm.irad[ie];
      const double cg = have_corr ? sm.cgsh[ie] : 1.0;
      const double gmn = gmin_t[q], gmx = gmax_t[q];
      const bool corr = fabs(cg - 1) > 1e-3;
      double ez = 0.0;
#pragma unroll 4
      for (int jj = 0; jj < RR_NG; jj++) {
        const double g = ((jj + 0.5) / RR_NG) * (gmx - gmn) + gmn;
        const double2 fl = __ldg(fgl_t + (size_t) jj * n2 + q);   // {frac_g, ln g}
        double e1 = fl.x;
        if (corr) e1 *= gshift_fluxboost_over_g(cg, g, fl.y, vp.gamma);
        else if (fabs(g - 1) > 1e-3) e1 *= exp((vp.gamma - 1) * fl.y);
        ez += e1;
      }
      double tfr = tf_t[q];
      if (ie == 0 && io < nrad - 1) tfr *= sm.scal[0];
      if (ie == nrad - 1 && io < nrad - 2) tfr *= sm.scal[1];
      sm.prod[io * RR_NR + ie] = ez * tfr * sm.emis_in[ie];
    }
    __syncthreads();
    if (t < nrad) {
      double sum = 0.0;
      for (int ie = 0; ie < nrad; ie++) sum += sm.prod[t * RR_NR + ie];
      if (have_corr) sum *= sm.cflux[t];
      sm.emis_ret[t] = sum;
    }
    __syncthreads();
    // back onto the fine grid (log interpolation in emissivity, linear in radius), add
    const int kk0 = bsearch_asc<double>(sm.rad, nrad, sm.re[NR - 1]);
    for (int i = t; i < NR; i += 256) {
      const double re = sm.re[i];
      int kk = count_le_asc(sm.rad, nrad, re) - 1;
      if (kk < kk0) kk = kk0;
      if (kk >= nrad - 1) {
        if (!(re - 1000.0 <= 1e-6)) sm.ints[1] = ST_TABLE_RANGE;
        kk = nrad - 2;
      }
      const double f = (re - sm.rad[kk]) / (sm.rad[kk + 1] - sm.rad[kk]);
      const double er = exp(f * log(sm.emis_ret[kk + 1]) + (1.0 - f) * log(sm.emis_ret[kk]));
      if (vp.return_rad == 1) sm.emis[i] += er; else sm.emis[i] = er;
    }
    __syncthreads();
  }
  for (int i = t; i < NR; i += 256) {
    g_emis[i] = sm.emis[i];
    if (pass == 1) g_de[i] = sm.del_emit[i];
  }
  if (t == 0 && sm.ints[1] != 0) S.status[v] = sm.ints[1];
}

// ---------------------------------------------------------------------------------- k_zone
// Replaces IonGradient::calculate_gradient (src/IonGradient.cpp:128-211,340-390,421-430),
// get_xilltab_indices_for_paramvals + the interpolation-factor part of interp_xill_table
// (src/xilltable.c:297-323,1054-1136), calc_xillver_normalization_change_source_to_disk and
// calc_normalization_factor_source (src/Xillspec.cpp:179-233,408-440; src/PrimarySource.h:279-293) and
// calc_rrad_corr_factors (src/Relxill.cpp:164-182; src/Xillspec.cpp:344-362,457-526).
struct ZoneSmem {
  double re[NR], y1[NR], y2[NR], y3[NR];
  double powtab[NCOARSE];
  double rmean[NZMAX], del_emit[NZMAX], irr[NZMAX], dinc[NZMAX], lxi[NZMAX], dens[NZMAX], ect[NZMAX], eshift[NZMAX];
  double nfac[NZMAX + 1], s2[NZMAX + 1];
  int b[NZMAX];
  int ints[4];
};

// cutoff power law on the coarse grid (src/Xillspec.cpp:215-233) reduced to the two band sums; warp-cooperative
__device__ void ecut_band_sums(const DevTables &T, const double *powtab, double ecut, double &S1, double &S2) {
  const int lane = threadIdx.x & 31;
  double a1 = 0.0, a2 = 0.0;
  const double ex0 = exp(1.0 / ecut);
  for (int i = lane; i < NCOARSE; i += 32) {
    const double e0 = T.ecoarse[i], e1 = T.ecoarse[i + 1];
    const double en = 0.5 * (e0 + e1);
    const double fl = ex0 * powtab[i] * exp(-en / ecut) * (e1 - e0);
    const double w = fl * 0.5 * (e0 + e1);
    if (T.coarse_m1[i]) a1 += w * 1e20 * 1.602177e-09;
    if (T.coarse_m2[i]) a2 += w;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  S1 = a1;
  S2 = a2;
}

// blackbody primary of the NS flavours (spec_blackbody, src/Xillspec.cpp:262-269) reduced to the same two band sums
__device__ void bb_band_sums(const DevTables &T, double ktbb, double &S1, double &S2) {
  const int lane = threadIdx.x & 31;
  double a1 = 0.0, a2 = 0.0;
  const double kt4 = pow(ktbb, 4);
  for (int i = lane; i < NCOARSE; i += 32) {
    const double e0 = T.ecoarse[i], e1 = T.ecoarse[i + 1];
    const double en = 0.5 * (e0 + e1);
    double fl = en * en / (kt4 * (exp(en / ktbb) - 1));
    fl *= (e1 - e0);
    const double w = fl * 0.5 * (e0 + e1);
    if (T.coarse_m1[i]) a1 += w * 1e20 * 1.602177e-09;
    if (T.coarse_m2[i]) a2 += w;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  S1 = a1;
  S2 = a2;
}

__global__ void __launch_bounds__(128) k_zone(const VPar *__restrict__ vps, DevTables T, Scratch S) {
  extern __shared__ __align__(16) unsigned char smraw[];
  ZoneSmem &sm = *reinterpret_cast<ZoneSmem *>(smraw);
  const int v = blockIdx.x, t = threadIdx.x;
  const VPar &vp = vps[v];
  if (S.status[v] != ST_OK) return;
  if (S.reuse && (S.reuse[v] & REUSE_ALL)) return;
  const int nz = vp.nz;
  const bool alpha = (vp.ion_grad_type == ION_ALPHA);
  const int nc_all = (T.xill[vp.xtab].npar == 6) ? 32 : 16;
  for (int i = t; i < NR; i += 128) {
    sm.re[i] = S.re[(size_t) v * NR + i];
    sm.y1[i] = S.del_emit[(size_t) v * NR + i];
    if (alpha) {
      sm.y2[i] = S.emis[(size_t) v * NR + i];
      sm.y3[i] = S.del_inc[(size_t) v * NR + i];
    }
  }
  if (t == 0) sm.ints[0] = 0;
  __syncthreads();
  // brackets of the zone centres in the (descending) fine grid
  if (t < nz) {
    const double x = 0.5 * (vp.zone[t] + vp.zone[t + 1]);
    sm.rmean[t] = x;
    const int c = count_gt_desc(sm.re, NR, x);
    sm.b[t] = (c >= 1 && c <= NR - 1) ? c - 1 : -1;
  }
  __syncthreads();
  if (t == 0) {  // inv_rebin_mean's cursor semantics: brackets must be found in strictly increasing order
    bool ok = true;
    if (sm.rmean[0] > sm.rmean[nz - 1] || sm.re[nz - 1] > sm.re[0]) ok = false;
    if (sm.rmean[0] < sm.re[NR - 1] || sm.rmean[nz - 1] > sm.re[0]) ok = false;
    int prev = -1;
    for (int in = nz - 1; in >= 0 && ok; in--) {
      if (sm.b[in] < 0 || sm.b[in] <= prev) ok = false;
      prev = sm.b[in];
    }
    if (!ok) sm.ints[0] = ST_ZONE;
  }
  __syncthreads();
  if (sm.ints[0] != 0) {
    if (t == 0) S.status[v] = sm.ints[0];
    return;
  }
  if (t < nz) {
    const int ii = sm.b[t];
    const double f = (sm.rmean[t] - sm.re[ii + 1]) / (sm.re[ii] - sm.re[ii + 1]);
    sm.del_emit[t] = lin1d(f, sm.y1[ii + 1], sm.y1[ii]);
    if (alpha) {
      sm.irr[t] = lin1d(f, sm.y2[ii + 1], sm.y2[ii]);
      sm.dinc[t] = lin1d(f, sm.y3[ii + 1], sm.y3[ii]);
    }
    sm.eshift[t] = (vp.emis_type == EMIS_LP) ? gi_potential_lp(sm.rmean[t], vp.a, vp.height, vp.beta, sm.del_emit[t]) : 1.0;
  }
  __syncthreads();
  if (t < nz) {
    double lxi, dens;
    if (vp.ion_grad_type == ION_PL) {  // src/IonGradient.cpp:174-182 (natural log/exp as written there)
      lxi = (exp(vp.lxi)) * pow((sm.rmean[t] / sm.rmean[0]), -1.0 * vp.iongrad_index);
      lxi = log(lxi);
      dens = vp.dens;
    } else if (alpha) {  // src/IonGradient.cpp:108-171
      const double rin = vp.zone[0];
      const double rad_lxi = ((11. / 9.) * (11. / 9.)) * rin;
      const int kk = bsearch_desc(sm.re, NR, rad_lxi);
      const double interp = (rad_lxi - sm.re[kk + 1]) / (sm.re[kk] - sm.re[kk + 1]);
      const double e_at = lin1d(interp, sm.y2[kk + 1], sm.y2[kk]);
      const double di_at = lin1d(interp, sm.y3[kk + 1], sm.y3[kk]);
      const double lxi_max = log10(4.0 * PI * e_at / density_ss73_zone_a(rad_lxi, rin) * (cos(PI / 4) / cos(di_at)));
      const double fac_lxi_norm = vp.lxi - lxi_max;
      const double density_min = density_ss73_zone_a((25. / 9.) * rin, rin);
      const double dn = vp.const_density ? 1.0 : density_ss73_zone_a(sm.rmean[t], rin) / density_min;   // src/IonGradient.cpp:155-158
      dens = log10(dn) + vp.dens;
      lxi = log10(4.0 * PI * sm.irr[t] / dn * (cos(PI / 4) / cos(sm.dinc[t])));
      lxi += fac_lxi_norm;
    } else {
      lxi = vp.lxi;
      dens = vp.dens;
    }
    if (lxi < 0.0) lxi = 0.0; else if (lxi > 4.7) lxi = 4.7;  // src/IonGradient.cpp:388
    sm.lxi[t] = lxi;
    sm.dens[t] = dens;
    sm.ect[t] = vp.ect * sm.eshift[t];
    S.eshift[(size_t) v * NZMAX + t] = sm.eshift[t];
    S.zlxi[(size_t) v * NZMAX + t] = lxi;
    S.zdens[(size_t) v * NZMAX + t] = dens;
    S.zect[(size_t) v * NZMAX + t] = sm.ect[t];

    // xillver corner nodes + weights of this zone
    const XillDev &X = T.xill[vp.xtab];
    float inp[8];
    inp[0] = (float) vp.gam; inp[1] = (float) vp.afe; inp[2] = (float) lxi; inp[3] = (float) sm.ect[t];
    inp[4] = (float) dens; inp[5] = (float) vp.ktbb; inp[6] = (float) vp.frac_pl_bb; inp[7] = 0.f;
    int ind[6];
    double fac[6];
    const int nax = X.npar - 1;  // all axes but the inclination
    for (int i = 0; i < nax; i++) {
      const int pind = X.pindex[i];
      const int n = X.nvals[i];
      int k = bsearch_asc<float>(X.vals[i], n, inp[pind]);
      if (k < 0) k = 0; else if (k > n - 2) k = n - 2;
      ind[i] = k;
      float val = inp[pind];
      const float lo = X.vals[i][0], hi = X.vals[i][n - 1];
      if (val < lo) val = lo; else if (val > hi) val = hi;
      fac[i] = (double) ((val - X.vals[i][k]) / (X.vals[i][k + 1] - X.vals[i][k]));
      if (pind == 3) {  // ensure_ecut_within_boundarys (src/xilltable.c:1054-1067): the reference looks the limits up on table
                        // axis number 3 (the global Ecut id), which is the Ecut axis itself except in the CO table
        if (sm.ect[t] <= (double) X.vals[3][0]) fac[i] = 0.0;
        if (sm.ect[t] >= (double) X.vals[3][X.nvals[3] - 1]) fac[i] = 1.0;
      }
    }
    const int off = (X.npar == 6) ? 1 : 0;
    const double f1 = fac[off], f2 = fac[off + 1], f3 = fac[off + 2], f4 = fac[off + 3];
    int *xr = S.xrow + ((size_t) v * NZMAX + t) * 32;
    double *xw = S.xw + ((size_t) v * NZMAX + t) * 32;
    // corner order of interp_5d_tab_incl (src/xilltable.c:836-873)
    const int bits[16] = {0x0, 0x1, 0x2, 0x4, 0x3, 0x5, 0x6, 0x7, 0x8, 0x9, 0xA, 0xC, 0xB, 0xD, 0xE, 0xF};
    for (int half = 0; half < (X.npar == 6 ? 2 : 1); half++) {
      for (int c = 0; c < 16; c++) {
        const int b1 = bits[c] & 1, b2 = (bits[c] >> 1) & 1, b3 = (bits[c] >> 2) & 1, b4 = (bits[c] >> 3) & 1;
        double w = (b1 ? f1 : (1.0 - f1)) * (b2 ? f2 : (1.0 - f2)) * (b3 ? f3 : (1.0 - f3)) * (b4 ? f4 : (1 - f4));
        long node;
        const int i1 = ind[off] + b1, i2 = ind[off + 1] + b2, i3 = ind[off + 2] + b3, i4 = ind[off + 3] + b4;
        if (X.npar == 5) node = (((long) i1 * X.nvals[1] + i2) * X.nvals[2] + i3) * X.nvals[3] + i4;
        else node = ((((long) (ind[0] + half) * X.nvals[1] + i1) * X.nvals[2] + i2) * X.nvals[3] + i3) * X.nvals[4] + i4;
        if (X.npar == 6) w = half ? fac[0] * w : (1.0 - fac[0]) * w;
        xr[half * 16 + c] = (int) node;
        xw[half * 16 + c] = w;
      }
    }
  }
  __syncthreads();
  // Factorised corner list for k_xill.  The multilinear weight of a corner is a product over the table
  // axes; Gamma and A_Fe are the same for all zones of a vector, so the table is first contracted over
  // those two axes (4 nodes, weights ga_w) and the zones only blend the remaining "rest" corners
  // (logXi, Ecut|kTe[, Dens]: 4 or 8 per zone).  A corner is filed under the slot given by the parities of
  // its node indices (the two nodes of a bracket have different parity, so the corners of a zone occupy
  // distinct slots, and a corner shared by neighbouring zones keeps its slot): xkey[zone][slot] = rest node
  // offset, xwsort[zone][slot] = weight.
  const int n_rest = nc_all / 4;
  if (t < nz) {
    const XillDev &X = T.xill[vp.xtab];
    const int nax = X.npar - 1;
    long stride[6];
    long acc_s = 1;
    for (int i = nax - 1; i >= 0; i--) { stride[i] = acc_s; acc_s *= X.nvals[i]; }
    // recompute this zone's bracket (same arithmetic as above) for the zone-varying axes
    float inp[8];
    inp[0] = (float) vp.gam; inp[1] = (float) vp.afe; inp[2] = (float) sm.lxi[t]; inp[3] = (float) sm.ect[t];
    inp[4] = (float) sm.dens[t]; inp[5] = (float) vp.ktbb; inp[6] = (float) vp.frac_pl_bb; inp[7] = 0.f;
    int ind[6];
    double fac[6];
    for (int i = 0; i < nax; i++) {
      const int pind = X.pindex[i];
      const int n = X.nvals[i];
      int k = bsearch_asc<float>(X.vals[i], n, inp[pind]);
      if (k < 0) k = 0; else if (k > n - 2) k = n - 2;
      ind[i] = k;
      float val = inp[pind];
      const float lo = X.vals[i][0], hi = X.vals[i][n - 1];
      if (val < lo) val = lo; else if (val > hi) val = hi;
      fac[i] = (double) ((val - X.vals[i][k]) / (X.vals[i][k + 1] - X.vals[i][k]));
      if (pind == 3) {
        if (sm.ect[t] <= (double) X.vals[3][0]) fac[i] = 0.0;
        if (sm.ect[t] >= (double) X.vals[3][X.nvals[3] - 1]) fac[i] = 1.0;
      }
    }
    // axes: two vector-level ones (the first two whose parameter cannot change from zone to zone: Gamma and A_Fe,
    // kTbb and A_Fe for the NS table, Gamma and A_CO for the CO table) and the rest (logXi, Ecut|kTe, Dens, ...)
    int ax_v[2], ax_r[4], nv = 0, nr = 0;
    for (int i = 0; i < nax; i++) {
      const int pi = X.pindex[i];
      if (nv < 2 && pi != 2 && pi != 3 && pi != 4) ax_v[nv++] = i; else ax_r[nr++] = i;
    }
    for (int c = 0; c < n_rest; c++) {
      long off = 0;
      double w = 1.0;
      for (int q = 0; q < nr; q++) {
        const int bit = (c >> q) & 1;
        off += (long) (ind[ax_r[q]] + bit) * stride[ax_r[q]];
        w *= bit ? fac[ax_r[q]] : (1.0 - fac[ax_r[q]]);
      }
      int slot = 0;
      for (int q = 0; q < nr; q++) slot |= ((ind[ax_r[q]] + ((c >> q) & 1)) & 1) << q;
      S.xkey[((size_t) v * NZMAX + t) * 8 + slot] = (int) off;
      S.xwsort[((size_t) v * NZMAX + t) * 8 + slot] = w;
    }
    if (t == 0) {
      int *ga_off = S.xga_off + (size_t) v * 4;
      double *ga_w = S.xga_w + (size_t) v * 4;
      for (int c = 0; c < 4; c++) {
        const int b0 = c & 1, b1 = (c >> 1) & 1;
        ga_off[c] = (int) ((long) (ind[ax_v[0]] + b0) * stride[ax_v[0]] + (long) (ind[ax_v[1]] + b1) * stride[ax_v[1]]);
        ga_w[c] = (b0 ? fac[ax_v[0]] : (1.0 - fac[ax_v[0]])) * (b1 ? fac[ax_v[1]] : (1.0 - fac[ax_v[1]]));
      }
    }
  }
  __syncthreads();
  if (t == 0) S.xn[v] = n_rest;
  // primary-spectrum normalisations: one warp per (zone | source); cutoff power law only here,
  // the nthcomp variant lives in k_zone_nthcomp
  if (vp.prim_type == PRIM_ECUT || vp.prim_type == PRIM_BB) {
    const bool bb = (vp.prim_type == PRIM_BB);
    for (int i = t; i < NCOARSE; i += 128) {
      const double en = 0.5 * (T.ecoarse[i] + T.ecoarse[i + 1]);
      sm.powtab[i] = bb ? 0.0 : pow(en, -vp.gam);
    }
    __syncthreads();
    const int warp = t >> 5, lane = t & 31;
    for (int job = warp; job <= nz; job += 4) {
      const double ecut = (job < nz) ? sm.ect[job] : vp.ect;
      double S1, S2;
      if (bb) bb_band_sums(T, vp.ktbb, S1, S2); else ecut_band_sums(T, sm.powtab, ecut, S1, S2);
      if (lane == 0) {
        const double norm = S1 / (1e15 / 4.0 / PI);
        sm.nfac[job] = 1. / norm;
        sm.s2[job] = S2;
      }
    }
    __syncthreads();
    if (t == 0) S.nsrc[v] = sm.nfac[nz];
    if (t < nz) {
      S.normch[(size_t) v * NZMAX + t] = sm.nfac[t] / sm.nfac[nz];
      if (vp.do_corr) {
        const XillDev &X = T.xill[vp.xtab];
        const int *xr = S.xrow + ((size_t) v * NZMAX + t) * 32;
        const double *xw = S.xw + ((size_t) v * NZMAX + t) * 32;
        double ef = 0.0, p1 = 0.0, p2 = 0.0;
        const int nc = (X.npar == 6) ? 32 : 16;
        for (int c = 0; c < nc; c++) {
          ef += xw[c] * X.node_ef[xr[c]];
          p1 += xw[c] * X.node_p1[xr[c]];
          p2 += xw[c] * X.node_p2[xr[c]];
        }
        const double direct = sm.s2[t] * sm.nfac[t];
        S.corr_flux[(size_t) v * NZMAX + t] = ef / direct;
        S.corr_gshift[(size_t) v * NZMAX + t] = (p1 / p2) / pow(1.5, vp.gam);
      }
    }
  }
}

// ---------------------------------------------------------------------------------- k_rows
// The (a, mu0) half of the transfer-function interpolation (interpol_a_mu0, src/Relprofile.cpp:39-80), once per TABLE
// radius: the fine grid has about ten radii per table interval, and every one of them needs the same two bilinearly
// interpolated rows.  One thread per (table radius, g*): 16 table floats -> {trff1, trff2, cosne1, cosne2} in fp64
// (exactly the values k_fine used to recompute per fine radius; the float -> double conversions of the 32 table
// values per fine-grid point kept the XU pipe busier than the FP64 pipe).
__global__ void __launch_bounds__(320) k_rows(const VPar *__restrict__ vps, DevTables T, Scratch S) {
  const int v = blockIdx.y;
  if (S.status[v] != ST_OK) return;
  if (S.reuse && S.reuse[v]) return;
  const int j = threadIdx.x % NG, it = blockIdx.x * 8 + threadIdx.x / NG;
  if (it >= REL_NRT) return;
  const int ia = S.brk_i[2 * v], im = S.brk_i[2 * v + 1];
  const double fa = S.brk_f[2 * v], fm = S.brk_f[2 * v + 1];
  const float4 *tc = reinterpret_cast<const float4 *>(T.rel_tc);
  const size_t o00 = (((size_t) ia * REL_NMU + im) * REL_NRT + it) * NG + j;
  const size_t o10 = (((size_t) (ia + 1) * REL_NMU + im) * REL_NRT + it) * NG + j;
  const size_t o01 = (((size_t) ia * REL_NMU + im + 1) * REL_NRT + it) * NG + j;
  const size_t o11 = (((size_t) (ia + 1) * REL_NMU + im + 1) * REL_NRT + it) * NG + j;
  const float4 a00 = __ldg(tc + o00), a10 = __ldg(tc + o10), a01 = __ldg(tc + o01), a11 = __ldg(tc + o11);
  double2 tr, co;
  tr.x = lin2d_f(fa, fm, a00.x, a10.x, a01.x, a11.x);
  tr.y = lin2d_f(fa, fm, a00.y, a10.y, a01.y, a11.y);
  co.x = lin2d_f(fa, fm, a00.z, a10.z, a01.z, a11.z);
  co.y = lin2d_f(fa, fm, a00.w, a10.w, a01.w, a11.w);
  // two planes per vector, [REL_NRT][NG] {branch 0, branch 1} each: the transfer functions (k_line stages runs of these
  // rows with one bulk copy), then the emission angles
  double2 *row = reinterpret_cast<double2 *>(S.relrow) + (size_t) v * REL_NRT * NG * 2 + (size_t) it * NG + j;
  row[0] = tr;
  row[REL_NRT * NG] = co;
}

// ---------------------------------------------------------------------------------- k_fine
// The g*-dependent half of interpol_relTable (src/Relprofile.cpp:39-80,280-293): bilinear (a, mu0) blend of
// the four table corners at the two bracketing table radii, then the radial lerp, for trff1/2 and cosne1/2.
// Fused with the per-radius part of the emission-angle distribution (the rel_cosne part of
// calc_relline_profile, src/Relprofile.cpp:907-938, get_cosne_bin :799-801) while the values are in registers:
// every (radius, g*) thread files its two branch contributions, then one thread per (radius, angle bin) adds
// them in the reference's order (g* ascending, branch 1 then 2).  The emission angles themselves are stored
// only when a later stage needs them (limb darkening in k_line, or the test probes).
__global__ void __launch_bounds__(320) k_fine(const VPar *__restrict__ vps, DevTables T, Scratch S, int n_incl,
                                              double e_first, double e_last, int store_cosne, int store_trff) {
  __shared__ double s_vx[8][NG], s_vy[8][NG];       // the two branch contributions of every (radius, g*)
  __shared__ int s_bin[8][NG];                      // their angle bins, 16 bits each (0xffff: none)
  __shared__ double s_rad[8][3];                    // per radius: gmin, gmax - gmin, r (2 pi r)^2 emis weight (< 0: off the grid)
  const int v = blockIdx.y;
  if (S.status[v] != ST_OK) return;
  if (S.reuse && S.reuse[v]) return;
  const int j = threadIdx.x % NG, rl = threadIdx.x / NG;
  const int i = blockIdx.x * 8 + rl;
  if (n_incl > 0 && threadIdx.x < 8) {   // per-radius factors of the distribution
    const int i2 = blockIdx.x * 8 + threadIdx.x;
    const double *re = S.re + (size_t) v * NR;
    const double gmin = S.gmin[(size_t) v * NR + i2], gmax = S.gmax[(size_t) v * NR + i2];
    const double r = re[i2], x1 = 2 * PI * r;
    s_rad[threadIdx.x][0] = gmin;
    s_rad[threadIdx.x][1] = gmax - gmin;
    s_rad[threadIdx.x][2] = ((gmax > e_first) && (gmin < e_last))
                                ? r * (x1 * x1) * S.emis[(size_t) v * NR + i2] * (trapez_single(re, i2, NR) / 2) : -1.0;
  }
  const int it = S.it[(size_t) v * NR + i];
  const double fr = S.fr[(size_t) v * NR + i];
  // the two table rows around this radius, already interpolated in (a, mu0) by k_rows: row `it` (larger radius) and
  // row `it+1` (smaller radius)
  const double2 *row = reinterpret_cast<const double2 *>(S.relrow) + (size_t) v * REL_NRT * NG * 2 + (size_t) it * NG + j;
  const double2 thi = row[0], chi = row[REL_NRT * NG], tlo = row[NG], clo = row[REL_NRT * NG + NG];
  const double t1_hi = thi.x, t2_hi = thi.y, c1_hi = chi.x, c2_hi = chi.y;
  const double t1_lo = tlo.x, t2_lo = tlo.y, c1_lo = clo.x, c2_lo = clo.y;
  double2 tr, co;
  tr.x = lin1d(fr, t1_lo, t1_hi);
  tr.y = lin1d(fr, t2_lo, t2_hi);
  co.x = lin1d(fr, c1_lo, c1_hi);
  co.y = lin1d(fr, c2_lo, c2_hi);
  const size_t o = ((size_t) v * NR + i) * NG + j;
  if (store_trff) reinterpret_cast<double2 *>(S.trff)[o] = tr;   // probes only: k_line interpolates its own rows
  if (store_cosne) reinterpret_cast<double2 *>(S.cosne)[o] = co;
  if (n_incl <= 0) return;
  __syncthreads();
  // ---- emission-angle distribution, per-radius part: contribution = r (2 pi g r)^2 / sqrt(g* - g*^2) trff emis
  //      weight dg*, grouped as [per radius] * g^2 * [per g*] * trff
  {
    int bins = 0xffffffff;
    double2 val = make_double2(0.0, 0.0);
    const double ar = s_rad[rl][2];
    if (ar >= 0.0) {
      const double g = T.gstar[j] * s_rad[rl][1] + s_rad[rl][0];
      const double common = ar * (g * g) * T.gstar_w[j];
      val.x = common * tr.x;
      val.y = common * tr.y;
      const int i0 = ((int) (n_incl * (1 - co.x) + 1)) - 1, i1 = ((int) (n_incl * (1 - co.y) + 1)) - 1;
      const bool bad = (val.x != val.x) || (val.y != val.y) || i0 < 0 || i0 >= n_incl || i1 < 0 || i1 >= n_incl;
      if (bad) S.status[v] = ST_NAN;   // any writer stores the same value
      else bins = i0 | (i1 << 16);
    }
    s_bin[rl][j] = bins;
    s_vx[rl][j] = val.x;
    s_vy[rl][j] = val.y;
  }
  __syncthreads();
  {
    // (radius, angle bin, quarter of the g* axis) per thread: 8 x 10 x 4 = 320; the quarters are then added in order
    const int c = threadIdx.x & 3, rm = threadIdx.x >> 2;
    const int r2 = rm / 10, m = rm - r2 * 10;
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < NG / 4; q++) {
      const int jj = c * (NG / 4) + q;
      const int b = s_bin[r2][jj];
      // the values are read only where the bin matches (one or two of a thread's twenty candidates): the kernel runs the
      // shared-memory data pipe, and an unconditional 128-bit read per candidate was most of its wavefronts
      if ((b & 0xffff) == m) s += s_vx[r2][jj];
      if ((b >> 16) == m) s += s_vy[r2][jj];
    }
    const double s1 = __shfl_down_sync(0xffffffffu, s, 1), s2 = __shfl_down_sync(0xffffffffu, s, 2),
                 s3 = __shfl_down_sync(0xffffffffu, s, 3);
    if (c == 0 && m < n_incl) S.distpart[((size_t) v * NR + blockIdx.x * 8 + r2) * 10 + m] = ((s + s1) + s2) + s3;
  }
}

// ---------------------------------------------------------------------------------- k_dist
// Emission-angle distribution per zone: sums of k_fine's per-radius parts over the zone's radii and the
// normalisation (src/Relprofile.cpp:783-795).
__global__ void __launch_bounds__(256) k_dist(const VPar *__restrict__ vps, DevTables T, Scratch S, int n_incl) {
  const int v = blockIdx.x, t = threadIdx.x;
  if (S.status[v] != ST_OK) return;
  if (S.reuse && S.reuse[v]) return;
  const VPar &vp = vps[v];
  const double *part = S.distpart + (size_t) v * NR * 10;
  const int *zfirst = S.zfirst + (size_t) v * (NZMAX + 1);
  const int nz = vp.nz;
  // Zones of up to DIST_RUN radii: one thread per (zone, angle bin) adds the zone's radii in ascending index order.
  // Longer zones (vectors with few zones: a one-zone model has 1000 radii, and ten threads of the block were walking
  // them) are cut into runs of DIST_RUN radii, one thread per (bin, run), and the runs are added in order — a function of
  // the zone's own length, so the result does not depend on the batch.
  constexpr int DIST_RUN = 64, DIST_MAXRUN = NR / DIST_RUN + NZMAX + 1;
  __shared__ double s_run[DIST_MAXRUN][MAX_INCL];
  __shared__ short s_rz[DIST_MAXRUN], s_rb[NZMAX + 1];   // zone of every run, first run of every zone
  int longest = 0;
  for (int z = 0; z < nz; z++) longest = max(longest, zfirst[z] - zfirst[z + 1]);
  if (longest <= DIST_RUN) {
    for (int job = t; job < nz * n_incl; job += 256) {
      const int z = job / n_incl, m = job - z * n_incl;
      double s = 0.0;
      for (int i = zfirst[z + 1]; i < zfirst[z]; i++) s += part[(size_t) i * 10 + m];   // the zone's radii, ascending index
      S.dist[((size_t) v * NZMAX + z) * MAX_INCL + m] = s;
    }
  } else {
    if (t == 0) {
      int r = 0;
      for (int z = 0; z < nz; z++) {
        s_rb[z] = (short) r;
        const int nrun = (zfirst[z] - zfirst[z + 1] + DIST_RUN - 1) / DIST_RUN;
        for (int q = 0; q < nrun; q++) s_rz[r++] = (short) z;
      }
      s_rb[nz] = (short) r;
    }
    __syncthreads();
    const int nrun = s_rb[nz];
    for (int job = t; job < nrun * n_incl; job += 256) {
      const int r = job / n_incl, m = job - r * n_incl;
      const int z = s_rz[r], i0 = zfirst[z + 1] + (r - s_rb[z]) * DIST_RUN, i1 = min(zfirst[z], i0 + DIST_RUN);
      double s = 0.0;
      for (int i = i0; i < i1; i++) s += part[(size_t) i * 10 + m];
      s_run[r][m] = s;
    }
    __syncthreads();
    for (int job = t; job < nz * n_incl; job += 256) {
      const int z = job / n_incl, m = job - z * n_incl;
      double s = 0.0;
      for (int r = s_rb[z]; r < s_rb[z + 1]; r++) s += s_run[r][m];
      S.dist[((size_t) v * NZMAX + z) * MAX_INCL + m] = s;
    }
  }
  __syncthreads();
  if (t < nz) {
    double *d = S.dist + ((size_t) v * NZMAX + t) * MAX_INCL;
    double s = 0.0;
    for (int m = 0; m < n_incl; m++) s += d[m];
    if (!(s > 1e-8)) S.status[v] = ST_ZONE;  // the reference asserts here (src/Relprofile.cpp:790)
    for (int m = 0; m < n_incl; m++) d[m] /= s;
  }
}

// ---------------------------------------------------------------------------------- k_linefinish
// renorm_relline_profile for one-zone line models (src/Relprofile.cpp:749-781) + copy to the output
__global__ void __launch_bounds__(256) k_linefinish(const VPar *__restrict__ vps, Scratch S, int n_ener, int ne_stride,
                                                    int nz_stride, double *__restrict__ out) {
  __shared__ double red[256];
  const int v = blockIdx.x, t = threadIdx.x;
  double *o = out + (size_t) v * n_ener;
  if (S.status[v] != ST_OK) {
    for (int j = t; j < n_ener; j += 256) o[j] = 0.0;
    return;
  }
  const VPar &vp = vps[v];
  const double *flux = S.relflux + (size_t) v * nz_stride * ne_stride;
  const int jlo = S.zrange[(size_t) v * NZMAX * 2], jhi = S.zrange[(size_t) v * NZMAX * 2 + 1];
  double part = 0.0;
  for (int j = t; j < n_ener; j += 256) part += (j >= jlo && j <= jhi) ? flux[j] : 0.0;
  const double sum = block_sum<256>(part, red);
  const double scale = vp.renorm ? vp.relline_norm / sum : 1.0;
  for (int j = t; j < n_ener; j += 256) {
    const double f = (j >= jlo && j <= jhi) ? flux[j] : 0.0;
    o[j] = vp.renorm ? f * scale : f;
  }
}

// ---------------------------------------------------------------------------------- k_xillver
// Standalone xillver / xillverCp (LocalModel::xillver_model, src/LocalModel.cpp:104-130): interpolation over
// all table axes including the inclination (interp_5d_tab / interp_6d_tab, src/xilltable.c:878-1044),
// semi-infinite-slab factor (norm_xillver_spec, src/Xillspec.cpp:528-545), rebin to the caller's grid and
// add_primary_component (src/Relbase.cpp:294-351; the nthcomp primary is added by k_xillver_prim_nth).
__global__ void __launch_bounds__(256) k_xillver(const VPar *__restrict__ vps, DevTables T, Scratch S, int which,
                                                 const double *__restrict__ user_e, int n_flux, double *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double *fx = reinterpret_cast<double *>(smraw);   // [stride]
  __shared__ double s_w[64], s_pow[NCOARSE], s_nsrc, s_fac0;
  __shared__ const float *s_row[64];
  const int v = blockIdx.x, t = threadIdx.x;
  const VPar &vp = vps[v];
  double *o = out + (size_t) v * n_flux;
  if (t == 0) S.status[v] = vp.status;
  if (vp.status != ST_OK) {
    for (int j = t; j < n_flux; j += 256) o[j] = 0.0;
    return;
  }
  const XillDev &X = T.xill[which];
  const int ne = X.n_ener, st = X.stride;
  const int ncorn = (X.npar == 6) ? 64 : 32;
  if (t == 0) {
    float inp[8];
    inp[0] = (float) vp.gam; inp[1] = (float) vp.afe; inp[2] = (float) vp.lxi; inp[3] = (float) vp.ect;
    inp[4] = (float) vp.dens; inp[5] = (float) vp.ktbb; inp[6] = (float) vp.frac_pl_bb; inp[7] = (float) vp.xincl;
    int ind[6];
    double fac[6];
    for (int i = 0; i < X.npar; i++) {
      const int pind = X.pindex[i];
      const int n = X.nvals[i];
      int k = bsearch_asc<float>(X.vals[i], n, inp[pind]);
      if (k < 0) k = 0; else if (k > n - 2) k = n - 2;
      ind[i] = k;
      float val = inp[pind];
      const float lo = X.vals[i][0], hi = X.vals[i][n - 1];
      if (val < lo) val = lo; else if (val > hi) val = hi;
      fac[i] = (double) ((val - X.vals[i][k]) / (X.vals[i][k + 1] - X.vals[i][k]));
      if (pind == 3) {   // limits from table axis number 3, like the reference (see k_zone)
        if (vp.ect <= (double) X.vals[3][0]) fac[i] = 0.0;
        if (vp.ect >= (double) X.vals[3][X.nvals[3] - 1]) fac[i] = 1.0;
      }
    }
    const int off = (X.npar == 6) ? 1 : 0;
    const int bits[16] = {0x0, 0x1, 0x2, 0x4, 0x3, 0x5, 0x6, 0x7, 0x8, 0x9, 0xA, 0xC, 0xB, 0xD, 0xE, 0xF};
    for (int half = 0; half < (X.npar == 6 ? 2 : 1); half++)
      for (int h5 = 0; h5 < 2; h5++)
        for (int c = 0; c < 16; c++) {
          const int b1 = bits[c] & 1, b2 = (bits[c] >> 1) & 1, b3 = (bits[c] >> 2) & 1, b4 = (bits[c] >> 3) & 1;
          const double w = (b1 ? fac[off] : (1.0 - fac[off])) * (b2 ? fac[off + 1] : (1.0 - fac[off + 1]))
                           * (b3 ? fac[off + 2] : (1.0 - fac[off + 2])) * (b4 ? fac[off + 3] : (1 - fac[off + 3]))
                           * (h5 ? fac[off + 4] : (1 - fac[off + 4]));
          const int i1 = ind[off] + b1, i2 = ind[off + 1] + b2, i3 = ind[off + 2] + b3, i4 = ind[off + 3] + b4,
                    i5 = ind[off + 4] + h5;
          long row;
          if (X.npar == 5) row = ((((long) i1 * X.nvals[1] + i2) * X.nvals[2] + i3) * X.nvals[3] + i4) * X.nvals[4] + i5;
          else row = (((((long) (ind[0] + half) * X.nvals[1] + i1) * X.nvals[2] + i2) * X.nvals[3] + i3) * X.nvals[4] + i4) * X.nvals[5] + i5;
          s_w[half * 32 + h5 * 16 + c] = w;
          s_row[half * 32 + h5 * 16 + c] = X.data + (size_t) row * st;
        }
    s_fac0 = fac[0];
  }
  if (vp.prim_type == PRIM_ECUT)
    for (int i = t; i < NCOARSE; i += 256) s_pow[i] = pow(0.5 * (T.ecoarse[i] + T.ecoarse[i + 1]), -vp.gam);
  __syncthreads();
  if ((vp.prim_type == PRIM_ECUT || vp.prim_type == PRIM_BB) && t < 32) {
    double S1, S2;
    if (vp.prim_type == PRIM_BB) bb_band_sums(T, vp.ktbb, S1, S2); else ecut_band_sums(T, s_pow, vp.ect, S1, S2);
    if (t == 0) { s_nsrc = 1. / (S1 / (1e15 / 4.0 / PI)); S.nsrc[v] = s_nsrc; }
  }
  const double slab = 0.5 * cos(vp.xincl * PI / 180);
  for (int e = t; e < ne; e += 256) {
    double f1 = s_w[0] * (double) __ldg(s_row[0] + e);
    for (int c = 1; c < 32; c++) f1 += s_w[c] * (double) __ldg(s_row[c] + e);
    if (ncorn == 64) {
      double f2 = s_w[32] * (double) __ldg(s_row[32] + e);
      for (int c = 33; c < 64; c++) f2 += s_w[c] * (double) __ldg(s_row[c] + e);
      f1 = lin1d(s_fac0, f1, f2);
    }
    fx[e] = f1 * slab;
  }
  __syncthreads();
  const double rfa = fabs(vp.refl_frac);
  const bool add_prim = (vp.refl_frac >= 0) && (vp.prim_type == PRIM_ECUT || vp.prim_type == PRIM_BB);
  const bool bb = (vp.prim_type == PRIM_BB);
  const double ex0 = exp(1.0 / vp.ect), kt4 = pow(vp.ktbb, 4);
  for (int j = t; j < n_flux; j += 256) {
    double elo = user_e[j], ehi = user_e[j + 1];
    if (vp.z > 0) { elo *= (1 + vp.z); ehi *= (1 + vp.z); }
    double f = rebin_bin(elo, ehi, X.ener, fx, ne) * rfa;
    if (add_prim) {
      const double en = 0.5 * (elo + ehi);
      double pr;
      if (bb) {
        pr = en * en / (kt4 * (exp(en / vp.ktbb) - 1));
        pr *= (ehi - elo);
      } else {
        pr = ex0 * pow(en, -vp.gam) * exp(-en / vp.ect) * (ehi - elo);
      }
      pr *= s_nsrc;
      f += pr;
    }
    o[j] = f;
  }
}

// ---------------------------------------------------------------------------------- launchers
static size_t g_smem_sys = 0, g_smem_zone = 0;

int line_kernel_init();
int nth_kernel_init();
int xill_kernel_init();
int conv_kernel_init();

int kernels_init() {
  g_smem_sys = sizeof(SysSmem);
  g_smem_zone = sizeof(ZoneSmem);
  cudaError_t e;
  e = cudaFuncSetAttribute(k_syspar, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) g_smem_sys);
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(k_zone, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) g_smem_zone);
  if (e != cudaSuccess) return 1;
  if (line_kernel_init() != 0) return 1;
  if (nth_kernel_init() != 0) return 1;
  if (xill_kernel_init() != 0) return 1;
  if (conv_kernel_init() != 0) return 1;
  return 0;
}

void launch_syspar(const VPar *vps, const DevTables &T, const Scratch &S, long n, int pass, cudaStream_t st) {
  k_syspar<<<(unsigned) n, 256, g_smem_sys, st>>>(vps, T, S, pass);
}
void launch_zone(const VPar *vps, const DevTables &T, const Scratch &S, long n, cudaStream_t st) {
  k_zone<<<(unsigned) n, 128, g_smem_zone, st>>>(vps, T, S);
}
void launch_fine(const VPar *vps, const DevTables &T, const Scratch &S, long n, int n_incl, double e_first,
                 double e_last, int store_cosne, int store_trff, cudaStream_t st) {
  dim3 grid_rows((REL_NRT + 7) / 8, (unsigned) n);
  k_rows<<<grid_rows, 320, 0, st>>>(vps, T, S);
  dim3 grid(NR / 8, (unsigned) n);
  k_fine<<<grid, 320, 0, st>>>(vps, T, S, n_incl, e_first, e_last, store_cosne, store_trff);
}
void launch_dist(const VPar *vps, const DevTables &T, const Scratch &S, long n, int n_incl, cudaStream_t st) {
  k_dist<<<(unsigned) n, 256, 0, st>>>(vps, T, S, n_incl);
}
void launch_linefinish(const VPar *vps, const Scratch &S, long n, int n_ener, double *out, cudaStream_t st) {
  k_linefinish<<<(unsigned) n, 256, 0, st>>>(vps, S, n_ener, S.ne_line_cap, S.nz_cap, out);
}
void launch_xillver(const VPar *vps, const DevTables &T, const Scratch &S, long n, int which, const double *user_e,
                    int n_flux, double *out, int stride, cudaStream_t st) {
  k_xillver<<<(unsigned) n, 256, (size_t) stride * sizeof(double), st>>>(vps, T, S, which, user_e, n_flux, out);
}
}  // namespace rx
