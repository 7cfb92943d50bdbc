// nthcomp.cu — thermal Comptonisation continuum (nthcomp) for the Cp models on the device.
//
// Replaces c_donthcomp and its callees for the one configuration relxill uses (disk-blackbody seed with
// kT_bb = 0.05 keV, reference src/relutility.c:625-632; src/donthcomp.c:200-301 f_thermlc__, :467-648
// f_thdscompton__, :651-793 f_spp__ + c_donthcomp) together with the primary-spectrum normalisations that
// k_zone computes for the cutoff power law (src/Xillspec.cpp:179-205,408-440, src/PrimarySource.h:279-293)
// and the returning-radiation flux correction (src/Xillspec.cpp:344-362,500-526).
//
// A relxilllpCp evaluation needs one Kompaneets solution per radial zone (kTe shifted into the zone's frame)
// plus one for the source; the reference solves 3*Nz+3 times because every consumer calls c_donthcomp again
// (SURVEY.md §3.5).  Here each distinct solution is computed once: one thread per solve runs the
// tridiagonal recurrence (inherently sequential in the photon-energy index), the per-solve work arrays are
// laid out [energy index][solve] so the threads of a block read and write them coalesced, and everything
// that depends only on the photon grid was tabulated at load (tables.cu, load_nthcomp).
#include <cuda_runtime.h>

#include "common.h"
#include "devutil.cuh"
#include "kernels.h"

namespace rx {

constexpr double LOG10E = 0.43429448190325182765;

// f_thdscompton__ + f_thermlc__ for one (theta, gamma); arrays strided by NTH_SOL. returns jmax.
__device__ int nth_solve(const DevTables &T, double theta, double gamma, double *gam, double *g, double *spt) {
  const double *x = T.nth_x, *w = T.nth_w, *c2 = T.nth_c2, *rel = T.nth_rel, *x3 = T.nth_x3, *dph = T.nth_dphdot;
  const double d1 = gamma + .5;
  const double tautom = sqrt(3. / (theta * (d1 * d1 - 2.25)) + 2.25) - 1.5;
  const double deltal = T.nth_deltal;
  const double xmax = theta * 40.;
  int jmax = (int) (LOG10E * log(xmax / T.nth_xmin) / .02) + 1;
  if (jmax > 899) jmax = 899;
  if (jmax < 4) jmax = 4;
  int jnr = T.nth_jnr, jrel = T.nth_jrel;
  if (jnr > jmax - 1) jnr = jmax - 1;
  if (jrel > jmax) jrel = jmax;
  const double xnr = x[jnr - 1], xr = x[jrel - 1];
  auto bet = [&](int j) -> double {   // j 1-based
    if (j > jrel) return 1 / tautom;
    const double taukn = tautom * rel[j - 1];
    if (j <= jnr - 1) return 1 / tautom / (taukn / 3 + 1);
    const double arg = (x[j - 1] - xnr) / (xr - xnr);
    const double flz = 1 - arg;
    return 1 / tautom / (taukn / 3 * flz + 1);
  };
  const double c20 = tautom / deltal;
  const double td = theta / deltal;
  const double x32 = w[0];
  const double aa = (td / x32 + .5) / (td / x32 - .5);
  // forward elimination: gam[j-1], g[j-1] for j = 2 .. jmax-1 (1-based j as in the reference)
  double gam_prev = 0.0, g_prev = 0.0;
  for (int j = 2; j <= jmax - 1; j++) {
    const double w1 = w[j - 1], w2 = w[j - 2];
    const double a = -c20 * c2[j - 1] * (td / w1 + .5);
    const double t1 = -c20 * c2[j - 1] * (.5 - td / w1);
    const double t2 = c20 * c2[j - 2] * (td / w2 + .5);
    const double t3 = x3[j - 1] * (tautom * bet(j));
    const double b = t1 + t2 + t3;
    const double c = c20 * c2[j - 2] * (.5 - td / w2);
    const double d = x[j - 1] * dph[j - 1];
    double alp, gg;
    if (j == 2) {
      alp = b + c * aa;
      gg = d / alp;
    } else {
      alp = b - c * gam_prev;
      // the last row also carries the (zero) boundary value u[jmax]: (d - a*0 - c*g)/alp
      gg = (j == jmax - 1) ? (d - a * 0. - c * g_prev) / alp : (d - c * g_prev) / alp;
    }
    gam_prev = a / alp;
    g_prev = gg;
    gam[(size_t) (j - 1) * NTH_SOL] = gam_prev;
    g[(size_t) (j - 1) * NTH_SOL] = gg;
  }
  // back substitution + escaping photon density -> E F_E
  for (int j = 0; j < NTH_MAX; j++) { if (j >= jmax - 1) spt[(size_t) j * NTH_SOL] = 0.0; }
  double u_next = g_prev;   // u[jmax-1] (1-based) = g[jmax-2]
  {
    const int j = jmax - 1;
    const double dphesc = x[j - 1] * x[j - 1] * u_next * bet(j) * tautom;
    spt[(size_t) (j - 1) * NTH_SOL] = dphesc * (x[j - 1] * x[j - 1]);
  }
  double u2 = u_next;   // will end as u of 1-based index 2
  for (int jj = jmax - 2; jj >= 2; jj--) {
    const double u = g[(size_t) (jj - 1) * NTH_SOL] - gam[(size_t) (jj - 1) * NTH_SOL] * u_next;
    const double dphesc = x[jj - 1] * x[jj - 1] * u * bet(jj) * tautom;
    spt[(size_t) (jj - 1) * NTH_SOL] = dphesc * (x[jj - 1] * x[jj - 1]);
    u_next = u;
    u2 = u;
  }
  {
    const double u = aa * u2;
    const double dphesc = x[0] * x[0] * u * bet(1) * tautom;
    spt[0] = dphesc * (x[0] * x[0]);
  }
  return jmax;
}

// value of the E F_E solution at photon energy e_kev (in the source frame after the redshift factor zfac =
// 1 + z): the interpolation of c_donthcomp :759-780.  nth = jmax.
__device__ __forceinline__ double nth_prim(const DevTables &T, const double *spt, int nth, double e_kev, double zfac) {
  const double *xth = T.nth_x;
  const double target = e_kev * zfac;
  // j = first 1-based index with NOT (xth[j-1]*511 < target)
  int lo = 0, hi = nth;
  while (lo < hi) {
    const int m = (lo + hi) >> 1;
    if (xth[m] * 511. < target) lo = m + 1; else hi = m;
  }
  const int j = lo + 1;
  if (j > nth) return 0.0;
  if (j > 1) {
    const int jl = j - 1;
    const double s0 = spt[(size_t) (jl - 1) * NTH_SOL], s1 = spt[(size_t) jl * NTH_SOL];
    return s0 + (e_kev / 511. * zfac - xth[jl - 1]) * (s1 - s0) / (xth[jl] - xth[jl - 1]);
  }
  return spt[0];
}

// 1 / f_spp__(1/xn): normalisation at 1 keV (observer frame), c_donthcomp :737-741 + f_spp__ :651-681
__device__ double nth_normfac(const DevTables &T, const double *spt, int nth, double zfac) {
  const double *xth = T.nth_x;
  const double xn = zfac / 511.;
  const double xx = 1 / (1 / xn);
  int ih = 2;
  while (ih < nth && xx > xth[ih - 1]) ++ih;
  const int il = ih - 1;
  const double s0 = spt[(size_t) (il - 1) * NTH_SOL], s1 = spt[(size_t) (ih - 1) * NTH_SOL];
  return 1 / (s0 + (s1 - s0) * (xx - xth[il - 1]) / (xth[ih - 1] - xth[il - 1]));
}

// photons per bin [e0, e1] (c_donthcomp :782-786)
__device__ __forceinline__ double nth_bin(const DevTables &T, const double *spt, int nth, double e0, double e1, double zfac,
                                          double normfac) {
  const double p0 = nth_prim(T, spt, nth, e0, zfac), p1 = nth_prim(T, spt, nth, e1, zfac);
  return (p1 / (e1 * e1) + p0 / (e0 * e0)) * .5 * (e1 - e0) * normfac;
}

// ---------------------------------------------------------------------------------- k_nth
// One CTA per vector of a Cp model, after k_zone: the Kompaneets solutions of the zones and the source,
// their xillver-normalisation integrals on the coarse grid, the normalisation-change factors and the
// returning-radiation flux-correction factors.
__global__ void __launch_bounds__(128) k_nth(const VPar *__restrict__ vps, DevTables T, Scratch S) {
  __shared__ double s_nfac[NZMAX + 1], s_s2[NZMAX + 1];
  __shared__ int s_jmax[NZMAX + 1];
  const int v = blockIdx.x, t = threadIdx.x;
  if (S.status[v] != ST_OK) return;
  if (S.reuse && (S.reuse[v] & REUSE_ALL)) return;
  const VPar &vp = vps[v];
  if (vp.prim_type != PRIM_NTHCOMP) return;
  const int nz = vp.nz;
  double *gam = S.nth_gam + (size_t) v * NTH_MAX * NTH_SOL, *g = S.nth_g + (size_t) v * NTH_MAX * NTH_SOL;
  double *spt = S.nth_spt + (size_t) v * NTH_MAX * NTH_SOL;
  if (t <= nz) {
    const double kte = (t < nz) ? S.zect[(size_t) v * NZMAX + t] : vp.ect;   // zone: kTe * energy shift; source: kTe
    const int jm = nth_solve(T, kte / 511., vp.gam, gam + t, g + t, spt + t);
    s_jmax[t] = jm;
    S.nth_jmax[(size_t) v * NTH_SOL + t] = jm;
  }
  __syncthreads();
  // band integrals on the coarse grid (z = 0: ener_shift = 1), one warp per solve
  const int warp = t >> 5, lane = t & 31;
  for (int job = warp; job <= nz; job += 4) {
    const double *sp = spt + job;
    const int nth = s_jmax[job];
    const double normfac = nth_normfac(T, sp, nth, 1.0);
    double a1 = 0.0, a2 = 0.0;
    for (int i = lane; i < NCOARSE; i += 32) {
      const double e0 = T.ecoarse[i], e1 = T.ecoarse[i + 1];
      const double fl = nth_bin(T, sp, nth, e0, e1, 1.0, normfac);
      const double wgt = fl * 0.5 * (e0 + e1);
      if (T.coarse_m1[i]) a1 += wgt * 1e20 * 1.602177e-09;
      if (T.coarse_m2[i]) a2 += wgt;
    }
    for (int o = 16; o > 0; o >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
      s_nfac[job] = 1. / (a1 / (1e15 / 4.0 / PI));
      s_s2[job] = a2;
    }
  }
  __syncthreads();
  if (t == 0) S.nsrc[v] = s_nfac[nz];
  if (t < nz) {
    S.normch[(size_t) v * NZMAX + t] = s_nfac[t] / s_nfac[nz];
    if (vp.do_corr) {
      const XillDev &X = T.xill[1];
      const int *xr = S.xrow + ((size_t) v * NZMAX + t) * 32;
      const double *xw = S.xw + ((size_t) v * NZMAX + t) * 32;
      double ef = 0.0, p1 = 0.0, p2 = 0.0;
      const int nc = (X.npar == 6) ? 32 : 16;
      for (int c = 0; c < nc; c++) {
        ef += xw[c] * X.node_ef[xr[c]];
        p1 += xw[c] * X.node_p1[xr[c]];
        p2 += xw[c] * X.node_p2[xr[c]];
      }
      const double direct = s_s2[t] * s_nfac[t];
      S.corr_flux[(size_t) v * NZMAX + t] = ef / direct;
      S.corr_gshift[(size_t) v * NZMAX + t] = (p1 / p2) / pow(1.5, vp.gam);
    }
  }
}

// primary spectrum of a Cp model on the convolution grid, added to the convolved reflection
// (PrimarySource::add_primary_spectrum, src/PrimarySource.cpp:66-125 with spec_nthcomp, src/Xillspec.cpp:243-253).
// k_conv has already applied the reflection scaling and left the 4096-bin result in `total`; this kernel adds
// the primary and rebins to the caller's grid.
__global__ void __launch_bounds__(256) k_prim_nth(const VPar *__restrict__ vps, DevTables T, Scratch S, double *total,
                                                  const double *__restrict__ user_e, int n_flux, double *out, int renorm3) {
  __shared__ double acc[NCONV];
  const int v = blockIdx.x, t = threadIdx.x;
  if (S.status[v] != ST_OK) return;
  const VPar &vp = vps[v];
  const bool reuse_all = S.reuse && (S.reuse[v] & REUSE_ALL);   // `total` already holds reflection + primary
  const int nz = vp.nz;
  const double *sp = S.nth_spt + (size_t) v * NTH_MAX * NTH_SOL + nz;
  const int nth = S.nth_jmax[(size_t) v * NTH_SOL + nz];
  const double zfac = 1 / vp.eshift_obs - 1 + 1;   // z + 1 with z = 1/shift - 1 (src/Xillspec.cpp:250)
  const double normfac = nth_normfac(T, sp, nth, zfac);
  double prim_scale = 1.0;
  if (vp.emis_type == EMIS_LP) {
    const double *rf = S.reflfrac + (size_t) v * 8;
    prim_scale = rf[4] / 0.5 * pow(vp.eshift_obs, vp.gam);
    if (vp.beta > 1e-4) prim_scale *= vp.doppler_obs * vp.doppler_obs;
  }
  const double nsrc = S.nsrc[v];
  const bool add_prim = (vp.refl_frac >= 0) && !reuse_all;
  double *tot = total + (size_t) v * NCONV;
  for (int i = t; i < NCONV; i += 256) {
    double val = tot[i];
    if (add_prim) {
      double pr = nth_bin(T, sp, nth, T.econv[i], T.econv[i + 1], zfac, normfac);
      pr *= nsrc;
      if (vp.emis_type == EMIS_LP) pr *= prim_scale;
      val += pr;
    }
    acc[i] = val;
    tot[i] = val;
  }
  __syncthreads();
  double rn = 1.0;   // RELXILL_RENORMALIZE: renorm_relxill_spectrum_1keV, src/Relxill.cpp:249-259
  if (renorm3) {
    const int i3 = T.conv_i3kev;
    rn = 1.0 / (acc[i3] / (T.econv[i3 + 1] - T.econv[i3]));
  }
  double *o = out + (size_t) v * n_flux;
  for (int j = t; j < n_flux; j += 256) {   // _rebin_spectrum (src/relutility.c:549-601), one output bin per thread
    double elo_o = user_e[j], ehi_o = user_e[j + 1];
    if (vp.z > 0) { elo_o *= (1 + vp.z); ehi_o *= (1 + vp.z); }
    const double *e0 = T.econv;
    double f = 0.0;
    if ((e0[0] <= ehi_o) && (e0[NCONV] >= elo_o)) {
      int imin = count_le_asc(e0, NCONV + 1, elo_o) - 1;
      if (imin < 0) imin = 0;
      int imax = count_le_asc(e0, NCONV + 1, ehi_o);
      if (imax > NCONV) imax = NCONV;
      imax -= 1;
      if (imax < 0) imax = 0;
      double elo = elo_o, ehi = ehi_o;
      if (elo < e0[imin]) elo = e0[imin];
      if (ehi > e0[imax + 1]) ehi = e0[imax + 1];
      if (imax == imin) f = (ehi - elo) / (e0[imin + 1] - e0[imin]) * acc[imin];
      else {
        const double dmin = (e0[imin + 1] - elo) / (e0[imin + 1] - e0[imin]);
        const double dmax = (ehi - e0[imax]) / (e0[imax + 1] - e0[imax]);
        f += acc[imin] * dmin + acc[imax] * dmax;
        for (int jj = imin + 1; jj <= imax - 1; jj++) f += acc[jj];
      }
    }
    o[j] = renorm3 ? f * rn : f;
  }
}

// nthcomp primary of the standalone xillverCp model on the caller's (redshifted) grid, added to k_xillver's
// reflection spectrum (add_primary_component, src/Relbase.cpp:294-351 with energy shift 1)
__global__ void __launch_bounds__(256) k_xillver_prim_nth(const VPar *__restrict__ vps, DevTables T, Scratch S,
                                                          const double *__restrict__ user_e, int n_flux, double *out) {
  const int v = blockIdx.x, t = threadIdx.x;
  if (S.status[v] != ST_OK) return;
  const VPar &vp = vps[v];
  if (!(vp.refl_frac >= 0)) return;
  const double *sp = S.nth_spt + (size_t) v * NTH_MAX * NTH_SOL;
  const int nth = S.nth_jmax[(size_t) v * NTH_SOL];
  const double normfac = nth_normfac(T, sp, nth, 1.0);
  const double nsrc = S.nsrc[v];
  double *o = out + (size_t) v * n_flux;
  for (int j = t; j < n_flux; j += 256) {
    double elo = user_e[j], ehi = user_e[j + 1];
    if (vp.z > 0) { elo *= (1 + vp.z); ehi *= (1 + vp.z); }
    o[j] += nth_bin(T, sp, nth, elo, ehi, 1.0, normfac) * nsrc;
  }
}

void launch_xillver_prim_nth(const VPar *vps, const DevTables &T, const Scratch &S, long n, const double *user_e, int n_flux,
                             double *out, cudaStream_t st) {
  k_xillver_prim_nth<<<(unsigned) n, 256, 0, st>>>(vps, T, S, user_e, n_flux, out);
}

void launch_nth(const VPar *vps, const DevTables &T, const Scratch &S, long n, cudaStream_t st) {
  k_nth<<<(unsigned) n, 128, 0, st>>>(vps, T, S);
}
void launch_prim_nth(const VPar *vps, const DevTables &T, const Scratch &S, long n, double *total, const double *user_e,
                     int n_flux, double *out, int renorm3, cudaStream_t st) {
  k_prim_nth<<<(unsigned) n, 256, 0, st>>>(vps, T, S, total, user_e, n_flux, out, renorm3);
}

}  // namespace rx
