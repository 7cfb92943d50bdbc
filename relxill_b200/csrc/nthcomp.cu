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
// (SURVEY.md §3.5).  Here each distinct solution is computed once, one thread per solve (the tridiagonal recurrence is
// sequential in the photon-energy index), 64 solves of consecutive vectors per CTA whatever the zone count is.
// The elimination coefficients never leave the SM: the forward sweep keeps a checkpoint every 32 steps in shared
// memory and the back substitution recomputes one 32-step segment at a time from its checkpoint (twice the flops of
// the sweep instead of 1.4 MB of HBM scratch per vector written and read back: the kernel was bound by that traffic).
// The zone solutions are not filed either: what is needed of them, two band integrals on the coarse grid and the
// value at 1 keV, is linear in the solution and is accumulated while it is produced (weights tabulated at load,
// tables.cu); only the source's solution goes to HBM for the primary spectrum.
#include <cuda_runtime.h>

#include "common.h"
#include "devutil.cuh"
#include "kernels.h"

namespace rx {

constexpr double LOG10E = 0.43429448190325182765;
constexpr int NTH_NT = 64;     // solves per CTA
constexpr int NTH_SEG = 32;    // steps between checkpoints
constexpr int NTH_NSEG = (NTH_MAX + NTH_SEG - 1) / NTH_SEG;

// Arithmetic.  The reference's rows cost six IEEE divisions each (td/w twice, taukn/3, 1/tautom/(...), a/alp, g/alp);
// with 51 solves x 2 x 900 rows per vector that made the kernel instruction-bound.  Here the node-only quotients are
// tabulated (tables.cu), td/w and the neighbouring row's coefficients are carried over instead of recomputed
// (t2_j = -a_{j-1}, c_j = -t1_{j-1} exactly), and the two remaining quotients per row are reciprocals refined by two
// Newton steps (~1 ulp).  The solution differs from the reference's by rounding only (measured against the golden
// vectors: see tests), the recurrence and its order are the reference's.
__device__ __forceinline__ double nth_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = r * (2.0 - x * r);
  r = r * (2.0 - x * r);
  return r;
}
struct NthSolve {   // per-thread constants of one (theta, gamma)
  double tautom, c20, td, aa, xnr, rxr;   // rxr = 1 / (xr - xnr)
  int jmax, jnr, jrel;
};
// 1 / (tautom bet(j)) = taukn/3 * flz + 1, the scattering term's denominator (f_thermlc__: bet)
__device__ __forceinline__ double nth_den(const DevTables &T, const NthSolve &q, int j) {   // j 1-based
  if (j > q.jrel) return 1.0;
  const double tk3 = q.tautom * T.nth_rel[j - 1] * (1.0 / 3.0);
  if (j <= q.jnr - 1) return tk3 + 1;
  const double flz = 1 - (T.nth_x[j - 1] - q.xnr) * q.rxr;
  return tk3 * flz + 1;
}
// carried from row to row: (gam, g) of the recurrence, td/w and the two coefficients the next row shares
struct NthRow { double gam, g, a, t1; };
// one row of the forward elimination (f_thermlc__, src/donthcomp.c:200-301), j = 2 .. jmax-1 (1-based as in the reference)
__device__ __forceinline__ void nth_step(const DevTables &T, const NthSolve &q, int j, NthRow &r) {
  const double p = q.c20 * T.nth_c2[j - 1], qw = q.td * T.nth_rw[j - 1];
  const double a = -p * (qw + .5);
  const double t1 = -p * (.5 - qw);
  const double t2 = -r.a;       // c20 c2[j-2] (td/w2 + .5)
  const double c = -r.t1;       // c20 c2[j-2] (.5 - td/w2)
  const double t3 = T.nth_x3[j - 1] * nth_rcp(nth_den(T, q, j));
  const double b = t1 + t2 + t3;
  const double d = T.nth_xd[j - 1];
  // row 2 closes on the boundary condition u[1] = aa u[2]; the last row also carries the (zero) boundary value u[jmax]
  const double alp = (j == 2) ? b + c * q.aa : b - c * r.gam;
  const double ra = nth_rcp(alp);
  r.g = (j == 2) ? d * ra : (d - c * r.g) * ra;
  r.gam = a * ra;
  r.a = a;
  r.t1 = t1;
}
// the coefficients a, t1 of row j (what row j + 1 takes over from it)
__device__ __forceinline__ void nth_coef(const DevTables &T, const NthSolve &q, int j, NthRow &r) {
  const double p = q.c20 * T.nth_c2[j - 1], qw = q.td * T.nth_rw[j - 1];
  r.a = -p * (qw + .5);
  r.t1 = -p * (.5 - qw);
}

// (the solution is contiguous: only the source's is filed)
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
    const double s0 = spt[jl - 1], s1 = spt[jl];
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
  const double s0 = spt[il - 1], s1 = spt[ih - 1];
  return 1 / (s0 + (s1 - s0) * (xx - xth[il - 1]) / (xth[ih - 1] - xth[il - 1]));
}

// photons per bin [e0, e1] (c_donthcomp :782-786)
__device__ __forceinline__ double nth_bin(const DevTables &T, const double *spt, int nth, double e0, double e1, double zfac,
                                          double normfac) {
  const double p0 = nth_prim(T, spt, nth, e0, zfac), p1 = nth_prim(T, spt, nth, e1, zfac);
  return (p1 / (e1 * e1) + p0 / (e0 * e0)) * .5 * (e1 - e0) * normfac;
}

// ---------------------------------------------------------------------------------- k_nth
// Thread = one Kompaneets solve: solve k of vector v (k < nz: zone k, kTe shifted into the zone's frame; k = nz: the
// source), 64 consecutive (v, k) pairs per CTA.  f_thdscompton__ + f_thermlc__ (src/donthcomp.c:200-301,467-648).
struct NthSmem {
  double2 ck[NTH_NSEG][NTH_NT];    // (gam, g) entering each segment
  double2 seg[NTH_SEG][NTH_NT];    // (gam, g) of the segment being back-substituted
};
__global__ void __launch_bounds__(NTH_NT) k_nth(const VPar *__restrict__ vps, DevTables T, Scratch S, long n, int per_vec) {
  extern __shared__ __align__(16) unsigned char smraw[];
  NthSmem &sm = *reinterpret_cast<NthSmem *>(smraw);
  const int t = threadIdx.x;
  const long gs = (long) blockIdx.x * NTH_NT + t;
  const long v = gs / per_vec;
  const int k = (int) (gs - v * per_vec);
  bool on = v < n;
  if (on) on = (S.status[v] == ST_OK) && !(S.reuse && (S.reuse[v] & REUSE_ALL)) && vps[v].prim_type == PRIM_NTHCOMP && k <= vps[v].nz;
  NthSolve q;
  q.jmax = 0;
  bool source = false;
  if (on) {
    const VPar &vp = vps[v];
    source = (k == vp.nz);
    const double kte = source ? vp.ect : S.zect[(size_t) v * NZMAX + k];   // zone: kTe * energy shift; source: kTe
    const double theta = kte / 511., gamma = vp.gam;
    const double d1 = gamma + .5;
    q.tautom = sqrt(3. / (theta * (d1 * d1 - 2.25)) + 2.25) - 1.5;
    const double xmax = theta * 40.;
    int jmax = (int) (LOG10E * log(xmax / T.nth_xmin) / .02) + 1;
    if (jmax > 899) jmax = 899;
    if (jmax < 4) jmax = 4;
    q.jmax = jmax;
    q.jnr = min(T.nth_jnr, jmax - 1);
    q.jrel = min(T.nth_jrel, jmax);
    q.xnr = T.nth_x[q.jnr - 1];
    q.rxr = 1.0 / (T.nth_x[q.jrel - 1] - q.xnr);
    q.c20 = q.tautom / T.nth_deltal;
    q.td = theta / T.nth_deltal;
    const double x32 = T.nth_w[0];
    q.aa = (q.td / x32 + .5) / (q.td / x32 - .5);
  }
  const int jmax = q.jmax;
  // ---- forward elimination, checkpoints only: segment s covers the rows j = 2 + 32 s .. 33 + 32 s
  NthRow row;
  nth_coef(T, q, 1, row);
  row.gam = 0.0; row.g = 0.0;
  int jtop = 0;
  for (int o = 16, m = jmax; o > 0; o >>= 1) { m = max(m, __shfl_xor_sync(0xffffffffu, m, o)); jtop = m; }
  if (jtop == 0) return;   // (a whole warp without work)
  for (int j = 2; j <= jtop - 1; j++) {
    if (((j - 2) & (NTH_SEG - 1)) == 0) sm.ck[(j - 2) / NTH_SEG][t] = make_double2(row.gam, row.g);
    if (j <= jmax - 1) nth_step(T, q, j, row);
  }
  // ---- back substitution + escaping photon density -> E F_E, segment by segment from the top
  const double *x = T.nth_x;
  double *spt = S.nth_spt + (size_t) (on ? v : 0) * NTH_MAX;
  if (on && source)
    for (int j = jmax - 1; j < NTH_MAX; j++) spt[j] = 0.0;
  const int ih1 = min(T.nth_ih1, jmax), il1 = ih1 - 1;   // f_spp__ bracket (1-based)
  double a1 = 0.0, a2 = 0.0, s_il = 0.0, s_ih = 0.0;
  auto emit = [&](int j, double u) {   // row j (1-based): E F_E at x_j
    const double val = T.nth_x4[j - 1] * u * nth_rcp(nth_den(T, q, j));   // x^2 u bet tautom x^2
    if (source) spt[j - 1] = val;
    a1 += T.nth_w1[j - 1] * val;
    a2 += T.nth_w2[j - 1] * val;
    if (j == il1) s_il = val;
    if (j == ih1) s_ih = val;
  };
  double u_next = row.g;    // u[jmax-1] (1-based) = g of the last row
  if (on) emit(jmax - 1, u_next);
  double u2 = u_next;       // will end as u of 1-based index 2
  for (int s = (jtop - 3) / NTH_SEG; s >= 0; s--) {
    const int j0 = 2 + s * NTH_SEG;
    // the segment's rows again, from its checkpoint
    const double2 c2 = sm.ck[s][t];
    NthRow c;
    nth_coef(T, q, j0 - 1, c);   // the coefficients the segment's first row shares with the row before it
    c.gam = c2.x; c.g = c2.y;
    for (int i = 0; i < NTH_SEG; i++) {
      const int j = j0 + i;
      if (j <= jmax - 1) nth_step(T, q, j, c);
      sm.seg[i][t] = make_double2(c.gam, c.g);
    }
    for (int i = NTH_SEG - 1; i >= 0; i--) {
      const int jj = j0 + i;
      if (on && jj <= jmax - 2) {   // rows jmax-2 .. 2
        const double2 gg = sm.seg[i][t];
        const double u = gg.y - gg.x * u_next;
        emit(jj, u);
        u_next = u;
        u2 = u;
      }
    }
  }
  if (!on) return;
  emit(1, q.aa * u2);
  // normalisation at 1 keV (f_spp__, src/donthcomp.c:651-681; z = 0) and the two band integrals
  const double normfac = 1 / (s_il + (s_ih - s_il) * (T.nth_xx1 - x[il1 - 1]) / (x[ih1 - 1] - x[il1 - 1]));
  S.nth_jmax[(size_t) v * NTH_SOL + k] = jmax;
  S.nth_nfac[(size_t) v * NTH_SOL + k] = 1. / ((a1 * normfac) / (1e15 / 4.0 / PI));
  S.nth_s2[(size_t) v * NTH_SOL + k] = a2 * normfac;
}

// per vector: source normalisation, normalisation-change factors and the returning-radiation flux-correction factors
// (src/Xillspec.cpp:179-205,344-362,408-440,500-526, src/PrimarySource.h:279-293)
__global__ void __launch_bounds__(64) k_nth_finish(const VPar *__restrict__ vps, DevTables T, Scratch S) {
  const int v = blockIdx.x, t = threadIdx.x;
  if (S.status[v] != ST_OK) return;
  if (S.reuse && (S.reuse[v] & REUSE_ALL)) return;
  const VPar &vp = vps[v];
  if (vp.prim_type != PRIM_NTHCOMP) return;
  const int nz = vp.nz;
  const double *s_nfac = S.nth_nfac + (size_t) v * NTH_SOL, *s_s2 = S.nth_s2 + (size_t) v * NTH_SOL;
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
  const double *sp = S.nth_spt + (size_t) v * NTH_MAX;
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
  const double *sp = S.nth_spt + (size_t) v * NTH_MAX;
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

int nth_kernel_init() {
  if (cudaFuncSetAttribute(k_nth, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(NthSmem)) != cudaSuccess) return 1;
  // three CTAs per SM: all of the shared-memory carve-out
  return cudaFuncSetAttribute(k_nth, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared) == cudaSuccess ? 0 : 1;
}
void launch_nth(const VPar *vps, const DevTables &T, const Scratch &S, long n, int nz_max, cudaStream_t st) {
  const int per_vec = nz_max + 1;
  const long total = n * per_vec;
  k_nth<<<(unsigned) ((total + NTH_NT - 1) / NTH_NT), NTH_NT, sizeof(NthSmem), st>>>(vps, T, S, n, per_vec);
  k_nth_finish<<<(unsigned) n, 64, 0, st>>>(vps, T, S);
}
void launch_prim_nth(const VPar *vps, const DevTables &T, const Scratch &S, long n, double *total, const double *user_e,
                     int n_flux, double *out, int renorm3, cudaStream_t st) {
  k_prim_nth<<<(unsigned) n, 256, 0, st>>>(vps, T, S, total, user_e, n_flux, out, renorm3);
}

}  // namespace rx
