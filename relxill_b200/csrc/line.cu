// line.cu — the relline profile kernel (the FP64 hot loop of the relativistic smearing).
// Compiled with FMA contraction on (build.py): its results feed no discrete decision other than the
// Romberg convergence test, which has a 2 % threshold.
//
// Replaces calc_relline_profile + integ_relline_bin + int_edge + int_romb + romberg_integration +
// relb_func (src/Relprofile.cpp:489-726,835-905) and the division by the bin energy of
// renorm_relline_profile (:757-762).
//
// Mapping.  One CTA (4 warps) per (vector, radial zone).  The zone's radii are processed in sub-batches; inside
// a sub-batch all (radius, energy-bin) pairs that the reference's double loop visits are flattened into one
// dense item list (the bins of one radius are contiguous: [ielo, iehi]), so lanes stay busy whatever the
// line width is.  The work is sorted by cost so that the lanes of a warp do the same thing:
//   set-up   lane = radius, one warp per task: first bin, last bin (closed-form index on the logarithmic
//            convolution grid, corrected against the tabulated edges) and the integrand at the two edge
//            nodes g* = h, 1-h that int_edge needs (once per radius instead of once per edge bin);
//            offsets by a warp scan.
//   pass A   one thread per item: analytic edge terms and the midpoint-rule bins (E < 0.95) are finished;
//            the bins that take the Romberg path are gathered in a dense list.
//   pass B   one thread per Romberg bin: two halvings (five abscissae) unconditionally; a bin that has not
//            converged by then (the horns of the profile, a few per cent of the bins) goes on a work list.
//   phase 2  one HALF WARP per listed bin: the 16 new abscissae of levels <= 4 are evaluated in parallel
//            across the lanes, the level sums come from one xor-butterfly, and the tableaus of 32 bins are
//            then built in parallel, one per lane.  Levels 5-6 (rare) are finished by the owning lane.
//   phase 3  per energy bin, the sub-batch's contributions are added in ascending-radius order into the
//            zone's output row -> no atomics on data, bit-reproducible, the reference's summation order.
//
// Arithmetic.  The two branches k = 0, 1 of the transfer function share everything but the interpolated
// trff value, so one evaluation of the integrand returns both (the reference calls relb_func twice).
// Romberg level n re-uses the function values of level n-1 (the reference re-evaluates them; the
// abscissae a + ii * pas are bit-identical because pas is halved exactly).  The integrand
//   pow(eg,3) / ((gmax-gmin) * sqrt(g* - g*^2)) * ftrf * emis          (src/Relprofile.cpp:506)
// is evaluated as eg^3 * rsqrt(g* - g*^2) * ftrf * (emis / (gmax-gmin)); the g* bracket is computed
// arithmetically (the grid is uniform), which can differ from the reference's binary search only when g*
// sits within an ulp of a node, where the piecewise-linear interpolant is continuous.
#include <cuda_runtime.h>

#include <cmath>

#include "common.h"
#include "devutil.cuh"
#include "kernels.h"

namespace rx {

struct RelbCtx {
  double gmin, gmax, del_g;
  double scale;          // emis / (gmax - gmin)
  const double2 *trff;   // [NG] {branch 0, branch 1} of this radius (global, L1-resident)
  const double2 *cosne;
  int limb;
};

#define GS_C ((1.0 - 2 * GFAC_H) / (NG - 1))
#define GS_INVC ((NG - 1) / (1.0 - 2 * GFAC_H))
// An FP64 immediate whose low word is not zero costs two moves at every use (a tenth of this kernel's instructions
// were such moves); as __constant__ data the same values ride in the instruction as constant-bank operands.
struct LnConst {
  double h, one_m_h, gs_c, gs_invc, prec, e95, sqrt_h;
  double inv[7];   // 1 / (4^ii - 1) of the Richardson step
};
__constant__ LnConst LK = {GFAC_H, 1.0 - GFAC_H, GS_C, GS_INVC, 0.02, 1.0 * 0.95, 0.0,
                           {0.0, 1.0 / 3.0, 1.0 / 15.0, 1.0 / 63.0, 1.0 / 255.0, 1.0 / 1023.0, 1.0 / 4095.0}};

// limb darkening / brightening factor of relb_func (src/Relprofile.cpp:508-518); out of line: the default
// law is isotropic and the logarithm would otherwise be replicated into every copy of the integrand
__device__ __noinline__ double2 relb2_limb(int ind, double inte, const double2 *cosne, int limb) {
  const double inte1 = 1.0 - inte;
  const double2 c0 = __ldg(cosne + ind), c1 = __ldg(cosne + ind + 1);
  const double m0 = inte * c0.x + inte1 * c1.x, m1 = inte * c0.y + inte1 * c1.y;
  if (limb == 1) return make_double2(1.0 + 2.06 * m0, 1.0 + 2.06 * m1);
  if (limb == 2) return make_double2(log(1.0 + 1.0 / m0), log(1.0 + 1.0 / m1));
  return make_double2(1.0, 1.0);
}

// both branches of relb_func (src/Relprofile.cpp:489-521) at energy eg
__device__ __forceinline__ void relb2(double eg, const RelbCtx &c, double &v0, double &v1) {
  const double egstar = (eg - c.gmin) * c.del_g;
  int ind = (int) ((egstar - LK.h) * LK.gs_invc);
  ind = ind < 0 ? 0 : (ind > NG - 2 ? NG - 2 : ind);
  const double inte = (egstar - (LK.h + LK.gs_c * (double) ind)) * LK.gs_invc;
  const double inte1 = 1.0 - inte;
  const double2 t0 = __ldg(c.trff + ind), t1 = __ldg(c.trff + ind + 1);
  const double common = (eg * eg * eg) * rsqrt(egstar - egstar * egstar) * c.scale;
  v0 = common * (inte * t0.x + inte1 * t1.x);   // (the reference's weights: inte on node ind, 1-inte on ind+1)
  v1 = common * (inte * t0.y + inte1 * t1.y);
  if (c.limb != 0) {
    const double2 f = relb2_limb(ind, inte, c.cosne, c.limb);
    v0 *= f.x;
    v1 *= f.y;
  }
}

// (obtprec > prec) of the reference, obtprec = fabs(t_new - t_old) / t_new  (src/Relprofile.cpp:575)
__device__ __forceinline__ bool not_converged(double t_new, double t_old) {
  const double d = fabs(t_new - t_old);
  if (t_new > 0.0) return d > LK.prec * t_new;
  if (t_new == 0.0) return d > 0.0;   // x/0 = inf > prec;  0/0 = NaN compares false
  return false;                        // negative (or NaN) quotient ends the loop
}

__device__ __forceinline__ double gstar2ener(double g, double gmin, double gmax) { return (g * (gmax - gmin) + gmin) * 1.0; }

// Richardson step of the Romberg tableau, t[ii] = (4^ii t[ii-1] - tprev[ii-1]) / (4^ii - 1), with the
// divisions replaced by the tabulated reciprocals
__device__ __forceinline__ double richardson(int ii, double cur_lo, double prev_lo) {
  const double r4[7] = {1.0, 4.0, 16.0, 64.0, 256.0, 1024.0, 4096.0};
  return (r4[ii] * cur_lo - prev_lo) * LK.inv[ii];
}

// grid_mode 0: the fixed convolution grid; 1: the caller's grid shifted by (1+z) and divided by lineE
// per vector (XspecSpectrum::shift_energy_grid_redshift / _1keV, src/XspecSpectrum.h:61-76)
__device__ __forceinline__ double line_edge(const double *egrid, int j, int grid_mode, double z, double lineE) {
  double e = __ldg(egrid + j);
  if (grid_mode) {
    if (z > 0) e *= (1 + z);
    e /= lineE;
  }
  return e;
}
// binary_search(ener, n+1, val) of the reference on the (possibly rescaled) grid: the last index k <= n_edges-2
// with edge[k] <= val (0 if none).  On the logarithmic convolution grid the index is computed in closed form
// and then corrected against the tabulated edges, so the result is the search's, without its dependent loads.
struct LineGrid {
  const double *e;
  int n_ener, mode;
  double log_lo, inv_dlog;   // mode 0: edge[k] ~ exp(log_lo + k / inv_dlog)
};
template <int GRID_MODE>
__device__ int line_index(const LineGrid &G, double val, double z, double lineE) {
  const int last = G.n_ener - 1;   // n_edges - 2
  if (GRID_MODE == 0) {
    int k = (int) floor((log(val) - G.log_lo) * G.inv_dlog);
    k = k < 0 ? 0 : (k > last ? last : k);
    while (k < last && __ldg(G.e + k + 1) <= val) k++;
    while (k > 0 && __ldg(G.e + k) > val) k--;
    return k;
  }
  int klo = 0, khi = G.n_ener;
  while (khi - klo > 1) {
    const int k = (khi + klo) >> 1;
    if (line_edge(G.e, k, 1, z, lineE) > val) khi = k; else klo = k;
  }
  return klo;
}

constexpr int LN_NT = 128;
constexpr int LN_BUF = 1024;     // contribution slots (items) per sub-batch
constexpr int LN_MAXR = 32;      // radii per sub-batch: one lane each in the set-up warp
constexpr int LN_MAXDEF = 256;   // work list of the deep Romberg bins
struct LnRad {
  double gmin, gmax, del_g, scale, weight;
  double nlo, nhi;               // `norm` of int_edge (src/Relprofile.cpp:585-621) at g* = h and g* = 1-h
  int ielo, iehi, off, gi;
};
struct LnSmem {
  double contrib[LN_BUF];
  double def_a[LN_MAXDEF], def_b[LN_MAXDEF];      // work list of phase 2: interval still to integrate
  double def_f[LN_MAXDEF][2];                     // ... and the integrand at its lower end (both branches)
  LnRad rad[LN_MAXR + 1];
  unsigned short rlist[LN_BUF];                   // items that take the Romberg path
  unsigned short def_item[LN_MAXDEF];
  unsigned char item_rad[LN_BUF];                 // sub-batch radius of every item (bit 7: retry flag)
  int nrad, nrom, ndef, overflow, cursor, jlo, jhi, resume;   // resume: first bin still to do of radius `cursor` (-1 = all)
  int zjlo, zjhi;
};

__device__ __forceinline__ void ln_ctx(const LnRad &lr, const double2 *g_trff, const double2 *g_cosne, int limb, RelbCtx &c) {
  c.gmin = lr.gmin; c.gmax = lr.gmax; c.del_g = lr.del_g; c.scale = lr.scale;
  c.trff = g_trff + (size_t) lr.gi * NG; c.cosne = g_cosne + (size_t) lr.gi * NG; c.limb = limb;
}

// int_edge (src/Relprofile.cpp:585-621) with the integrand at the edge node already summed into `norm`
__device__ __forceinline__ double edge_term(double blo, double bhi, double norm, double gmin, double gmax) {
  double lo, hi;
  if (blo <= 0.5) { lo = blo; hi = bhi; }
  else { lo = 1.0 - bhi; hi = 1.0 - blo; }
  return 2 * norm * (sqrt(hi) - sqrt(lo)) * 1.0 * (gmax - gmin);
}

// The decision part of integ_relline_bin (src/Relprofile.cpp:650-726): the analytic edge terms of the bin
// (returned in flu when EDGES) and the interval [rlo, rhi] left for quadrature.  Returns 0: nothing left,
// 1: midpoint rule, 2: Romberg (int_romb, :628-647).
template <bool EDGES>
__device__ __forceinline__ int bin_split(const LnRad &lr, double rlo0, double rhi0, double &rlo, double &rhi, double &flu) {
  flu = 0.0;
  double gblo = (rlo0 / 1.0 - lr.gmin) * lr.del_g;
  if (gblo < 0.0) gblo = 0.0; else if (gblo > 1.0) gblo = 1.0;
  double gbhi = (rhi0 / 1.0 - lr.gmin) * lr.del_g;
  if (gbhi < 0.0) gbhi = 0.0; else if (gbhi > 1.0) gbhi = 1.0;
  if (gbhi == 0) return 0;
  rlo = rlo0; rhi = rhi0;
  const bool at_lo = gblo <= LK.h, at_hi = gbhi >= LK.one_m_h;
  double lo_hhi = LK.h, hi_hlo = LK.one_m_h;
  if (at_lo) {
    rlo = gstar2ener(LK.h, lr.gmin, lr.gmax);
    if (gbhi <= LK.h) { lo_hhi = gbhi; rlo = -1.0; }
  }
  if (at_hi) {
    rhi = gstar2ener(LK.one_m_h, lr.gmin, lr.gmax);
    if (gblo >= LK.one_m_h) { hi_hlo = gblo; rhi = -1.0; }
  }
  if (EDGES && (at_lo || at_hi)) {   // lower-edge term first, like the reference; one call unless the bin spans both edges
    const double eb_lo = at_lo ? gblo : hi_hlo, eb_hi = at_lo ? lo_hhi : gbhi, e_norm = at_lo ? lr.nlo : lr.nhi;
    flu = flu + edge_term(eb_lo, eb_hi, e_norm, lr.gmin, lr.gmax);
    if (at_lo && at_hi) flu = flu + edge_term(hi_hlo, gbhi, lr.nhi, lr.gmin, lr.gmax);
  }
  if ((rhi >= 0) && (rlo >= 0)) return (rlo >= LK.e95) ? 2 : 1;
  return 0;
}

// Romberg levels 5 and 6 of one bin (rare: a fraction of a per cent of the listed bins), by the thread that owns
// the bin's tableau.  The new abscissae of a level are summed in ascending order like the reference's loop.
// Everything is passed by value so that the caller's tableau stays in registers.
struct DeepIn {
  double tp[2][5];   // tableau rows after level 4
  double res[2];
  int done[2];
};
__device__ __noinline__ double romberg_deep(double a, double pas, RelbCtx c, DeepIn in) {
  double tprev[2][7], res[2] = {in.res[0], in.res[1]};
  bool done[2] = {in.done[0] != 0, in.done[1] != 0};
  for (int k = 0; k < 2; k++)
    for (int ii = 0; ii < 5; ii++) tprev[k][ii] = in.tp[k][ii];
  const double pas6 = pas / 64.0;
  double pasn = pas / 16.0;
  double sum[2] = {tprev[0][0] / pasn, tprev[1][0] / pasn};   // level-4 trapezoid sums, from cur[0] = sum * pasn
  double last[2] = {res[0], res[1]};
  for (int n = 5; n <= 6; n++) {
    if (done[0] && done[1]) break;
    pasn = pasn * 0.5;
    double o[2] = {0.0, 0.0};
    const int step = (n == 5) ? 4 : 2, first = (n == 5) ? 2 : 1;   // level 5: p = 2 mod 4, level 6: odd p (of 64)
    for (int p = first; p < 64; p += step) {
      double w0, w1;
      relb2(a + pas6 * p, c, w0, w1);
      o[0] += w0;
      o[1] += w1;
    }
    for (int k = 0; k < 2; k++) {
      sum[k] += o[k];
      if (done[k]) continue;
      double cur[7];
      cur[0] = sum[k] * pasn;
      for (int ii = 1; ii <= n; ii++) cur[ii] = richardson(ii, cur[ii - 1], tprev[k][ii - 1]);
      res[k] = cur[n];
      if (!not_converged(cur[n], last[k])) done[k] = true;
      last[k] = cur[n];
      for (int ii = 0; ii <= n; ii++) tprev[k][ii] = cur[ii];
    }
  }
  return res[0] + res[1];
}

// 9 resident CTAs per SM (56 registers, 36 warps): the kernel is latency-bound on its dependent FP64 chains, and the
// extra warps are worth more than the few spilled values (measured 11.8 ms at 6 CTAs / 80 registers, 10.0 ms at 8 / 64,
// 9.4 ms at 9 / 56; at 10 / 48 the spills win: 10.4 ms)
template <int GRID_MODE>
__global__ void __launch_bounds__(LN_NT, 9) k_line(const VPar *__restrict__ vps, DevTables T, Scratch S, LineGrid G,
                                                   int ne_stride, int nz_stride) {
  __shared__ __align__(16) LnSmem sm;
  const unsigned FULL = 0xffffffffu;
  const int v = blockIdx.y, z = blockIdx.x, t = threadIdx.x;
  if (S.status[v] != ST_OK) return;
  if (S.reuse && S.reuse[v]) return;   // the zone profiles of the previous run stand
  const VPar &vp = vps[v];
  if (z >= vp.nz) return;
  const double *egrid = G.e;
  const int n_ener = G.n_ener;
  constexpr int grid_mode = GRID_MODE;
  const double zred = vp.z, lineE = vp.lineE;
  const int limb = vp.limb;
  const double e_first = line_edge(egrid, 0, grid_mode, zred, lineE);
  const double e_last = line_edge(egrid, n_ener, grid_mode, zred, lineE);
  // radii of this zone (izone[] is non-increasing along the descending-radius fine grid; k_syspar tabulated
  // the first index of every zone)
  const int ia = S.zfirst[(size_t) v * (NZMAX + 1) + z + 1], ib = S.zfirst[(size_t) v * (NZMAX + 1) + z];
  double *flux = S.relflux + ((size_t) v * nz_stride + z) * ne_stride;   // doubles as the zone accumulator
  const double *g_re = S.re + (size_t) v * NR;
  const double2 *g_trff = reinterpret_cast<const double2 *>(S.trff) + (size_t) v * NR * NG;
  const double2 *g_cosne = reinterpret_cast<const double2 *>(S.cosne) + (size_t) v * NR * NG;
  if (t == 0) { sm.cursor = ia; sm.resume = -1; sm.zjlo = n_ener; sm.zjhi = -1; }
  __syncthreads();

  while (true) {
    const int cur = sm.cursor;
    if (cur >= ib) break;
    const int resume = sm.resume;
    const int zjlo_old = sm.zjlo, zjhi_old = sm.zjhi;
    // ---- sub-batch set-up: radius r = lane, one warp per task (first bin | last bin | edge norm at g* = h | at 1-h)
    {
      const int r = t & 31, task = t >> 5;
      const int i = cur + r;
      LnRad &lr = sm.rad[r];
      if (i < ib) {
        const double gmin = S.gmin[(size_t) v * NR + i], gmax = S.gmax[(size_t) v * NR + i];
        const double del_g = 1. / (gmax - gmin);
        const double scale = del_g * S.emis[(size_t) v * NR + i];
        const bool on_grid = (gmax > e_first) && (gmin < e_last);  // src/Relprofile.cpp:863-878
        if (task == 0) {
          lr.gmin = gmin; lr.gmax = gmax; lr.del_g = del_g; lr.scale = scale;
          lr.weight = trapez_single(g_re, i, NR) / 2;
          lr.gi = i;
          int ielo = 0;
          if (on_grid) {
            ielo = line_index<GRID_MODE>(G, gmin < e_first ? e_first : gmin, zred, lineE);
            if (r == 0 && resume >= 0) ielo = resume;   // rest of a radius wider than the buffer
          }
          lr.ielo = ielo;
        } else if (task == 1) {
          lr.iehi = on_grid ? line_index<GRID_MODE>(G, gmax > e_last ? e_last : gmax, zred, lineE) : -1;
        } else {
          RelbCtx c;
          c.gmin = gmin; c.gmax = gmax; c.del_g = del_g; c.scale = scale;
          c.trff = g_trff + (size_t) i * NG; c.cosne = g_cosne + (size_t) i * NG; c.limb = limb;
          double n0, n1;
          relb2(gstar2ener(task == 2 ? GFAC_H : 1.0 - GFAC_H, gmin, gmax), c, n0, n1);
          double norm = 0.0;
          norm = norm + n0;
          norm = norm + n1;
          norm = norm * sqrt(GFAC_H);
          if (task == 2) lr.nlo = norm; else lr.nhi = norm;
        }
      } else if (task == 0) {
        lr.gi = i; lr.ielo = 0;
      } else if (task == 1) {
        lr.iehi = -1;
      }
    }
    __syncthreads();
    if (t >= 32) {
      // while warp 0 scans, the other warps request the transfer-function rows of the sub-batch's radii (640 bytes =
      // five lines each): the integrand's table look-ups (data-dependent, one L2 round trip each) then hit in L1
      const int nr = min(ib - cur, LN_MAXR);
      for (int q = t - 32; q < nr * 5; q += LN_NT - 32)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char *>(g_trff + (size_t) cur * NG) + (size_t) q * 128));
    }
    if (t < 32) {   // offsets of the radii that fit the buffer: warp scan over the bin counts
      const int r = t;
      const bool valid = cur + r < ib;
      int ielo = sm.rad[r].ielo, iehi = sm.rad[r].iehi;
      int w = valid ? max(iehi - ielo + 1, 0) : 0;
      int next_resume = -1;
      if (r == 0 && w > LN_BUF) {                       // a single radius wider than the buffer: take a piece
        w = LN_BUF;
        next_resume = ielo + w;
        iehi = next_resume - 1;
        sm.rad[0].iehi = iehi;
      }
      next_resume = __shfl_sync(FULL, next_resume, 0);
      int cum = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(FULL, cum, o);
        if (r >= o) cum += up;
      }
      const unsigned fits = __ballot_sync(FULL, valid && cum <= LN_BUF && (next_resume < 0 || r == 0));
      const int n = __popc(fits);                       // `fits` is a prefix: cum is monotone, valid a prefix
      const bool mine = r < n;
      int jlo = (mine && w > 0) ? ielo : n_ener, jhi = (mine && w > 0) ? iehi : -1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        jlo = min(jlo, __shfl_xor_sync(FULL, jlo, o));
        jhi = max(jhi, __shfl_xor_sync(FULL, jhi, o));
      }
      if (mine) sm.rad[r].off = cum - w;
      if (r == n - 1) sm.rad[n].off = cum;
      if (r == 0) {
        sm.nrad = n;
        sm.nrom = 0;
        sm.ndef = 0;
        sm.overflow = 0;
        sm.jlo = jlo;
        sm.jhi = jhi;
        sm.zjlo = min(zjlo_old, jlo);
        sm.zjhi = max(zjhi_old, jhi);
        sm.resume = next_resume;
        sm.cursor = (next_resume >= 0) ? cur + n - 1 : cur + n;
      }
    }
    __syncthreads();
    const int nrad = sm.nrad;
    const int nitems = sm.rad[nrad].off;
    for (int r = t >> 5; r < nrad; r += LN_NT / 32)   // item -> radius map, one warp per radius
      for (int q = sm.rad[r].off + (t & 31); q < sm.rad[r + 1].off; q += 32) sm.item_rad[q] = (unsigned char) r;
    __syncthreads();
    // ---- pass A: one thread per item.  Edge terms and midpoint bins are finished here; the bins that take the
    // Romberg path are gathered in a dense list (order irrelevant: every bin's value is computed on its own).
    for (int base = 0; base < nitems; base += LN_NT) {
      const int item = base + t;
      int cls = 0;
      if (item < nitems) {
        const LnRad &lr = sm.rad[sm.item_rad[item]];
        const int j = lr.ielo + (item - lr.off);
        const double elo = line_edge(egrid, j, grid_mode, zred, lineE), ehi = line_edge(egrid, j + 1, grid_mode, zred, lineE);
        double rlo, rhi, flu;
        cls = bin_split<true>(lr, elo, ehi, rlo, rhi, flu);
        if (cls == 1) {
          RelbCtx c;
          ln_ctx(lr, g_trff, g_cosne, limb, c);
          double m0, m1;
          relb2((rhi + rlo) / 2.0, c, m0, m1);
          double f2 = 0.0;
          f2 += m0 * (rhi - rlo);
          f2 += m1 * (rhi - rlo);
          flu = flu + f2;
        }
        sm.contrib[item] = flu;
      }
      const unsigned rom = __ballot_sync(FULL, cls == 2);
      if (rom) {
        int pos = 0;
        if ((t & 31) == 0) pos = atomicAdd(&sm.nrom, __popc(rom));
        pos = __shfl_sync(FULL, pos, 0);
        if (cls == 2) sm.rlist[pos + __popc(rom & ((1u << (t & 31)) - 1))] = (unsigned short) item;
      }
    }
    __syncthreads();
    const int nrom = sm.nrom;
    for (int round = 0;; round++) {
      // ---- pass B: one thread per Romberg bin, two halvings (five abscissae) evaluated unconditionally.  A bin that
      // has not converged by then (the horns of the profile) goes on the work list of phase 2; if the list is
      // full the bin is flagged (bit 7 of item_rad) and retried after phase 2 has drained the list, so the
      // result never depends on the order in which threads reach the list.
      for (int base = 0; base < nrom; base += LN_NT) {
        const int q = base + t;
        bool defer = false;
        int item = 0;
        double ra = 0.0, rb = 0.0, fa0 = 0.0, fa1 = 0.0;
        if (q < nrom) {
          item = sm.rlist[q];
          const int ir = sm.item_rad[item];
          if (round == 0 || (ir & 0x80)) {
            const LnRad &lr = sm.rad[ir & 0x7f];
            const int j = lr.ielo + (item - lr.off);
            const double elo = line_edge(egrid, j, grid_mode, zred, lineE), ehi = line_edge(egrid, j + 1, grid_mode, zred, lineE);
            double flu;
            bin_split<false>(lr, elo, ehi, ra, rb, flu);
            RelbCtx c;
            ln_ctx(lr, g_trff, g_cosne, limb, c);
            // Romberg on [ra, rb] for both branches (src/Relprofile.cpp:524-579), levels 0..2
            const double pas = rb - ra, pas1 = pas / 2.0, pas2 = pas1 / 2.0;
            // abscissae in the order a, b, a+pas/4, a+pas/2, a+3pas/4: the level-2 trapezoid sum is built as
            // ((ta + f(q1)) + f(mid)) + f(q3), the reference's ascending order.  Rolled: one copy of the integrand.
            double ta[2] = {0.0, 0.0}, fm[2] = {0.0, 0.0}, x2[2] = {0.0, 0.0};
#pragma unroll 1
            for (int pt = 0; pt < 5; pt++) {
              const double eg = (pt == 0) ? ra : (pt == 1) ? rb : (pt == 2) ? ra + pas2 * 1 : (pt == 3) ? ra + pas1 * 1 : ra + pas2 * 3;
              double w0, w1;
              relb2(eg, c, w0, w1);
              if (pt == 0) { fa0 = w0; fa1 = w1; }
              else if (pt == 1) { ta[0] = (fa0 + w0) / 2.0; ta[1] = (fa1 + w1) / 2.0; x2[0] = ta[0]; x2[1] = ta[1]; }
              else {
                if (pt == 3) { fm[0] = w0; fm[1] = w1; }
                x2[0] += w0;
                x2[1] += w1;
              }
            }
            double rsum = 0.0;
#pragma unroll
            for (int k = 0; k < 2; k++) {
              const double t00 = ta[k] * pas;
              const double t01 = (ta[k] + fm[k]) * pas1;
              const double t10 = richardson(1, t01, t00);
              double r = t10;
              if (not_converged(t10, t00)) {
                const double t02 = x2[k] * pas2;
                const double t11 = richardson(1, t02, t01);
                const double t20 = richardson(2, t11, t10);
                defer |= not_converged(t20, t10);
                r = t20;
              }
              rsum += r;
            }
            sm.item_rad[item] = (unsigned char) (ir & 0x7f);
            if (!defer) sm.contrib[item] = sm.contrib[item] + rsum;
          }
        }
        const unsigned dm = __ballot_sync(FULL, defer);
        if (dm) {
          int pos = 0;
          if ((t & 31) == 0) pos = atomicAdd(&sm.ndef, __popc(dm));
          pos = __shfl_sync(FULL, pos, 0);
          if (defer) {
            const int d = pos + __popc(dm & ((1u << (t & 31)) - 1));
            if (d < LN_MAXDEF) {
              sm.def_item[d] = (unsigned short) item;
              sm.def_a[d] = ra;
              sm.def_b[d] = rb;
              sm.def_f[d][0] = fa0;
              sm.def_f[d][1] = fa1;
            } else {
              sm.item_rad[item] |= 0x80;
              sm.overflow = 1;
            }
          }
        }
      }
      __syncthreads();
      // ---- phase 2: the listed bins at full Romberg depth.  One HALF WARP evaluates the 16 new abscissae of levels
      // 1..4 of a bin in parallel (lane h takes point h+1 of 16; f(a) comes from pass B); one xor-butterfly
      // gives the sums of the points that are new at each level.  The lane whose index equals the iteration
      // keeps them, so after 16 iterations every lane owns one bin and all tableaus are built in parallel.
      {
        const int ndef = min(sm.ndef, LN_MAXDEF);
        constexpr int NH = LN_NT / 16;
        const int half = t >> 4, h = t & 15, hb = t & 16;
        for (int g0 = 0; g0 < ndef; g0 += 16 * NH) {
          const int nit = min(16, (ndef - g0 + NH - 1) / NH);
          double kn[2][4], kfb[2];
#pragma unroll
          for (int k = 0; k < 2; k++) { kfb[k] = 0.0; kn[k][0] = kn[k][1] = kn[k][2] = kn[k][3] = 0.0; }
          for (int it = 0; it < nit; it++) {
            const int d = g0 + it * NH + half;
            double v0 = 0.0, v1 = 0.0;
            if (d < ndef) {
              RelbCtx c;
              ln_ctx(sm.rad[sm.item_rad[sm.def_item[d]] & 0x7f], g_trff, g_cosne, limb, c);
              const double a = sm.def_a[d], b = sm.def_b[d];
              const double pas4 = (b - a) / 16.0;
              relb2(h == 15 ? b : a + pas4 * (h + 1), c, v0, v1);
            }
#pragma unroll
            for (int k = 0; k < 2; k++) {
              const double vv = k ? v1 : v0;
              const double fb = __shfl_sync(FULL, vv, hb + 15);
              // class sums over the interior points p = 1..15 (lane h = p - 1): p = 8; p = 4 mod 8; p = 2 mod 4; odd p
              double x = (h == 15) ? 0.0 : vv;
              const double n1 = __shfl_sync(FULL, x, hb + 7);
              x += __shfl_xor_sync(FULL, x, 8);
              const double n2 = __shfl_sync(FULL, x, hb + 3);
              x += __shfl_xor_sync(FULL, x, 4);
              const double n3 = __shfl_sync(FULL, x, hb + 1);
              x += __shfl_xor_sync(FULL, x, 2);
              const double n4 = __shfl_sync(FULL, x, hb + 0);
              if (h == it) { kfb[k] = fb; kn[k][0] = n1; kn[k][1] = n2; kn[k][2] = n3; kn[k][3] = n4; }
            }
          }
          const int d = g0 + h * NH + half;   // the bin this lane kept
          if (h < nit && d < ndef) {
            const int item = sm.def_item[d];
            const double a = sm.def_a[d], b = sm.def_b[d];
            const double pas = b - a;
            double tprev[2][5], res[2] = {0.0, 0.0};
            bool done[2] = {false, false};
#pragma unroll
            for (int k = 0; k < 2; k++) {
              const double ta = (sm.def_f[d][k] + kfb[k]) / 2.0;
              tprev[k][0] = ta * pas;
              double last = tprev[k][0], pasn = pas, sum = ta;
#pragma unroll
              for (int n = 1; n <= 4; n++) {
                pasn = pasn * 0.5;
                sum += kn[k][n - 1];
                if (!done[k]) {
                  double cur[5];
                  cur[0] = sum * pasn;
#pragma unroll
                  for (int ii = 1; ii <= 4; ii++) if (ii <= n) cur[ii] = richardson(ii, cur[ii - 1], tprev[k][ii - 1]);
                  res[k] = cur[n];
                  if (!not_converged(cur[n], last)) done[k] = true;
                  last = cur[n];
#pragma unroll
                  for (int ii = 0; ii <= 4; ii++) if (ii <= n) tprev[k][ii] = cur[ii];
                }
              }
            }
            double rtot = res[0] + res[1];
            if (!(done[0] && done[1])) {
              RelbCtx c;
              ln_ctx(sm.rad[sm.item_rad[item] & 0x7f], g_trff, g_cosne, limb, c);
              DeepIn in;
#pragma unroll
              for (int k = 0; k < 2; k++) {
#pragma unroll
                for (int ii = 0; ii < 5; ii++) in.tp[k][ii] = tprev[k][ii];
                in.res[k] = res[k];
                in.done[k] = done[k] ? 1 : 0;
              }
              rtot = romberg_deep(a, pas, c, in);
            }
            sm.contrib[item] = sm.contrib[item] + rtot;
          }
        }
      }
      const int again = sm.overflow;
      __syncthreads();
      if (!again) break;
      if (t == 0) { sm.ndef = 0; sm.overflow = 0; }
      __syncthreads();
    }
    // ---- phase 3: ordered accumulation (ascending radius index = the reference's loop order) into the zone's
    // row.  Bins joining the zone's range start from zero; bins the sub-batch does not touch stay as they are.
    {
      const int jlo = sm.jlo, jhi = sm.jhi;
      const int nlo = min(zjlo_old, jlo), nhi = max(zjhi_old, jhi);
      for (int j = nlo + t; j <= nhi; j += LN_NT) {
        const bool was = (j >= zjlo_old && j <= zjhi_old), now = (j >= jlo && j <= jhi);
        if (was && !now) continue;
        double a = was ? flux[j] : 0.0;
        if (now) {
          for (int r = 0; r < nrad; r++) {
            const LnRad &lr = sm.rad[r];
            if (j >= lr.ielo && j <= lr.iehi) a += sm.contrib[lr.off + (j - lr.ielo)] * lr.weight;
          }
        }
        flux[j] = a;
      }
    }
    __syncthreads();
  }
  // only the bins this zone touched are written; the range travels with the row
  const int zjlo = sm.zjlo, zjhi = sm.zjhi;
  if (t == 0) {
    S.zrange[((size_t) v * NZMAX + z) * 2] = zjlo;
    S.zrange[((size_t) v * NZMAX + z) * 2 + 1] = zjhi;
  }
  for (int j = zjlo + t; j <= zjhi; j += LN_NT) {
    const double elo = line_edge(egrid, j, grid_mode, zred, lineE), ehi = line_edge(egrid, j + 1, grid_mode, zred, lineE);
    flux[j] = flux[j] / (0.5 * (elo + ehi));
  }
}

// ---------------------------------------------------------------------------------- launcher
int line_kernel_init() {
  cudaError_t e = cudaFuncSetAttribute(k_line<0>, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(k_line<1>, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared);
  return e == cudaSuccess ? 0 : 1;
}

void launch_line(const VPar *vps, const DevTables &T, const Scratch &S, long n, const double *egrid, int n_ener,
                 int grid_mode, int nz_max, cudaStream_t st) {
  dim3 grid(nz_max, (unsigned) n);
  LineGrid G;
  G.e = egrid; G.n_ener = n_ener; G.mode = grid_mode;
  G.log_lo = std::log(CONV_EMIN);
  G.inv_dlog = (double) NCONV / (std::log(CONV_EMAX) - std::log(CONV_EMIN));
  if (grid_mode == 0) k_line<0><<<grid, LN_NT, 0, st>>>(vps, T, S, G, S.ne_line_cap, S.nz_cap);
  else k_line<1><<<grid, LN_NT, 0, st>>>(vps, T, S, G, S.ne_line_cap, S.nz_cap);
}
int line_max_bins() { return 1 << 24; }   // the zone accumulator lives in the output row: no shared-memory limit

}  // namespace rx
