// line.cu — the relline profile kernel (the FP64 hot loop of the relativistic smearing).
// Compiled with FMA contraction on (build.py): its results feed no discrete decision other than the
// Romberg convergence test, which has a 2 % threshold.
//
// Replaces calc_relline_profile + integ_relline_bin + int_edge + int_romb + romberg_integration +
// relb_func (src/Relprofile.cpp:489-726,835-905) and the division by the bin energy of
// renorm_relline_profile (:757-762).
//
// Mapping.  One CTA per (vector, radial zone).  The zone's radii are processed in sub-batches; inside a
// sub-batch all (radius, energy-bin) pairs that the reference's double loop visits are flattened into one
// dense item list (the bins of one radius are contiguous: [ielo, iehi]), so lanes stay busy whatever the
// line width is.
//   phase 1  one thread per item.  Edge terms, midpoint bins and Romberg bins up to two halvings are
//            finished here; a bin whose Romberg integral has not converged by then (the horns of the
//            profile, a few per cent of the bins) is pushed on a shared-memory work list.
//   phase 2  one HALF WARP per listed bin: the dyadic abscissae of the deeper Romberg levels are
//            evaluated in parallel across the lanes (17 points for levels <= 4, 65 for levels <= 6) and
//            the tableau is built from class sums of one xor-butterfly.
//   phase 3  per energy bin, the sub-batch's contributions are added in ascending-radius order into the
//            zone accumulator -> no atomics on data, bit-reproducible, the reference's summation order.
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

// both branches of relb_func (src/Relprofile.cpp:489-521) at energy eg
__device__ __forceinline__ void relb2(double eg, const RelbCtx &c, double &v0, double &v1) {
  const double egstar = (eg - c.gmin) * c.del_g;
  int ind = (int) ((egstar - GFAC_H) * GS_INVC);
  ind = ind < 0 ? 0 : (ind > NG - 2 ? NG - 2 : ind);
  const double inte = (egstar - (GFAC_H + GS_C * (double) ind)) * GS_INVC;
  const double inte1 = 1.0 - inte;
  const double2 t0 = __ldg(c.trff + ind), t1 = __ldg(c.trff + ind + 1);
  const double common = (eg * eg * eg) * rsqrt(egstar - egstar * egstar) * c.scale;
  v0 = common * (inte * t0.x + inte1 * t1.x);   // (the reference's weights: inte on node ind, 1-inte on ind+1)
  v1 = common * (inte * t0.y + inte1 * t1.y);
  if (c.limb != 0) {
    const double2 c0 = __ldg(c.cosne + ind), c1 = __ldg(c.cosne + ind + 1);
    const double m0 = inte * c0.x + inte1 * c1.x, m1 = inte * c0.y + inte1 * c1.y;
    if (c.limb == 1) { v0 *= (1.0 + 2.06 * m0); v1 *= (1.0 + 2.06 * m1); }
    else if (c.limb == 2) { v0 *= log(1.0 + 1.0 / m0); v1 *= log(1.0 + 1.0 / m1); }
  }
}

// (obtprec > prec) of the reference, obtprec = fabs(t_new - t_old) / t_new  (src/Relprofile.cpp:575)
__device__ __forceinline__ bool not_converged(double t_new, double t_old) {
  const double d = fabs(t_new - t_old);
  if (t_new > 0.0) return d > 0.02 * t_new;
  if (t_new == 0.0) return d > 0.0;   // x/0 = inf > prec;  0/0 = NaN compares false
  return false;                        // negative (or NaN) quotient ends the loop
}

__device__ __forceinline__ double gstar2ener(double g, double gmin, double gmax) { return (g * (gmax - gmin) + gmin) * 1.0; }

__device__ double int_edge(double blo, double bhi, const RelbCtx &c) {  // src/Relprofile.cpp:585-621 (h = GFAC_H)
  double hex, lo, hi;
  if (blo <= 0.5) { hex = GFAC_H; lo = blo; hi = bhi; }
  else { hex = 1.0 - GFAC_H; lo = 1.0 - bhi; hi = 1.0 - blo; }
  double n0, n1;
  relb2(gstar2ener(hex, c.gmin, c.gmax), c, n0, n1);
  double norm = 0.0;
  norm = norm + n0;
  norm = norm + n1;
  norm = norm * sqrt(GFAC_H);
  return 2 * norm * (sqrt(hi) - sqrt(lo)) * 1.0 * (c.gmax - c.gmin);
}

// Romberg on [a, b] for both branches, at most two halvings (src/Relprofile.cpp:524-579 with the loop cut
// after niter = 2).  Returns true and the sum of the two integrals if both branches converged.
__device__ bool romberg2_capped(double a, double b, const RelbCtx &c, double &out, double (&fend)[2]) {
  double fa0, fa1, fb0, fb1, m0, m1;
  relb2(a, c, fa0, fa1);
  relb2(b, c, fb0, fb1);
  fend[0] = fa0; fend[1] = fa1;
  const double ta0 = (fa0 + fb0) / 2.0, ta1 = (fa1 + fb1) / 2.0;
  const double pas = b - a;
  const double t00_0 = ta0 * pas, t00_1 = ta1 * pas;
  const double pas1 = pas / 2.0;
  relb2(a + pas1 * 1, c, m0, m1);
  const double t01_0 = (ta0 + m0) * pas1, t01_1 = (ta1 + m1) * pas1;
  const double t10_0 = (4.0 * t01_0 - t00_0) / 3.0, t10_1 = (4.0 * t01_1 - t00_1) / 3.0;
  const bool nc0 = not_converged(t10_0, t00_0), nc1 = not_converged(t10_1, t00_1);
  if (!nc0 && !nc1) { out = t10_0 + t10_1; return true; }
  const double pas2 = pas1 / 2.0;
  double q0, q1, u0, u1;
  relb2(a + pas2 * 1, c, q0, q1);
  relb2(a + pas2 * 3, c, u0, u1);
  double r0 = t10_0, r1 = t10_1;
  bool bad = false;
  if (nc0) {
    const double t02 = (((ta0 + q0) + m0) + u0) * pas2;
    const double t11 = (4.0 * t02 - t01_0) / 3.0;
    const double t20 = (16.0 * t11 - t10_0) / 15.0;
    bad |= not_converged(t20, t10_0);
    r0 = t20;
  }
  if (nc1) {
    const double t02 = (((ta1 + q1) + m1) + u1) * pas2;
    const double t11 = (4.0 * t02 - t01_1) / 3.0;
    const double t20 = (16.0 * t11 - t10_1) / 15.0;
    bad |= not_converged(t20, t10_1);
    r1 = t20;
  }
  out = r0 + r1;
  return !bad;
}

// Richardson step of the Romberg tableau, t[ii] = (4^ii t[ii-1] - tprev[ii-1]) / (4^ii - 1), with the
// divisions replaced by the tabulated reciprocals
__device__ __forceinline__ double richardson(int ii, double cur_lo, double prev_lo) {
  const double r4[7] = {1.0, 4.0, 16.0, 64.0, 256.0, 1024.0, 4096.0};
  const double inv[7] = {0.0, 1.0 / 3.0, 1.0 / 15.0, 1.0 / 63.0, 1.0 / 255.0, 1.0 / 1023.0, 1.0 / 4095.0};
  return (r4[ii] * cur_lo - prev_lo) * inv[ii];
}

// Full-depth Romberg of one bin by a HALF warp (16 lanes; the two halves of a warp work on different bins;
// `active` is uniform per half; all 32 lanes must call).  All lanes of the half return the sum of the two
// branch integrals.  Levels 1..4 use the 17 dyadic points of spacing (b-a)/16 (lane h evaluates point h+1,
// lane 0 also the lower end point), levels 5..6 the 65 points of spacing (b-a)/64.  The level sums come from
// one xor-butterfly per depth: after the steps 8,4,2 the lanes whose index has the same low bits hold the sum
// of their residue class, i.e. exactly the points that are new at one Romberg level.
__device__ double romberg2_half(double a, double b, const RelbCtx &c, bool active, const double *fend) {
  const unsigned FULL = 0xffffffffu;
  const int h = threadIdx.x & 15;          // lane inside the half
  const int base = threadIdx.x & 16;       // first lane of this half inside the warp
  const double pas = b - a;
  double res[2] = {0.0, 0.0};
  bool done[2] = {!active, !active};
  double tprev[2][7], ta[2];
  // ---- depth 4: point p = h + 1 (p = 16 is the upper end point), lane 0 additionally p = 0
  const double pas4 = pas / 16.0;
  double v[2] = {0.0, 0.0};
  if (active) relb2(h == 15 ? b : a + pas4 * (h + 1), c, v[0], v[1]);   // f(a) comes from phase 1
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const double fa = active ? fend[k] : 0.0, fb = __shfl_sync(FULL, v[k], base + 15);
    // class sums over the interior points p = 1..15 (lane h = p - 1): odd p, p = 2 mod 4, p = 4 mod 8, p = 8
    double x = (h == 15) ? 0.0 : v[k];
    const double n1 = __shfl_sync(FULL, x, base + 7);                 // p = 8
    x += __shfl_xor_sync(FULL, x, 8);
    const double n2 = __shfl_sync(FULL, x, base + 3);                 // p = 4, 12
    x += __shfl_xor_sync(FULL, x, 4);
    const double n3 = __shfl_sync(FULL, x, base + 1);                 // p = 2, 6, 10, 14
    x += __shfl_xor_sync(FULL, x, 2);
    const double n4 = __shfl_sync(FULL, x, base + 0);                 // odd p
    if (done[k]) continue;
    ta[k] = (fa + fb) / 2.0;
    tprev[k][0] = ta[k] * pas;
    double last = tprev[k][0], pasn = pas, sum = ta[k];
    const double newp[5] = {0.0, n1, n2, n3, n4};
#pragma unroll
    for (int n = 1; n <= 4; n++) {
      pasn = pasn * 0.5;
      sum += newp[n];
      if (!done[k]) {
        double cur[7];
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
  // ---- depth 6 (rare): 63 interior points, lane h takes p = h + 1 + 16 q, q = 0..3 (p = 64 is the end point)
  const bool need6 = !(done[0] && done[1]);
  if (__any_sync(FULL, need6)) {
    const double pas6 = pas / 64.0;
    double w0[4], w1[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      w0[q] = 0.0; w1[q] = 0.0;
      const int pidx = h + 1 + 16 * q;
      if (need6 && pidx < 64) relb2(a + pas6 * pidx, c, w0[q], w1[q]);
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
      // new points of level 6: odd p; of level 5: p = 2 mod 4 (the others were used by levels <= 4)
      double o6 = 0.0, o5 = 0.0;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int pidx = h + 1 + 16 * q;
        const double val = k ? w1[q] : w0[q];
        if (pidx < 64) {
          if (pidx & 1) o6 += val;
          else if ((pidx & 3) == 2) o5 += val;
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        o6 += __shfl_xor_sync(FULL, o6, o);
        o5 += __shfl_xor_sync(FULL, o5, o);
      }
      if (done[k]) continue;
      // the level-4 trapezoid sum is recovered from the tableau row: cur[0] = sum * pasn
      double pasn = pas / 16.0;
      double sum = tprev[k][0] / pasn;
      double last = res[k];
      const double newp[2] = {o5, o6};
#pragma unroll
      for (int n = 5; n <= 6; n++) {
        pasn = pasn * 0.5;
        sum += newp[n - 5];
        if (!done[k]) {
          double cur[7];
          cur[0] = sum * pasn;
#pragma unroll
          for (int ii = 1; ii <= 6; ii++) if (ii <= n) cur[ii] = richardson(ii, cur[ii - 1], tprev[k][ii - 1]);
          res[k] = cur[n];
          if (!not_converged(cur[n], last)) done[k] = true;
          last = cur[n];
#pragma unroll
          for (int ii = 0; ii <= 6; ii++) if (ii <= n) tprev[k][ii] = cur[ii];
        }
      }
    }
  }
  return res[0] + res[1];
}

// integ_relline_bin (src/Relprofile.cpp:650-726) without the deep Romberg levels: returns the bin integral,
// or (deferred = true) the edge terms only, with [ra, rb] the interval still to be integrated.
__device__ double integ_bin_phase1(const RelbCtx &c, double rlo0, double rhi0, bool &deferred, double &ra, double &rb, double (&fend)[2]) {
  deferred = false;
  double flu = 0.0;
  double gblo = (rlo0 / 1.0 - c.gmin) * c.del_g;
  if (gblo < 0.0) gblo = 0.0; else if (gblo > 1.0) gblo = 1.0;
  double gbhi = (rhi0 / 1.0 - c.gmin) * c.del_g;
  if (gbhi < 0.0) gbhi = 0.0; else if (gbhi > 1.0) gbhi = 1.0;
  if (gbhi == 0) return 0.0;
  double rlo = rlo0, rhi = rhi0, hlo, hhi;
  if (gblo <= GFAC_H) {
    hlo = gblo;
    hhi = GFAC_H;
    rlo = gstar2ener(GFAC_H, c.gmin, c.gmax);
    if (gbhi <= GFAC_H) { hhi = gbhi; rlo = -1.0; }
    flu = flu + int_edge(hlo, hhi, c);
  }
  if (gbhi >= (1.0 - GFAC_H)) {
    hhi = gbhi;
    hlo = 1.0 - GFAC_H;
    rhi = gstar2ener(1 - GFAC_H, c.gmin, c.gmax);
    if (gblo >= (1.0 - GFAC_H)) { hlo = gblo; rhi = -1.0; }
    flu = flu + int_edge(hlo, hhi, c);
  }
  if ((rhi >= 0) && (rlo >= 0)) {
    if (rlo >= 1.0 * 0.95) {  // src/Relprofile.cpp:628-647
      double f2;
      if (romberg2_capped(rlo, rhi, c, f2, fend)) flu = flu + f2;
      else { deferred = true; ra = rlo; rb = rhi; }
    } else {
      double m0, m1;
      relb2((rhi + rlo) / 2.0, c, m0, m1);
      double f2 = 0.0;
      f2 += m0 * (rhi - rlo);
      f2 += m1 * (rhi - rlo);
      flu = flu + f2;
    }
  }
  return flu;
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
// binary_search(ener, n+1, val) of the reference on the (possibly rescaled) grid
__device__ int line_bsearch(const double *egrid, int n_edges, double val, int grid_mode, double z, double lineE) {
  int klo = 0, khi = n_edges - 1;
  while (khi - klo > 1) {
    const int k = (khi + klo) >> 1;
    if (line_edge(egrid, k, grid_mode, z, lineE) > val) khi = k; else klo = k;
  }
  return klo;
}

constexpr int LN_NT = 256;
constexpr int LN_BUF = 2048;   // contribution slots per sub-batch
constexpr int LN_MAXR = 64;    // radii per sub-batch
struct LnRad {
  double gmin, gmax, del_g, scale, weight;
  int ielo, iehi, off, gi;
};
struct LnSmem {
  double contrib[LN_BUF];
  double def_a[LN_BUF / 4], def_b[LN_BUF / 4];   // work list of phase 2: interval still to integrate
  double def_f[LN_BUF / 4][2];                    // ... and the integrand at its lower end (both branches)
  LnRad rad[LN_MAXR + 1];
  unsigned short def_item[LN_BUF / 4];
  unsigned char item_rad[LN_BUF];                 // sub-batch radius of every item
  int nrad, ndef, overflow, cursor, jlo, jhi, resume;   // resume: first bin still to do of radius `cursor` (-1 = all)
  int zjlo, zjhi;
};
constexpr int LN_MAXDEF = LN_BUF / 4;

__device__ __forceinline__ void ln_ctx(const LnRad &lr, const double2 *g_trff, const double2 *g_cosne, int limb, RelbCtx &c) {
  c.gmin = lr.gmin; c.gmax = lr.gmax; c.del_g = lr.del_g; c.scale = lr.scale;
  c.trff = g_trff + (size_t) lr.gi * NG; c.cosne = g_cosne + (size_t) lr.gi * NG; c.limb = limb;
}

__global__ void __launch_bounds__(LN_NT, 3) k_line(const VPar *__restrict__ vps, DevTables T, Scratch S,
                                                   const double *__restrict__ egrid, int n_ener, int grid_mode,
                                                   int ne_stride, int nz_stride, int n_acc) {
  extern __shared__ __align__(16) unsigned char smraw[];
  LnSmem &sm = *reinterpret_cast<LnSmem *>(smraw);
  double *acc = reinterpret_cast<double *>(smraw + sizeof(LnSmem));   // [n_acc]
  const int v = blockIdx.y, z = blockIdx.x, t = threadIdx.x;
  if (S.status[v] != ST_OK) return;
  const VPar &vp = vps[v];
  if (z >= vp.nz) return;
  const double zred = vp.z, lineE = vp.lineE;
  const int limb = vp.limb;
  const double e_first = line_edge(egrid, 0, grid_mode, zred, lineE);
  const double e_last = line_edge(egrid, n_ener, grid_mode, zred, lineE);
  // radii of this zone (izone[] is non-increasing along the descending-radius fine grid; k_syspar tabulated
  // the first index of every zone)
  const int ia = S.zfirst[(size_t) v * (NZMAX + 1) + z + 1], ib = S.zfirst[(size_t) v * (NZMAX + 1) + z];
  double *flux = S.relflux + ((size_t) v * nz_stride + z) * ne_stride;
  const double *g_re = S.re + (size_t) v * NR;
  const double2 *g_trff = reinterpret_cast<const double2 *>(S.trff) + (size_t) v * NR * NG;
  const double2 *g_cosne = reinterpret_cast<const double2 *>(S.cosne) + (size_t) v * NR * NG;
  for (int j = t; j < n_acc; j += LN_NT) acc[j] = 0.0;
  if (t == 0) { sm.cursor = ia; sm.resume = -1; sm.zjlo = n_ener; sm.zjhi = -1; }
  __syncthreads();

  while (true) {
    const int cur = sm.cursor;
    if (cur >= ib) break;
    const int resume = sm.resume;
    // ---- sub-batch set-up: one thread per radius
    if (t < LN_MAXR) {
      const int i = cur + t;
      LnRad lr;
      lr.gi = i;
      lr.ielo = 0;
      lr.iehi = -1;
      if (i < ib) {
        lr.gmin = S.gmin[(size_t) v * NR + i];
        lr.gmax = S.gmax[(size_t) v * NR + i];
        lr.del_g = 1. / (lr.gmax - lr.gmin);
        lr.scale = lr.del_g * S.emis[(size_t) v * NR + i];
        lr.weight = trapez_single(g_re, i, NR) / 2;
        if ((lr.gmax > e_first) && (lr.gmin < e_last)) {  // src/Relprofile.cpp:863-878
          double egmin = lr.gmin, egmax = lr.gmax;
          if (egmin < e_first) egmin = e_first;
          if (egmax > e_last) egmax = e_last;
          lr.ielo = line_bsearch(egrid, n_ener + 1, egmin, grid_mode, zred, lineE);
          lr.iehi = line_bsearch(egrid, n_ener + 1, egmax, grid_mode, zred, lineE);
          if (t == 0 && resume >= 0) lr.ielo = resume;   // rest of a radius wider than the buffer
        }
      }
      sm.rad[t] = lr;
    }
    __syncthreads();
    if (t == 0) {
      int off = 0, n = 0, jlo = n_ener, jhi = -1, next_resume = -1;
      while (n < LN_MAXR && cur + n < ib) {
        int w = sm.rad[n].iehi - sm.rad[n].ielo + 1;
        if (off + w > LN_BUF) {
          if (n > 0) break;
          w = LN_BUF;                                   // a single radius wider than the buffer: take a piece
          next_resume = sm.rad[n].ielo + w;
          sm.rad[n].iehi = next_resume - 1;
        }
        sm.rad[n].off = off;
        off += (w > 0 ? w : 0);
        if (w > 0) { jlo = min(jlo, sm.rad[n].ielo); jhi = max(jhi, sm.rad[n].iehi); }
        n++;
        if (next_resume >= 0) break;
      }
      sm.rad[n].off = off;
      sm.nrad = n;
      sm.ndef = 0;
      sm.overflow = 0;
      sm.jlo = jlo;
      sm.jhi = jhi;
      sm.zjlo = min(sm.zjlo, jlo);
      sm.zjhi = max(sm.zjhi, jhi);
      sm.resume = next_resume;
      sm.cursor = (next_resume >= 0) ? cur + n - 1 : cur + n;
    }
    __syncthreads();
    const int nrad = sm.nrad;
    const int nitems = sm.rad[nrad].off;
    for (int r = t >> 5; r < nrad; r += LN_NT / 32)   // item -> radius map, one warp per radius
      for (int q = sm.rad[r].off + (t & 31); q < sm.rad[r + 1].off; q += 32) sm.item_rad[q] = (unsigned char) r;
    __syncthreads();
    // ---- phase 1: one thread per item.  A bin whose Romberg integral needs more than two halvings goes on
    // the work list; if the list is full the item is flagged (bit 7 of item_rad) and retried after phase 2
    // has drained the list, so the result never depends on the order in which threads reach the list.
    for (int round = 0;; round++) {
      for (int item = t; item < nitems; item += LN_NT) {
        const int ir = sm.item_rad[item];
        if (round > 0 && !(ir & 0x80)) continue;
        const LnRad &lr = sm.rad[ir & 0x7f];
        const int j = lr.ielo + (item - lr.off);
        RelbCtx c;
        ln_ctx(lr, g_trff, g_cosne, limb, c);
        const double elo = line_edge(egrid, j, grid_mode, zred, lineE), ehi = line_edge(egrid, j + 1, grid_mode, zred, lineE);
        bool deferred;
        double ra, rb, fend[2];
        const double val = integ_bin_phase1(c, elo, ehi, deferred, ra, rb, fend);
        sm.item_rad[item] = (unsigned char) (ir & 0x7f);
        if (deferred) {
          const int d = atomicAdd(&sm.ndef, 1);
          if (d < LN_MAXDEF) {
            sm.def_item[d] = (unsigned short) item;
            sm.def_a[d] = ra;
            sm.def_b[d] = rb;
            sm.def_f[d][0] = fend[0];
            sm.def_f[d][1] = fend[1];
          } else {
            sm.item_rad[item] = (unsigned char) (ir | 0x80);
            sm.overflow = 1;
          }
        }
        sm.contrib[item] = val;
      }
      __syncthreads();
      // ---- phase 2: one half warp per listed bin
      {
        const int ndef = min(sm.ndef, LN_MAXDEF);
        const int half = t >> 4, hl = t & 15;
        const int nloop = (ndef + LN_NT / 16 - 1) / (LN_NT / 16);
        for (int it = 0; it < nloop; it++) {
          const int d = it * (LN_NT / 16) + half;
          const bool active = d < ndef;
          RelbCtx c;
          double ra = 0.0, rb = 1.0;
          int item = 0;
          if (active) {
            item = sm.def_item[d];
            ln_ctx(sm.rad[sm.item_rad[item] & 0x7f], g_trff, g_cosne, limb, c);
            ra = sm.def_a[d];
            rb = sm.def_b[d];
          } else {
            ln_ctx(sm.rad[0], g_trff, g_cosne, limb, c);
          }
          const double f2 = romberg2_half(ra, rb, c, active, sm.def_f[active ? d : 0]);
          if (active && hl == 0) sm.contrib[item] = sm.contrib[item] + f2;
        }
      }
      const int again = sm.overflow;
      __syncthreads();
      if (!again) break;
      if (t == 0) { sm.ndef = 0; sm.overflow = 0; }
      __syncthreads();
    }
    // ---- phase 3: ordered accumulation (ascending radius index = the reference's loop order)
    for (int j = sm.jlo + t; j <= sm.jhi; j += LN_NT) {
      double a = acc[j];
      for (int r = 0; r < nrad; r++) {
        const LnRad &lr = sm.rad[r];
        if (j >= lr.ielo && j <= lr.iehi) a += sm.contrib[lr.off + (j - lr.ielo)] * lr.weight;
      }
      acc[j] = a;
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
    flux[j] = acc[j] / (0.5 * (elo + ehi));
  }
}

// ---------------------------------------------------------------------------------- launcher
int line_kernel_init() {
  cudaError_t e = cudaFuncSetAttribute(k_line, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (200 * 1024));
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(k_line, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared);
  return e == cudaSuccess ? 0 : 1;
}

void launch_line(const VPar *vps, const DevTables &T, const Scratch &S, long n, const double *egrid, int n_ener,
                 int grid_mode, int nz_max, cudaStream_t st) {
  dim3 grid(nz_max, (unsigned) n);
  const int n_acc = ((n_ener + 31) / 32) * 32;
  const size_t sm = sizeof(LnSmem) + (size_t) n_acc * sizeof(double);
  k_line<<<grid, LN_NT, sm, st>>>(vps, T, S, egrid, n_ener, grid_mode, S.ne_line_cap, S.nz_cap, n_acc);
}
int line_max_bins() { return (int) ((200 * 1024 - sizeof(LnSmem)) / sizeof(double)); }

}  // namespace rx
