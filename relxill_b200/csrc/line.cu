// line.cu — the relline profile kernel (the FP64 hot loop of the relativistic smearing).
// Compiled with FMA contraction on (build.py): its results feed no discrete decision other than the
// Romberg convergence test, which has a 2 % threshold.
//
// Replaces calc_relline_profile + integ_relline_bin + int_edge + int_romb + romberg_integration +
// relb_func (src/Relprofile.cpp:489-726,835-905), the radial half of interpol_relTable (:280-293) and the division
// by the bin energy of renorm_relline_profile (:757-762).
//
// Mapping.  One CTA of ONE WARP per (vector, radial zone[, run of the zone's radii]); no block barrier anywhere.  The
// warp takes the radii in sub-batches of <= LN_R = 8:
//   staging  the (a, mu0)-interpolated transfer-function rows of the TABLE radii that bracket the sub-batch (k_rows'
//            output, contiguous) come in with ONE bulk asynchronous copy (cp.async.bulk + mbarrier: the TMA unit moves
//            them while the lanes compute bin ranges); the radial interpolation onto the sub-batch's fine radii is
//            then done from shared memory into shared memory.  The fine transfer functions never exist in HBM.
//   set-up   lane = (radius, task): first bin | last bin (closed-form index on the logarithmic convolution grid,
//            corrected against the tabulated edges) | the integrand at the edge nodes g* = h | g* = 1-h that
//            int_edge needs (once per radius instead of once per edge bin).
//   main     BIN-STATIONARY.  The sub-batch's bin range is cut into tiles of 15 bins anchored at the bin where the
//            quadrature rule changes (E = 0.95, src/Relprofile.cpp:633), so a tile is all midpoint-rule or all
//            Romberg.  The two half-warps work on two consecutive radii at a time, lane = bin EDGE (16 edges = 15
//            bins), and walk the radii in ascending order with the bin's sum in a register: no atomics, no
//            contribution buffer.
//              midpoint tiles: one evaluation of the integrand per bin.
//              Romberg tiles:  the integrand at the bin edges is evaluated once per edge and shared by the two
//                              neighbouring bins through a shuffle (the reference evaluates it twice); levels 1-2
//                              (three more abscissae) for all lanes.  The bins that have not converged by then (the
//                              horns, ~13 %) are QUEUED in shared memory with their tableau (the queue lives in the
//                              row stage, dead by then) and worked off once per sub-batch, or when the queue is full:
//                              levels 3 and 4 one lane per queued bin, what is left after level 4 (~1e-4 of the
//                              bins) by a whole new integration, one lane each (deep_flush).
//            The zone's row in HBM is the accumulator between sub-batches (read once, written once per tile and
//            sub-batch); the queued bins' results are added to it when the queue is worked off, so a bin's sum runs in
//            ascending radius order except for those late terms: deterministic, a function of the vector alone.
//   split    vectors with few zones (line and convolution models, one-zone relxill flavours) have every zone cut into
//            line_parts(nz) runs of radii, one CTA each (a single XSPEC call of relline would otherwise be ONE warp
//            walking 1000 radii); k_linemerge adds the partial rows in order.
// What the measurements of round 2 taught (profiles/README.md): with the integrand inlined at every quadrature site
// the kernel was bound by instruction fetch (independent warps, 90 KB of code: stall_no_instruction 3.5 per issue), so
// everything outside the two tile loops is written for size (one out-of-line copy of the integrand, rolled loops);
// Romberg state kept in registers across the deep levels spilled (local memory through a 28 KB L1), hence the queue.
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
#include <cstdint>
#include <cstdlib>

#include "common.h"
#include "devutil.cuh"
#include "kernels.h"

namespace rx {

#define GS_C ((1.0 - 2 * GFAC_H) / (NG - 1))
#define GS_INVC ((NG - 1) / (1.0 - 2 * GFAC_H))
// An FP64 immediate whose low word is not zero costs two moves at every use; as __constant__ data the same values
// ride in the instruction as constant-bank operands.
struct LnConst {
  double h, one_m_h, gs_c, gs_invc, prec, e95, sqrt_h;
  double inv[7];   // 1 / (4^ii - 1) of the Richardson step
};
__constant__ LnConst LK = {GFAC_H, 1.0 - GFAC_H, GS_C, GS_INVC, 0.02, 1.0 * 0.95, 0.0,
                           {0.0, 1.0 / 3.0, 1.0 / 15.0, 1.0 / 63.0, 1.0 / 255.0, 1.0 / 1023.0, 1.0 / 4095.0}};

constexpr int LN_NT = 32;      // one warp per CTA
constexpr int LN_MINB = 20;     // resident CTAs per SM asked of the compiler: 96 registers (22 CTAs / 80 registers: spills, 9.2 ms against 8.6)
constexpr int LN_R = 8;          // radii per sub-batch
constexpr int LN_ROWS = 5;       // table rows staged per sub-batch (a sub-batch is cut where its bracket would not fit)
constexpr int LN_TB = 15;        // bins per tile (16 edges: one half warp)
constexpr int LN_FS = NG + 1;    // row stride of the fine rows in shared memory (padded: consecutive radii on different banks)

struct LnRad {
  double gmin, del_g, dgm;       // dgm = gmax - gmin
  double scale;                  // emis / (gmax - gmin)
  double weight;
  double nlo, nhi;               // `norm` of int_edge (src/Relprofile.cpp:585-621) at g* = h and g* = 1-h
  double ehlo, ehhi;             // gstar2ener(h), gstar2ener(1-h)
  int ielo, iehi, gi, pad;
};
// A Romberg bin that has not converged after level 2 waits here for levels 3+ (deep_flush)
struct LnDeep {
  double a, pas;                 // the bin's quadrature interval [a, a + pas]
  double sum[2];                 // trapezoid sums (in units of the level's step) per branch; at the end sum[0] = the bin's result
  double tq[2][3];               // tableau row without its first entry (= sum * step); a finished branch: [0] = its result
  int rsel, j;                   // radius of the sub-batch, energy bin
  int flags, pad;                // bits 0-1: branch finished; bits 8-..: level still to do (1: all of it, 3, 4, 5; 0: done)
};
// the queue lives in the row stage, dead in the main loop; at most one entry per lane (deep_flush)
constexpr int LN_DQ = ((LN_ROWS * NG * 16) / (int) sizeof(LnDeep)) < 32 ? ((LN_ROWS * NG * 16) / (int) sizeof(LnDeep)) : 32;
struct LnSmem {
  union {
    double2 rows[LN_ROWS][NG];   // bulk-copy destination
    LnDeep dq[LN_DQ];
  };
  double2 fine[LN_R][LN_FS];     // {branch 0, branch 1} of the sub-batch's radii
  LnRad rad[LN_R + 2];
  unsigned long long mbar;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 lds_d2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
// ---- bulk asynchronous copy (TMA unit), completion on an mbarrier
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy reads of dst are done (after a barrier)
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t phase) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(a), "r"(phase) : "memory");
  } while (!done);
}

struct RelbCtx {
  double gmin, del_g;
  double scale;          // emis / (gmax - gmin)
  uint32_t row;          // shared-memory address of this radius' fine row: [NG] {branch 0, branch 1}
  const double2 *cosne;  // global; only for limb != 0
  int limb;
};

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

// both branches of relb_func (src/Relprofile.cpp:489-521) at energy eg.  ONE out-of-line copy, everything passed in
// registers: the kernel's warps walk their zones independently, and with the integrand inlined at every quadrature
// site the instruction cache was what bounded the kernel (ncu: stall_no_instruction 3.5 per issue).
__device__ __forceinline__ double2 relb2_fi(double eg, double gmin, double del_g, double scale, uint32_t row, int limb,
                                        const double2 *cosne) {
  const double egstar = (eg - gmin) * del_g;
  int ind = (int) ((egstar - LK.h) * LK.gs_invc);
  ind = ind < 0 ? 0 : (ind > NG - 2 ? NG - 2 : ind);
  const double inte = (egstar - (LK.h + LK.gs_c * (double) ind)) * LK.gs_invc;
  const double inte1 = 1.0 - inte;
  const double2 t0 = lds_d2(row + ind * 16), t1 = lds_d2(row + ind * 16 + 16);
  const double common = (eg * eg * eg) * rsqrt(egstar - egstar * egstar) * scale;
  double2 v;
  v.x = common * (inte * t0.x + inte1 * t1.x);   // (the reference's weights: inte on node ind, 1-inte on ind+1)
  v.y = common * (inte * t0.y + inte1 * t1.y);
  if (limb != 0) {
    const double2 f = relb2_limb(ind, inte, cosne, limb);
    v.x *= f.x;
    v.y *= f.y;
  }
  return v;
}
// the out-of-line copy (set-up, deep levels, whole integrations)
__device__ __noinline__ double2 relb2_f(double eg, double gmin, double del_g, double scale, uint32_t row, int limb,
                                        const double2 *cosne) {
  return relb2_fi(eg, gmin, del_g, scale, row, limb, cosne);
}
__device__ __forceinline__ void relb2(double eg, const RelbCtx &c, double &v0, double &v1) {
  const double2 v = relb2_f(eg, c.gmin, c.del_g, c.scale, c.row, c.limb, c.cosne);
  v0 = v.x;
  v1 = v.y;
}
// inlined: the quadrature loops of the main loop (independent evaluations overlap)
__device__ __forceinline__ void relb2i(double eg, const RelbCtx &c, double &v0, double &v1) {
  const double2 v = relb2_fi(eg, c.gmin, c.del_g, c.scale, c.row, c.limb, c.cosne);
  v0 = v.x;
  v1 = v.y;
}

// (obtprec > prec) of the reference, obtprec = fabs(t_new - t_old) / t_new  (src/Relprofile.cpp:575)
__device__ __forceinline__ bool not_converged(double t_new, double t_old) {
  const double d = fabs(t_new - t_old);
  if (t_new > 0.0) return d > LK.prec * t_new;
  if (t_new == 0.0) return d > 0.0;   // x/0 = inf > prec;  0/0 = NaN compares false
  return false;                        // negative (or NaN) quotient ends the loop
}

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
__device__ __noinline__ int line_index(const LineGrid &G, double val, double z, double lineE) {
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

// int_edge (src/Relprofile.cpp:585-621) with the integrand at the edge node already summed into `norm`
__device__ __forceinline__ double edge_term(double blo, double bhi, double norm, double dgm) {
  double lo, hi;
  if (blo <= 0.5) { lo = blo; hi = bhi; }
  else { lo = 1.0 - bhi; hi = 1.0 - blo; }
  return 2 * norm * (sqrt(hi) - sqrt(lo)) * 1.0 * dgm;
}

// Full Romberg integration of [a, a + pas] by one lane (src/Relprofile.cpp:524-579, both branches sharing the abscissae).
// Out of line and rare: (i) a bin that has not converged after the cooperative level 4, ~1e-4 of the Romberg bins;
// (ii) the bin below E = 0.95 whose lower limit is raised to the analytic edge interval's end above 0.95 (a few per
// vector).  The new abscissae of a level are summed in ascending order.
__device__ __noinline__ double romberg_bin(double a, double pas, RelbCtx c) {
  double fa0, fa1, fb0, fb1;
  relb2(a, c, fa0, fa1);
  relb2(a + pas, c, fb0, fb1);
  double tprev[2][7], res[2], sum[2] = {(fa0 + fb0) / 2.0, (fa1 + fb1) / 2.0};
  bool done[2] = {false, false};
#pragma unroll 1
  for (int k = 0; k < 2; k++) {
    tprev[k][0] = sum[k] * pas;
    res[k] = tprev[k][0];
  }
#pragma unroll 1
  for (int n = 1; n <= 6; n++) {
    if (done[0] && done[1]) break;
    const double pasn = pas * (1.0 / (double) (1 << n));
    double o[2] = {0.0, 0.0};
#pragma unroll 1
    for (int p = 1; p < (1 << n); p += 2) {
      double w0, w1;
      relb2(a + pasn * p, c, w0, w1);
      o[0] += w0;
      o[1] += w1;
    }
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
      sum[k] += o[k];
      if (done[k]) continue;
      double prev = tprev[k][0], cur = sum[k] * pasn, r4 = 1.0;
      tprev[k][0] = cur;
#pragma unroll 1
      for (int ii = 1; ii <= n; ii++) {   // t[ii] = (4^ii t[ii-1] - tprev[ii-1]) / (4^ii - 1)
        r4 *= 4.0;
        cur = (r4 * cur - prev) * LK.inv[ii];
        prev = tprev[k][ii];
        tprev[k][ii] = cur;
      }
      if (!not_converged(cur, res[k])) done[k] = true;
      res[k] = cur;
    }
  }
  double r = 0.0;
  r += res[0];
  r += res[1];
  return r;
}

__device__ __forceinline__ void ln_ctx(const LnSmem &sm, int r, const double2 *g_cosne, int limb, RelbCtx &c) {
  const LnRad &lr = sm.rad[r];
  c.gmin = lr.gmin; c.del_g = lr.del_g; c.scale = lr.scale;
  c.row = smem_u32(&sm.fine[r][0]);
  c.cosne = g_cosne + (size_t) lr.gi * NG; c.limb = limb;
}

// Levels 3+ of the queued bins (src/Relprofile.cpp:553-576), one lane per queued bin: the level's new abscissae, then
// the bin's tableau in the queue entry.  (Level 4 compacted over the warp, 8 lanes per bin with an xor-butterfly, was
// slower: 8.47 against 8.36 ms.)  What is left after level 4 (~1e-4 of the Romberg bins) and the bins queued for a whole
// integration are finished by one lane each too.  The results (weighted with
// the radius' area weight) are added to the zone's row in HBM; queue order = radius order, so the sum is deterministic.
// Written for size, not speed (rolled loops, one call site of the integrand): see relb2_f.
__device__ __forceinline__ void deep_tableau(LnDeep &d, int L, double s0, double s1) {
  const double pasn = d.pas * (1.0 / (double) (1 << L)), pasp = pasn * 2.0;
  int flags = d.flags & 3;
  bool more = false;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    double prev = d.sum[k] * pasp;   // first tableau entry of the previous level
    const double sum = d.sum[k] + (k ? s1 : s0);
    d.sum[k] = sum;
    if (flags & (1 << k)) continue;
    double cur = sum * pasn, r4 = 1.0, last = 0.0;
#pragma unroll
    for (int ii = 1; ii <= 4; ii++) {
      if (ii > L) break;   // t[ii] = (4^ii t[ii-1] - tprev[ii-1]) / (4^ii - 1); tq[ii-1] holds t[ii]
      r4 *= 4.0;
      cur = (r4 * cur - prev) * LK.inv[ii];
      if (ii < L) { prev = d.tq[k][ii - 1]; d.tq[k][ii - 1] = cur; last = prev; }
    }
    if (!not_converged(cur, last)) {
      flags |= 1 << k;
      d.tq[k][0] = cur;
    } else {
      if (L == 3) d.tq[k][2] = cur;
      more = true;
    }
  }
  // not converged after level 4 (~1e-4 of the Romberg bins): the bin is integrated again as a whole by one lane
  d.flags = flags | ((more ? (L == 3 ? 4 : 1) : 0) << 8);
}
__device__ __forceinline__ void deep_flush(LnSmem &sm, int ndq, int lane, const double2 *g_cosne, int limb, double *flux) {
  const unsigned FULL = 0xffffffffu;
  // the bins' sums so far (the tiles they belong to have been written): fetched now, needed at the very end
  const double before = (lane < ndq) ? flux[sm.dq[lane].j] : 0.0;
  // ---- levels 3 and 4, one lane per entry: the level's new abscissae a + p pas / 2^L (p odd) in ascending order,
  // then the tableau (src/Relprofile.cpp:553-576)
#pragma unroll 1
  for (int L = 3; L <= 4; L++) {
    LnDeep &d = sm.dq[min(lane, ndq - 1)];
    const bool go = (lane < ndq) && ((d.flags >> 8) == L);
    if (!__any_sync(FULL, go)) break;   // nothing goes on
    RelbCtx c;
    ln_ctx(sm, d.rsel, g_cosne, limb, c);
    const double a = d.a, pasl = d.pas * (L == 3 ? 0.125 : 0.0625);
    double s0 = 0.0, s1 = 0.0;
#pragma unroll 1
    for (int p = 1; p < (1 << L); p += 2) {
      double w0, w1;
      relb2(a + pasl * (double) p, c, w0, w1);
      s0 += w0;
      s1 += w1;
    }
    if (go) { if (L == 3) deep_tableau(d, 3, s0, s1); else deep_tableau(d, 4, s0, s1); }
    __syncwarp();
  }
  // ---- results, one lane per entry (whole integrations, level code 1, are done here), added to the zone's row in
  // HBM: the tiles the entries come from have been written.  Entries of the same bin (different radii) are summed in
  // queue order by the first of them.
  double cwt = 0.0;
  int j = -1 - lane;
  if (lane < ndq) {
    LnDeep &d = sm.dq[lane];
    j = d.j;
    double r;
    if ((d.flags >> 8) == 1) {
      RelbCtx c;
      ln_ctx(sm, d.rsel, g_cosne, limb, c);
      r = romberg_bin(d.a, d.pas, c);
    } else {
      r = 0.0;
      r += d.tq[0][0];
      r += d.tq[1][0];
    }
    cwt = r * sm.rad[d.rsel].weight;
    d.sum[0] = cwt;
  }
  const unsigned same = __match_any_sync(FULL, j);
  __syncwarp();
  if (lane < ndq && lane == __ffs(same) - 1) {
    double tot = 0.0;
#pragma unroll 1
    for (unsigned mm = same; mm; mm &= mm - 1) tot += sm.dq[__ffs(mm) - 1].sum[0];   // ascending lanes = queue order
    flux[j] = before + tot;
  }
  __syncwarp();
}

// analytic edge terms of a bin that reaches into [0, h] or [1-h, 1] (integ_relline_bin, src/Relprofile.cpp:650-726;
// the decisions in g* like the reference's)
__device__ __noinline__ double edge_terms(double Ea, double Eb, double gmin, double del_g, double dgm, double nlo, double nhi) {
  double ga = (Ea / 1.0 - gmin) * del_g;
  if (ga < 0.0) ga = 0.0; else if (ga > 1.0) ga = 1.0;
  double gb = (Eb / 1.0 - gmin) * del_g;
  if (gb < 0.0) gb = 0.0; else if (gb > 1.0) gb = 1.0;
  double flu = 0.0;
  if (gb == 0) return flu;
  if (ga <= LK.h) flu = flu + edge_term(ga, (gb <= LK.h) ? gb : LK.h, nlo, dgm);   // lower-edge term first, like the reference
  if (gb >= LK.one_m_h) flu = flu + edge_term((ga >= LK.one_m_h) ? ga : LK.one_m_h, gb, nhi, dgm);
  return flu;
}

// One warp per CTA: no block barrier anywhere (the only synchronisation is the mbarrier of the bulk copy and
// __syncwarp around the shared-memory stage).  24 resident CTAs per SM at 8.8 KB of shared memory and 80 registers.
template <int GRID_MODE>
__global__ void __launch_bounds__(LN_NT, LN_MINB) k_line(const VPar *__restrict__ vps, DevTables T, Scratch S, LineGrid G,
                                                   int ne_stride, int nz_stride) {
  __shared__ __align__(128) LnSmem sm;
  const unsigned FULL = 0xffffffffu;
  const int v = blockIdx.x, z = blockIdx.y, lane = threadIdx.x;
  if (S.status[v] != ST_OK) return;
  if (S.reuse && S.reuse[v]) return;   // the zone profiles of the previous run stand
  const VPar &vp = vps[v];
  if (z >= vp.nz) return;
  const double *egrid = G.e;
  const int n_ener = G.n_ener;
  constexpr int grid_mode = GRID_MODE;
  const double zred = vp.z, lineE = vp.lineE;
  const int limb = vp.limb;
  // radii of this zone (izone[] is non-increasing along the descending-radius fine grid; k_syspar tabulated
  // the first index of every zone)
  int ia = S.zfirst[(size_t) v * (NZMAX + 1) + z + 1], ib = S.zfirst[(size_t) v * (NZMAX + 1) + z];
  // Vectors with few zones (line and convolution models, the one-zone relxill flavours): the zone's radii are cut into
  // line_parts(nz) runs, one CTA each, whose partial rows k_linemerge adds in ascending order; the division by the bin
  // energy is left to it.  The cut depends on the vector's own zone count only, so a spectrum does not depend on the
  // batch it is evaluated in.
  const int nsplit = line_parts(vp.nz);
  if ((int) blockIdx.z >= nsplit) return;
  const bool whole = nsplit == 1;
  const int row = whole ? z : z * nsplit + (int) blockIdx.z;
  if (!whole) {
    const int len = ((ib - ia + nsplit - 1) / nsplit + LN_R - 1) / LN_R * LN_R;
    ia = min(ib, ia + (int) blockIdx.z * len);
    ib = min(ib, ia + len);
  }
  double *flux = S.relflux + ((size_t) v * nz_stride + row) * ne_stride;   // doubles as the accumulator
  const int *g_it = S.it + (size_t) v * NR;
  const double2 *g_rows = reinterpret_cast<const double2 *>(S.relrow) + (size_t) v * REL_NRT * NG * 2;   // trff plane
  const double2 *g_cosne = reinterpret_cast<const double2 *>(S.cosne) + (size_t) v * NR * NG;
  int j95 = 0;
  if (lane == 0) {
    mbar_init(&sm.mbar, 1);
    // first bin of the Romberg rule: bins j >= j95 have E_lo >= 0.95 (src/Relprofile.cpp:633); tiles are anchored there
    const int k = line_index<GRID_MODE>(G, LK.e95, zred, lineE);
    j95 = (line_edge(egrid, k, grid_mode, zred, lineE) >= LK.e95) ? k : k + 1;
  }
  j95 = __shfl_sync(FULL, j95, 0);
  int zlo = n_ener, zhi = -1;   // bins of the zone's row written so far (the same in every lane)

  uint32_t phase = 0;
  for (int cur = ia; cur < ib;) {
    // ---- the sub-batch: as many radii as fit LN_R and whose table bracket fits LN_ROWS
    int n, it0;
    {
      const int i = min(cur + (lane & (LN_R - 1)), ib - 1);
      const int it = g_it[i];
      it0 = __shfl_sync(FULL, it, 0);
      const unsigned ok = __ballot_sync(FULL, (lane < LN_R) && (cur + lane < ib) && (it + 2 - it0 <= LN_ROWS));
      n = __ffs(~ok) - 1;   // ok is a prefix (it[] is non-decreasing); lane 0 always fits
      const int it1 = __shfl_sync(FULL, it, n - 1);
      if (lane == 0) bulk_load(&sm.rows[0][0], g_rows + (size_t) it0 * NG, (uint32_t) (it1 + 2 - it0) * NG * 16, &sm.mbar);
    }
    // ---- set-up: lane = (radius, task).  While the rows are in flight, task 0: record + first bin | 1: last bin
    const int r_su = lane & (LN_R - 1), task = lane / LN_R;
    if (task < 2) {
      LnRad &lr = sm.rad[r_su];
      const int i = cur + r_su;
      if (r_su < n) {
        const double e_first = line_edge(egrid, 0, grid_mode, zred, lineE);
        const double e_last = line_edge(egrid, n_ener, grid_mode, zred, lineE);
        const double gmin = S.gmin[(size_t) v * NR + i], gmax = S.gmax[(size_t) v * NR + i];
        const bool on_grid = (gmax > e_first) && (gmin < e_last);  // src/Relprofile.cpp:863-878
        if (task == 0) {
          const double del_g = 1. / (gmax - gmin);
          lr.gmin = gmin; lr.del_g = del_g; lr.dgm = gmax - gmin;
          lr.scale = del_g * S.emis[(size_t) v * NR + i];
          lr.weight = trapez_single(S.re + (size_t) v * NR, i, NR) / 2;
          lr.ehlo = (GFAC_H * (gmax - gmin) + gmin) * 1.0;
          lr.ehhi = ((1.0 - GFAC_H) * (gmax - gmin) + gmin) * 1.0;
          lr.gi = i;
          lr.ielo = on_grid ? line_index<GRID_MODE>(G, gmin < e_first ? e_first : gmin, zred, lineE) : n_ener;
        } else {
          lr.iehi = on_grid ? line_index<GRID_MODE>(G, gmax > e_last ? e_last : gmax, zred, lineE) : -1;
        }
      } else if (task == 0) {   // padding of an odd sub-batch: a radius without bins
        lr.gmin = 0.0; lr.del_g = 1.0; lr.dgm = 1.0; lr.scale = 0.0; lr.weight = 0.0; lr.ehlo = 0.0; lr.ehhi = 1.0;
        lr.nlo = 0.0; lr.nhi = 0.0; lr.gi = cur; lr.ielo = n_ener;
      } else {
        lr.iehi = -1;
      }
    }
    // ---- radial interpolation of the staged rows onto the sub-batch's radii (src/Relprofile.cpp:280-293)
    mbar_wait(&sm.mbar, phase);
    phase ^= 1;
#pragma unroll 1
    for (int q = lane; q < n * NG; q += 32) {
      const int r = q / NG, jg = q - r * NG;
      const int i = cur + r;
      const int it = g_it[i] - it0;
      const double fr = S.fr[(size_t) v * NR + i];
      const double2 hi = sm.rows[it][jg], lo = sm.rows[it + 1][jg];   // row `it`: larger radius
      double2 tr;
      tr.x = __dadd_rn(__dmul_rn(fr, hi.x), __dmul_rn(1.0 - fr, lo.x));   // lin1d(fr, lo, hi), uncontracted like k_fine's
      tr.y = __dadd_rn(__dmul_rn(fr, hi.y), __dmul_rn(1.0 - fr, lo.y));
      sm.fine[r][jg] = tr;
    }
    __syncwarp();
    // ---- set-up, tasks 2 and 3: the integrand at the two edge nodes g* = h, 1-h
    if (task >= 2 && r_su < n) {
      RelbCtx c;
      ln_ctx(sm, r_su, g_cosne, limb, c);
      const LnRad &lr = sm.rad[r_su];
      double n0, n1;
      relb2(task == 2 ? lr.ehlo : lr.ehhi, c, n0, n1);
      double norm = 0.0;
      norm = norm + n0;
      norm = norm + n1;
      norm = norm * sqrt(GFAC_H);
      if (task == 2) sm.rad[r_su].nlo = norm; else sm.rad[r_su].nhi = norm;
    }
    // the sub-batch's bin range
    int jlo, jhi;
    {
      const bool valid = r_su < n;
      const int ielo = valid ? sm.rad[r_su].ielo : n_ener, iehi = valid ? sm.rad[r_su].iehi : -1;
      jlo = (iehi >= ielo) ? ielo : n_ener; jhi = (iehi >= ielo) ? iehi : -1;
#pragma unroll
      for (int o = LN_R / 2; o > 0; o >>= 1) {
        jlo = min(jlo, __shfl_xor_sync(FULL, jlo, o));
        jhi = max(jhi, __shfl_xor_sync(FULL, jhi, o));
      }
    }
    const int zold_lo = zlo, zold_hi = zhi;
    if (jhi >= jlo && zhi >= zlo) {   // close a gap between the zone's range so far and this sub-batch's
      if (zhi < jlo) jlo = zhi + 1;
      if (zlo > jhi) jhi = zlo - 1;
    }
    zlo = min(zlo, jlo); zhi = max(zhi, jhi);
    __syncwarp();

    // ---- main loop: tile of 15 bins, half warp = radius, lane = bin edge; the Romberg tiles first
    if (jhi >= jlo) {
      const int half = lane >> 4, hl = lane & 15;
      int ndq = 0;   // bins waiting in the deep queue
      // tiles anchored at j95: tile k covers bins [j95 + k LN_TB, j95 + (k+1) LN_TB)
      const int k_lo = (jlo - j95 >= 0) ? (jlo - j95) / LN_TB : -((j95 - jlo + LN_TB - 1) / LN_TB);
      const int k_hi = (jhi - j95 >= 0) ? (jhi - j95) / LN_TB : -((j95 - jhi + LN_TB - 1) / LN_TB);
      // The bins that need Romberg levels 3+ wait in the queue; it is worked off (one copy of that code) when it cannot
      // take another visit (`again`: the tile is written out, resumed afterwards at the same radius pair) and after the
      // sub-batch's last tile.
      unsigned pm = 0;   // radius pairs of the current tile still to visit
      for (int k = k_hi, again = 0;;) {
        if (again || k < k_lo) {
          __syncwarp();
          if (ndq) { deep_flush(sm, ndq, lane, g_cosne, limb, flux); ndq = 0; }
          if (k < k_lo) break;
        }
        const int j = j95 + k * LN_TB + hl;                     // this lane's edge; its bin if hl < 15
        const double Ea = line_edge(egrid, min(max(j, 0), n_ener), grid_mode, zred, lineE);
        const double Eb = line_edge(egrid, min(max(j + 1, 0), n_ener), grid_mode, zred, lineE);
        const bool binlane = (hl < LN_TB) && (j >= jlo) && (j <= jhi);
        double acc = (binlane && (again || ((j >= zold_lo) && (j <= zold_hi)))) ? flux[j] : 0.0;
        if (!again) {   // the pairs with a radius whose bins reach into this tile (lane r < 8 looks at radius r)
          const int tlo = max(j95 + k * LN_TB, jlo), thi = min(j95 + k * LN_TB + LN_TB - 1, jhi);
          const LnRad &lq = sm.rad[lane & (LN_R - 1)];
          const unsigned m = __ballot_sync(FULL, (lane < n) && (lq.iehi >= tlo) && (lq.ielo <= thi)) & 0xffu;
          pm = (m | (m >> 1)) & 0x55u;
        }
        {
          if (k < 0) {
            // ---------------- midpoint-rule tile (int_romb with lo < 0.95, src/Relprofile.cpp:628-647)
            for (; pm; pm &= pm - 1) {
              const int rsel = (__ffs(pm) - 1) + half;
              const LnRad &lr = sm.rad[rsel];
              const bool in = binlane && (j >= lr.ielo) && (j <= lr.iehi);
              // limits of the quadrature: the bin, cut at the ends of the analytic edge intervals (the reference takes
              // these decisions in g*; an ulp of difference moves the cut by an ulp)
              const double ehlo = lr.ehlo, ehhi = lr.ehhi;
              const bool e_lo = Ea < ehlo, e_hi = Eb > ehhi;
              const double Xa = e_lo ? ehlo : Ea, Xb = e_hi ? ehhi : Eb;
              const double w = Xb - Xa;
              RelbCtx c;
              c.gmin = lr.gmin; c.del_g = lr.del_g; c.scale = lr.scale;
              c.row = smem_u32(&sm.fine[rsel][0]);
              c.cosne = g_cosne + (size_t) lr.gi * NG; c.limb = limb;
              double flu = 0.0;
              if (in && (e_lo || e_hi)) flu = edge_terms(Ea, Eb, c.gmin, c.del_g, lr.dgm, lr.nlo, lr.nhi);
              double m0, m1;
              relb2i((Xb + Xa) / 2.0, c, m0, m1);
              const bool quad = in && w > 0.0;
              const bool hard = quad && (Xa >= LK.e95);   // the lower limit was raised past 0.95: Romberg, by deep_flush
              if (quad && !hard) {
                double f2 = 0.0;
                f2 += m0 * w;
                f2 += m1 * w;
                flu = flu + f2;
              }
              if (__any_sync(FULL, hard)) {
                const unsigned dm = __ballot_sync(FULL, hard);
                if (ndq + __popc(dm) > LN_DQ) break;   // no room: flush, then this visit again
                if (hard) {
                  LnDeep &d = sm.dq[ndq + __popc(dm & ((1u << lane) - 1))];
                  d.a = Xa; d.pas = w; d.rsel = rsel; d.j = j; d.flags = 1 << 8;
                }
                ndq += __popc(dm);
                __syncwarp();
              }
              // ascending-radius accumulation: the lower half's radius first, then the upper half's
              const double own = in ? flu * lr.weight : 0.0;
              const double oth = __shfl_down_sync(FULL, own, 16);   // the upper half's radius comes second
              acc = (acc + own) + oth;                              // (only the lower half's sum is kept)
            }
          } else {
            // ---------------- Romberg tile (src/Relprofile.cpp:524-579), both branches
            for (; pm; pm &= pm - 1) {
              const int rsel = (__ffs(pm) - 1) + half;
              const LnRad &lr = sm.rad[rsel];
              const bool in = binlane && (j >= lr.ielo) && (j <= lr.iehi);
              const double ehlo = lr.ehlo, ehhi = lr.ehhi;
              // every lane evaluates the integrand at its own lower edge, clamped to the quadrature's range: that is
              // the lower limit of its bin and the upper limit of the neighbour's
              const double Xa = (Ea < ehlo) ? ehlo : ((Ea > ehhi) ? ehhi : Ea);
              const double Xb = (Eb < ehlo) ? ehlo : ((Eb > ehhi) ? ehhi : Eb);
              const double pas = Xb - Xa;
              const bool romb = in && (pas > 0.0);
              RelbCtx c;
              c.gmin = lr.gmin; c.del_g = lr.del_g; c.scale = lr.scale;
              c.row = smem_u32(&sm.fine[rsel][0]);
              c.cosne = g_cosne + (size_t) lr.gi * NG; c.limb = limb;
              double flu = 0.0;
              if (in && (Ea < ehlo || Eb > ehhi)) flu = edge_terms(Ea, Eb, c.gmin, c.del_g, lr.dgm, lr.nlo, lr.nhi);
              double sum[2], tq[2][4], res[2];
              bool done[2];
              int need = 0;
              {
                double fa0, fa1, fm0, fm1;
                relb2i(Xa, c, fa0, fa1);
                const double fb0 = __shfl_down_sync(FULL, fa0, 1, 16), fb1 = __shfl_down_sync(FULL, fa1, 1, 16);
                const double pas1 = pas / 2.0, pas2 = pas1 / 2.0;
                relb2i(Xa + pas1 * 1, c, fm0, fm1);
                double t01[2];
                bool lvl2 = false;
  #pragma unroll
                for (int kk = 0; kk < 2; kk++) {
                  const double ta = ((kk ? fa1 : fa0) + (kk ? fb1 : fb0)) / 2.0;
                  sum[kk] = ta;
                  const double t00 = ta * pas;
                  t01[kk] = (ta + (kk ? fm1 : fm0)) * pas1;
                  const double t10 = richardson(1, t01[kk], t00);
                  tq[kk][0] = t10;
                  res[kk] = t10;
                  done[kk] = !not_converged(t10, t00);
                  lvl2 |= !done[kk];
                }
                if (__any_sync(FULL, romb && lvl2)) {
                  double q0, q1, u0, u1;
                  relb2i(Xa + pas2 * 1, c, q0, q1);
                  relb2i(Xa + pas2 * 3, c, u0, u1);
  #pragma unroll
                  for (int kk = 0; kk < 2; kk++) {
                    // ((ta + f(q1)) + f(mid)) + f(q3): the reference's ascending order
                    sum[kk] = ((sum[kk] + (kk ? q1 : q0)) + (kk ? fm1 : fm0)) + (kk ? u1 : u0);
                    if (!done[kk]) {
                      const double t02 = sum[kk] * pas2;
                      const double t11 = richardson(1, t02, t01[kk]);
                      const double t20 = richardson(2, t11, tq[kk][0]);
                      done[kk] = !not_converged(t20, res[kk]);
                      res[kk] = t20;
                      tq[kk][0] = t11; tq[kk][1] = t20;
                      if (!done[kk]) need = 3;
                    }
                  }
                  if (!romb) need = 0;
                }
              }
              if (__any_sync(FULL, need == 3)) {   // queue the bins that go on to level 3
                const unsigned dm = __ballot_sync(FULL, need == 3);
                if (ndq + __popc(dm) > LN_DQ) break;   // no room (a visit defers at most 30 bins): flush, then this visit again
                if (need == 3) {
                  LnDeep &d = sm.dq[ndq + __popc(dm & ((1u << lane) - 1))];
                  d.a = Xa; d.pas = pas; d.rsel = rsel; d.j = j;
  #pragma unroll
                  for (int kk = 0; kk < 2; kk++) {
                    d.sum[kk] = sum[kk];
                    d.tq[kk][0] = done[kk] ? res[kk] : tq[kk][0];
                    d.tq[kk][1] = tq[kk][1];
                  }
                  d.flags = (done[0] ? 1 : 0) | (done[1] ? 2 : 0) | (3 << 8);
                }
                ndq += __popc(dm);
                __syncwarp();
              }
              if (romb && need == 0) {
                double rsum = 0.0;
                rsum += res[0];
                rsum += res[1];
                flu = flu + rsum;
              }
              const double own = in ? flu * lr.weight : 0.0;
              const double oth = __shfl_down_sync(FULL, own, 16);   // the upper half's radius comes second
              acc = (acc + own) + oth;                              // (only the lower half's sum is kept)
            }
          }
        }
        // only the bins this zone touched are written; the range travels with the row
        if (binlane && half == 0) flux[j] = acc;
        again = pm != 0;
        if (!again) k--;
      }
    }
    cur += n;
    __syncwarp();   // the shared-memory stage and the row in HBM are taken over by the next sub-batch
  }
  if (whole) {   // division by the bin energy (renorm_relline_profile, src/Relprofile.cpp:757-762); k_linemerge's otherwise
    for (int j = zlo + lane; j <= zhi; j += 32) {
      const double elo = line_edge(egrid, j, grid_mode, zred, lineE), ehi = line_edge(egrid, j + 1, grid_mode, zred, lineE);
      flux[j] = flux[j] / (0.5 * (elo + ehi));
    }
  }
  if (lane == 0) {
    int *zr = whole ? S.zrange + ((size_t) v * NZMAX + z) * 2 : S.zrpart + ((size_t) v * LINE_PARTS + row) * 2;
    zr[0] = zlo;
    zr[1] = zhi;
  }
}

// Sum of the partial rows of a split zone (ascending runs = ascending radius index), division by the bin energy
// (renorm_relline_profile, src/Relprofile.cpp:757-762) and the zone's range.  One CTA per vector, zones in ascending
// order: the merged row z never lies behind a partial row that is still to be read (z <= z nsplit).
template <int GRID_MODE>
__global__ void __launch_bounds__(128) k_linemerge(const VPar *__restrict__ vps, Scratch S, LineGrid G, int ne_stride, int nz_stride,
                                                   int nz_max) {
  const int v = blockIdx.x, t = threadIdx.x;
  if (S.status[v] != ST_OK) return;
  if (S.reuse && S.reuse[v]) return;
  const VPar &vp = vps[v];
  const int nz = min(vp.nz, nz_max), nsplit = line_parts(vp.nz);
  if (nsplit == 1) return;
  double *base = S.relflux + (size_t) v * nz_stride * ne_stride;
  const int *zp = S.zrpart + (size_t) v * LINE_PARTS * 2;
  for (int z = 0; z < nz; z++) {
    int lo = G.n_ener, hi = -1;
    for (int s = 0; s < nsplit; s++) {
      const int a = zp[(z * nsplit + s) * 2], b = zp[(z * nsplit + s) * 2 + 1];
      if (b >= a) { lo = min(lo, a); hi = max(hi, b); }
    }
    double *out = base + (size_t) z * ne_stride;
    for (int j = lo + t; j <= hi; j += 128) {
      double a = 0.0;
      for (int s = 0; s < nsplit; s++) {
        const int r = z * nsplit + s;
        if (j >= zp[r * 2] && j <= zp[r * 2 + 1]) a += base[(size_t) r * ne_stride + j];
      }
      const double elo = line_edge(G.e, j, GRID_MODE, vp.z, vp.lineE), ehi = line_edge(G.e, j + 1, GRID_MODE, vp.z, vp.lineE);
      out[j] = a / (0.5 * (elo + ehi));
    }
    if (t == 0) {
      S.zrange[((size_t) v * NZMAX + z) * 2] = lo;
      S.zrange[((size_t) v * NZMAX + z) * 2 + 1] = hi;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------- launcher
int line_kernel_init() {
  // all of the carve-out as shared memory: 9.5 KB (+1 KB reserved) per one-warp CTA is what limits residency
  cudaError_t e = cudaFuncSetAttribute(k_line<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(k_line<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  return e == cudaSuccess ? 0 : 1;
}

// profile rows the arena needs per vector for zone counts in [nz_min, nz_max]
int line_rows(int nz_min, int nz_max) {
  int r = nz_max;
  for (int nz = std::max(1, nz_min); nz <= nz_max; nz++) r = std::max(r, nz * line_parts(nz));
  return r;
}
void launch_line(const VPar *vps, const DevTables &T, const Scratch &S, long n, const double *egrid, int n_ener,
                 int grid_mode, int nz_min, int nz_max, cudaStream_t st) {
  int parts = 1;
  for (int nz = std::max(1, nz_min); nz <= nz_max; nz++) parts = std::max(parts, line_parts(nz));
  dim3 grid((unsigned) n, nz_max, parts);   // zone-major launch order: the inner zones (most radii, widest profiles) first
  LineGrid G;
  G.e = egrid; G.n_ener = n_ener; G.mode = grid_mode;
  G.log_lo = std::log(CONV_EMIN);
  G.inv_dlog = (double) NCONV / (std::log(CONV_EMAX) - std::log(CONV_EMIN));
  if (grid_mode == 0) k_line<0><<<grid, LN_NT, 0, st>>>(vps, T, S, G, S.ne_line_cap, S.nz_cap);
  else k_line<1><<<grid, LN_NT, 0, st>>>(vps, T, S, G, S.ne_line_cap, S.nz_cap);
  if (parts > 1) {
    if (grid_mode == 0) k_linemerge<0><<<(unsigned) n, 128, 0, st>>>(vps, S, G, S.ne_line_cap, S.nz_cap, nz_max);
    else k_linemerge<1><<<(unsigned) n, 128, 0, st>>>(vps, S, G, S.ne_line_cap, S.nz_cap, nz_max);
  }
}
int line_launches(int nz_min, int nz_max) {
  for (int nz = std::max(1, nz_min); nz <= nz_max; nz++) if (line_parts(nz) > 1) return 2;
  return 1;
}
int line_max_bins() { return 1 << 24; }   // the zone accumulator lives in the output row: no shared-memory limit

}  // namespace rx
