// line.cu — the relline profile kernel (the FP64 hot loop of the relativistic smearing).
// Compiled with FMA contraction on (build.py): its results feed no discrete decision other than the
// Romberg convergence test, which has a 2 % threshold.
#include <cuda_runtime.h>

#include "common.h"
#include "devutil.cuh"
#include "kernels.h"

namespace rx {

// ---------------------------------------------------------------------------------- k_line
// Relline profile: calc_relline_profile + integ_relline_bin + int_edge + int_romb + romberg_integration +
// relb_func (src/Relprofile.cpp:489-726,835-905) and the division by the bin energy of
// renorm_relline_profile (:757-762).
//
// One CTA per (vector, radial zone).  The zone's radii are processed in sub-batches; inside a sub-batch all
// (radius, energy-bin) pairs that the reference's double loop visits are flattened into one dense item list
// (the bins of one radius are contiguous: [ielo, iehi]), so lanes stay busy whatever the line width is.
//   phase 1  every item is integrated with the Romberg depth capped at 2; the few bins that have not
//            converged by then (the horns of the profile) are pushed on a shared-memory work list
//   phase 2  the work list is integrated densely with the full depth (same arithmetic from scratch)
//   phase 3  per energy bin, the sub-batch's contributions are added in ascending-radius order into the
//            zone accumulator -> no atomics on data, bit-reproducible, the reference's summation order
struct RelbCtx {
  double gmin, gmax, del_g, emis;
  double scale;          // del_g * emis
  const double2 *trff;   // [NG] {branch 0, branch 1} of this radius (global, L1-resident)
  const double2 *cosne;
  const double *gstar;   // shared memory: g* nodes [NG] followed by the inverse node spacings [NG-1]
  int limb;
};

// The integrand (src/Relprofile.cpp:489-521):
//   pow(eg,3) / ((gmax-gmin) * sqrt(g* - g*^2)) * ftrf * emis [* limb]
// evaluated as eg^3 * rsqrt(g* - g*^2) * ftrf * (emis / (gmax-gmin)): one rsqrt instead of a sqrt and two
// divisions (the node-spacing division becomes a multiplication by the tabulated inverse spacing).
__device__ __forceinline__ double relb_func(double eg, int k, const RelbCtx &c) {
  const double egstar = (eg - c.gmin) * c.del_g;
  // bracket in the (uniform up to rounding) g* grid: same result as binary_search(gstar, 40, egstar)
  int ind = (int) ((egstar - GFAC_H) * ((NG - 1) / (1.0 - 2 * GFAC_H)));
  ind = ind < 0 ? 0 : (ind > NG - 2 ? NG - 2 : ind);
  if (ind > 0 && c.gstar[ind] > egstar) ind--;
  else if (ind < NG - 2 && c.gstar[ind + 1] <= egstar) ind++;
  const double inte = (egstar - c.gstar[ind]) * c.gstar[NG + ind];
  const double inte1 = 1.0 - inte;
  const double2 t0 = __ldg(c.trff + ind), t1 = __ldg(c.trff + ind + 1);
  const double ftrf = inte * (k ? t0.y : t0.x) + inte1 * (k ? t1.y : t1.x);
  const double val = (eg * eg * eg) * rsqrt(egstar - egstar * egstar) * ftrf * c.scale;
  if (c.limb == 0) return val;
  const double2 c0 = __ldg(c.cosne + ind), c1 = __ldg(c.cosne + ind + 1);
  const double fmu0 = inte * (k ? c0.y : c0.x) + inte1 * (k ? c1.y : c1.x);
  double limb = 1.0;
  if (c.limb == 1) limb = (1.0 + 2.06 * fmu0);
  else if (c.limb == 2) limb = log(1.0 + 1.0 / fmu0);
  return val * limb;
}

// Romberg integration, src/Relprofile.cpp:524-579.  itermax = 5 is the reference; a smaller cap returns
// with converged = false when the precision goal has not been met yet.
template <int ITERMAX>
__device__ double romberg(double a, double b, int k, const RelbCtx &c, bool &converged) {
  const double prec = 0.02;
  double obtprec = 1.0;
  double prev[ITERMAX + 2], cur[ITERMAX + 2];
  int niter = 0;
  const double r0 = relb_func(a, k, c);
  const double rb = relb_func(b, k, c);
  const double ta = (r0 + rb) / 2.0;
  double pas = b - a;
  prev[0] = ta * pas;
  double last_diag = prev[0];
  while ((obtprec > prec) && (niter <= ITERMAX)) {
    niter++;
    pas = pas / 2.0;
    double s = ta;
    const int npts = (1 << niter) - 1;
    for (int ii = 1; ii <= npts; ii++) s += relb_func(a + pas * ii, k, c);
    cur[0] = s * pas;
    double r = 1.0;
#pragma unroll
    for (int ii = 1; ii <= ITERMAX + 1; ii++) {
      if (ii <= niter) {
        r *= 4.0;
        cur[ii] = (r * cur[ii - 1] - prev[ii - 1]) / (r - 1.0);
      }
    }
    double diag = cur[0];
#pragma unroll
    for (int ii = 1; ii <= ITERMAX + 1; ii++) if (ii == niter) diag = cur[ii];
    obtprec = fabs(diag - last_diag) / diag;
    last_diag = diag;
#pragma unroll
    for (int ii = 0; ii < ITERMAX + 2; ii++) prev[ii] = cur[ii];
  }
  converged = !(obtprec > prec);
  return last_diag;
}

__device__ __forceinline__ double gstar2ener(double g, double gmin, double gmax) { return (g * (gmax - gmin) + gmin) * 1.0; }

__device__ double int_edge(double blo, double bhi, const RelbCtx &c) {  // src/Relprofile.cpp:585-621 (h = GFAC_H)
  double hex, lo, hi;
  if (blo <= 0.5) { hex = GFAC_H; lo = blo; hi = bhi; }
  else { hex = 1.0 - GFAC_H; lo = 1.0 - bhi; hi = 1.0 - blo; }
  double norm = 0.0;
  const double eh = gstar2ener(hex, c.gmin, c.gmax);
  norm = norm + relb_func(eh, 0, c);
  norm = norm + relb_func(eh, 1, c);
  norm = norm * sqrt(GFAC_H);
  return 2 * norm * (sqrt(hi) - sqrt(lo)) * 1.0 * (c.gmax - c.gmin);
}

// src/Relprofile.cpp:650-726.  ITERMAX < 5: `complete` tells whether the capped Romberg runs converged
// (if not, the caller repeats the bin with ITERMAX = 5).
template <int ITERMAX>
__device__ double integ_relline_bin(const RelbCtx &c, double rlo0, double rhi0, bool &complete) {
  complete = true;
  double flu = 0.0;
  double gblo = (rlo0 / 1.0 - c.gmin) * c.del_g;
  if (gblo < 0.0) gblo = 0.0; else if (gblo > 1.0) gblo = 1.0;
  double gbhi = (rhi0 / 1.0 - c.gmin) * c.del_g;
  if (gbhi < 0.0) gbhi = 0.0; else if (gbhi > 1.0) gbhi = 1.0;
  if (gbhi == 0) return 0.0;
  double rlo = rlo0, rhi = rhi0, hlo, hhi;
  const bool edge_lo = (gblo <= GFAC_H), edge_hi = (gbhi >= (1.0 - GFAC_H));
  if (edge_lo) {
    rlo = gstar2ener(GFAC_H, c.gmin, c.gmax);
    if (gbhi <= GFAC_H) rlo = -1.0;
  }
  if (edge_hi) {
    rhi = gstar2ener(1 - GFAC_H, c.gmin, c.gmax);
    if (gblo >= (1.0 - GFAC_H)) rhi = -1.0;
  }
  const bool do_romb = (rhi >= 0) && (rlo >= 0) && (rlo >= 1.0 * 0.95);
  double f2 = 0.0;
  if (do_romb) {  // src/Relprofile.cpp:628-647
    bool c0, c1;
    f2 += romberg<ITERMAX>(rlo, rhi, 0, c, c0);
    if (ITERMAX < 5 && !c0) { complete = false; return 0.0; }
    f2 += romberg<ITERMAX>(rlo, rhi, 1, c, c1);
    if (ITERMAX < 5 && !c1) { complete = false; return 0.0; }
  }
  if (edge_lo) {
    hlo = gblo;
    hhi = GFAC_H;
    if (gbhi <= GFAC_H) hhi = gbhi;
    flu = flu + int_edge(hlo, hhi, c);
  }
  if (edge_hi) {
    hhi = gbhi;
    hlo = 1.0 - GFAC_H;
    if (gblo >= (1.0 - GFAC_H)) hlo = gblo;
    flu = flu + int_edge(hlo, hhi, c);
  }
  if ((rhi >= 0) && (rlo >= 0)) {
    if (!do_romb) {
      const double mid = (rhi + rlo) / 2.0;
      f2 += relb_func(mid, 0, c) * (rhi - rlo);
      f2 += relb_func(mid, 1, c) * (rhi - rlo);
    }
    flu = flu + f2;
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
  double gmin, gmax, del_g, emis, weight;
  int ielo, iehi, off, gi;
};
struct LnSmem {
  double contrib[LN_BUF];
  LnRad rad[LN_MAXR + 1];
  double gstar[2 * NG];
  unsigned short list[LN_BUF];
  int nrad, ndef, cursor, jlo, jhi, resume;   // resume: first bin still to do of radius `cursor` (-1 = all)
};

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
  const int *izone = S.izone + (size_t) v * NR;
  // radii of this zone: izone[] is non-increasing along the (descending-radius) fine grid
  int ia, ib;
  {
    int lo = 0, hi = NR;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (izone[m] > z) lo = m + 1; else hi = m; }
    ia = lo;
    hi = NR;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (izone[m] >= z) lo = m + 1; else hi = m; }
    ib = lo;
  }
  double *flux = S.relflux + ((size_t) v * nz_stride + z) * ne_stride;
  const double *g_re = S.re + (size_t) v * NR;
  const double2 *g_trff = reinterpret_cast<const double2 *>(S.trff) + (size_t) v * NR * NG;
  const double2 *g_cosne = reinterpret_cast<const double2 *>(S.cosne) + (size_t) v * NR * NG;
  if (t < NG) sm.gstar[t] = T.gstar[t];
  if (t < NG - 1) sm.gstar[NG + t] = 1.0 / (T.gstar[t + 1] - T.gstar[t]);
  for (int j = t; j < n_acc; j += LN_NT) acc[j] = 0.0;
  if (t == 0) { sm.cursor = ia; sm.resume = -1; }
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
        lr.emis = S.emis[(size_t) v * NR + i];
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
      sm.jlo = jlo;
      sm.jhi = jhi;
      sm.resume = next_resume;
      sm.cursor = (next_resume >= 0) ? cur + n - 1 : cur + n;
    }
    __syncthreads();
    const int nrad = sm.nrad;
    const int nitems = sm.rad[nrad].off;
    // ---- phase 1: all items, Romberg depth capped
    for (int item = t; item < nitems; item += LN_NT) {
      int lo = 0, hi = nrad;   // radius of this item: last r with off[r] <= item
      while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (sm.rad[m].off <= item) lo = m; else hi = m; }
      const LnRad &lr = sm.rad[lo];
      const int j = lr.ielo + (item - lr.off);
      RelbCtx c;
      c.gmin = lr.gmin; c.gmax = lr.gmax; c.del_g = lr.del_g; c.emis = lr.emis; c.scale = lr.del_g * lr.emis;
      c.trff = g_trff + (size_t) lr.gi * NG; c.cosne = g_cosne + (size_t) lr.gi * NG; c.gstar = sm.gstar; c.limb = limb;
      const double elo = line_edge(egrid, j, grid_mode, zred, lineE), ehi = line_edge(egrid, j + 1, grid_mode, zred, lineE);
      bool complete;
      const double val = integ_relline_bin<1>(c, elo, ehi, complete);
      if (complete) sm.contrib[item] = val;
      else sm.list[atomicAdd(&sm.ndef, 1)] = (unsigned short) item;
    }
    __syncthreads();
    // ---- phase 2: the bins that need the full Romberg depth
    const int ndef = sm.ndef;
    for (int d = t; d < ndef; d += LN_NT) {
      const int item = sm.list[d];
      int lo = 0, hi = nrad;
      while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (sm.rad[m].off <= item) lo = m; else hi = m; }
      const LnRad &lr = sm.rad[lo];
      const int j = lr.ielo + (item - lr.off);
      RelbCtx c;
      c.gmin = lr.gmin; c.gmax = lr.gmax; c.del_g = lr.del_g; c.emis = lr.emis; c.scale = lr.del_g * lr.emis;
      c.trff = g_trff + (size_t) lr.gi * NG; c.cosne = g_cosne + (size_t) lr.gi * NG; c.gstar = sm.gstar; c.limb = limb;
      const double elo = line_edge(egrid, j, grid_mode, zred, lineE), ehi = line_edge(egrid, j + 1, grid_mode, zred, lineE);
      bool complete;
      sm.contrib[item] = integ_relline_bin<5>(c, elo, ehi, complete);
    }
    __syncthreads();
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
  for (int j = t; j < n_ener; j += LN_NT) {
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
