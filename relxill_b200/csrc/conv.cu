// conv.cu — k_conv: per-zone FFT convolution, primary spectrum and rebin to the caller's grid.
// Compiled with FMA contraction on (build.py): butterflies and weighted sums only; the few discrete
// decisions (zone skipped when its profile sums to < 1e-12, bin searches) compare sums / table values that
// involve no multiply-add.
//
// relxill_convolution_multizone (src/Relxill.cpp:432-482) with fftw_conv_spectrum + calcFFTNormFactor
// (src/Relbase.cpp:119-213), PrimarySource::add_primary_spectrum (src/PrimarySource.cpp:66-125) and
// rebin_and_normalize_relxill_for_xspec (src/Relxill.cpp:261-278).  One CTA per vector.
//
// FFT work per zone is cut from three transforms (reference) to one:
//   * the two real inputs (xillver spectrum, rotated line profile) share one complex forward transform;
//   * the band sum of the convolved zone that the normalisation needs is a linear functional of the
//     product spectrum, sum_i w_i out_i = sum_k P[k] conj(W[k]) with W = DFT(band/cf) tabulated at
//     load, so it is taken in the frequency domain and no per-zone inverse transform is needed;
//   * the normalised product spectra of all zones are accumulated in the frequency domain and one inverse
//     transform per vector brings the sum back (inverse = forward transform of the conjugate).
// The transform is a 4096-point radix-16 Stockham autosort FFT in shared memory: 3 passes (16^3 = 4096), one 16-point
// butterfly per thread per pass, 256 threads with 16 points each.  The work array is complex-interleaved (one 128-bit
// shared access per point) and padded by one point every 16, which makes every pass's gather and scatter
// bank-conflict free per quarter-warp.
// With the radix-8 transform of round 1 (4 passes, 512 threads at 64 registers) the L1/shared-memory data pipe bounded
// the kernel (ncu: wavefronts at ~70 % of peak, FP64 pipe 45 %: 64 shared accesses per 8 points and zone, plus the
// register spills, which ride the same pipe).  Radix 16 exchanges through shared memory twice instead of three times
// (36 accesses per 8 points and zone), needs 6 block barriers per transform instead of 8 and, at 128 registers per
// thread for the same register file per CTA, spills nothing.  The rest of the zone loop moves as little as possible
// through that pipe too: the packing phase is a plain coalesced load (k_xill files the zone spectra on the convolution
// grid), the last pass stores only the half of the result that other threads need, the accumulators are 128-bit.
#include <cuda_runtime.h>

#include "common.h"
#include "devutil.cuh"
#include "kernels.h"

namespace rx {

constexpr int CONV_R = 16;                  // radix: points per thread
constexpr int CONV_NT = NCONV / CONV_R;     // 256
constexpr int CV_PADN = NCONV + NCONV / 16;
__device__ __forceinline__ int cv_pad(int i) { return i + (i >> 4); }

struct ConvSmem {
  double2 z[CV_PADN];
  double2 acc[NCONV / 2 + 1];                    // accumulated half spectrum (one 128-bit access per bin); reused as
                                                 // the final 4096-bin result (4098 doubles)
  double red[4 * (CONV_NT / 32)];
  double bc[8];
  double nyq;                        // product spectrum at the Nyquist bin (thread 0)
  double part[4 * (CONV_NT / 32)];   // per-warp partial sums of the packing phase, finished after the transform
  double2 tw1[16];                   // twiddles of pass 1: exp(-2 pi i k / 256), k < 16
};

// sums NV values over the block (fixed order: shuffle tree inside warps, then over the warps)
template <int NV>
__device__ __forceinline__ void block_sum_n(double (&v)[NV], ConvSmem &sm) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
#pragma unroll
  for (int q = 0; q < NV; q++)
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; q++) sm.red[q * (CONV_NT / 32) + w] = v[q];
  }
  __syncthreads();
  if (t < NV) {
    double s = 0.0;
    for (int i = 0; i < CONV_NT / 32; i++) s += sm.red[t * (CONV_NT / 32) + i];
    sm.bc[t] = s;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; q++) v[q] = sm.bc[q];
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void cmul(double &xr, double &xi, double wr, double wi) {
  const double a = xr * wr - xi * wi;
  xi = xr * wi + xi * wr;
  xr = a;
}
__device__ __forceinline__ double2 cprod(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// 4-point DFT (forward sign) of (x0, x1, x2, x3) -> natural order, in place
__device__ __forceinline__ void fft4(double &r0, double &i0, double &r1, double &i1, double &r2, double &i2, double &r3, double &i3) {
  const double ar = r0 + r2, ai = i0 + i2, br = r0 - r2, bi = i0 - i2;
  const double cr = r1 + r3, ci = i1 + i3, dr = r1 - r3, di = i1 - i3;
  r0 = ar + cr; i0 = ai + ci;
  r2 = ar - cr; i2 = ai - ci;
  r1 = br + di; i1 = bi - dr;   // b - i d
  r3 = br - di; i3 = bi + dr;   // b + i d
}
// 16-point DFT (forward sign), 4 x 4 decomposition (n = 4 n1 + n2 -> k = k1 + 4 k2), outputs in natural order
__device__ __forceinline__ void fft16(double (&r)[16], double (&i)[16]) {
  const double h = 0.70710678118654752440, c = 0.92387953251128675613, s = 0.38268343236508977173;
  // step 1: for every n2, a 4-point DFT over n1 of x[4 n1 + n2]; the result y[n2][k1] stays at index 4 k1 + n2
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) fft4(r[n2], i[n2], r[4 + n2], i[4 + n2], r[8 + n2], i[8 + n2], r[12 + n2], i[12 + n2]);
  // step 2: twiddles W16^(n2 k1) on y[n2][k1] (index 4 k1 + n2)
  {
    double x, y;
    // k1 = 1: W^1, W^2, W^3
    x = r[5]; y = i[5]; r[5] = x * c + y * s; i[5] = y * c - x * s;
    x = r[6]; y = i[6]; r[6] = (x + y) * h; i[6] = (y - x) * h;
    x = r[7]; y = i[7]; r[7] = x * s + y * c; i[7] = y * s - x * c;
    // k1 = 2: W^2, W^4 = -i, W^6
    x = r[9]; y = i[9]; r[9] = (x + y) * h; i[9] = (y - x) * h;
    x = r[10]; y = i[10]; r[10] = y; i[10] = -x;
    x = r[11]; y = i[11]; r[11] = (y - x) * h; i[11] = (-x - y) * h;
    // k1 = 3: W^3, W^6, W^9 = -W^1
    x = r[13]; y = i[13]; r[13] = x * s + y * c; i[13] = y * s - x * c;
    x = r[14]; y = i[14]; r[14] = (y - x) * h; i[14] = (-x - y) * h;
    x = r[15]; y = i[15]; r[15] = -(x * c + y * s); i[15] = -(y * c - x * s);
  }
  // step 3: for every k1, a 4-point DFT over n2 of y[n2][k1] -> X[k1 + 4 k2] at index 4 k1 + k2
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) fft4(r[4 * k1], i[4 * k1], r[4 * k1 + 1], i[4 * k1 + 1], r[4 * k1 + 2], i[4 * k1 + 2], r[4 * k1 + 3], i[4 * k1 + 3]);
  // natural order: X[k1 + 4 k2] sits at 4 k1 + k2 -> transpose the 4 x 4 index
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b2 = a + 1; b2 < 4; b2++) {
      double tt = r[4 * a + b2]; r[4 * a + b2] = r[4 * b2 + a]; r[4 * b2 + a] = tt;
      tt = i[4 * a + b2]; i[4 * a + b2] = i[4 * b2 + a]; i[4 * b2 + a] = tt;
    }
}

// Forward FFT of 4096 points; all CONV_NT threads call.  The input arrives in registers: thread j holds the points
// j + 256 q, q = 0..15 (exactly what the first pass needs), so the packing code hands its values over without a
// round trip through shared memory.  Result in the padded array z.
// The twiddles of pass p are w^q with w = exp(-2 pi i k / (16 ns)), k = j mod ns: they depend on the thread only.
// The 16 distinct ones of pass 1 come from a small shared table (tw1), the last pass's w is re-read per transform (one
// L1-resident 16-byte load); the powers come from squarings and products, applied as soon as they exist.
// The transform comes in two pieces, the first pass and the two twiddled ones, so that the caller can put work
// between them when the 32 packed values have left the registers.
// UPPER: only the upper half of the result (points 2048..4095, the thread's q = 8..15) is written to z; the lower
// half (points j + 256 q, q < 8) stays in r/im — all the split of two real transforms needs, since the partner of
// point k is point 4096 - k.
// First pass of the transform: a 16-point DFT of the thread's own points, written to z (no twiddles, no barrier)
__device__ __forceinline__ void fft4096_first(double (&r)[16], double (&im)[16], double2 *z) {
  const int j = threadIdx.x;
  fft16(r, im);
#pragma unroll
  for (int q = 0; q < 16; q++) z[cv_pad((j << 4) + q)] = make_double2(r[q], im[q]);
}
// The two twiddled passes; starts with the barrier that publishes the first pass
template <bool UPPER>
__device__ __forceinline__ void fft4096_rest(double (&r)[16], double (&im)[16], double2 *z, const double2 *tw1, const double2 *tw2) {
  const int j = threadIdx.x;
  __syncthreads();
#pragma unroll
  for (int pass = 1; pass < 3; pass++) {
    const int ns = 1 << (4 * pass);          // 16, 256
    const int k = j & (ns - 1);
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const double2 c = z[cv_pad(j + q * (NCONV / 16))];
      r[q] = c.x;
      im[q] = c.y;
    }
    {
      const double2 w1 = (pass == 1) ? tw1[k] : __ldg(tw2 + j);
      cmul(r[1], im[1], w1.x, w1.y);
      const double2 w2 = make_double2(w1.x * w1.x - w1.y * w1.y, 2.0 * w1.x * w1.y);
      cmul(r[2], im[2], w2.x, w2.y);
      const double2 w3 = cprod(w2, w1);
      cmul(r[3], im[3], w3.x, w3.y);
      const double2 w4 = make_double2(w2.x * w2.x - w2.y * w2.y, 2.0 * w2.x * w2.y);
      cmul(r[4], im[4], w4.x, w4.y);
      const double2 w5 = cprod(w4, w1);
      cmul(r[5], im[5], w5.x, w5.y);
      const double2 w6 = cprod(w4, w2);
      cmul(r[6], im[6], w6.x, w6.y);
      const double2 w7 = cprod(w4, w3);
      cmul(r[7], im[7], w7.x, w7.y);
      const double2 w8 = make_double2(w4.x * w4.x - w4.y * w4.y, 2.0 * w4.x * w4.y);
      cmul(r[8], im[8], w8.x, w8.y);
      double2 w = cprod(w8, w1);
      cmul(r[9], im[9], w.x, w.y);
      w = cprod(w8, w2);
      cmul(r[10], im[10], w.x, w.y);
      w = cprod(w8, w3);
      cmul(r[11], im[11], w.x, w.y);
      w = cprod(w8, w4);
      cmul(r[12], im[12], w.x, w.y);
      w = cprod(w8, w5);
      cmul(r[13], im[13], w.x, w.y);
      w = cprod(w8, w6);
      cmul(r[14], im[14], w.x, w.y);
      w = cprod(w8, w7);
      cmul(r[15], im[15], w.x, w.y);
    }
    fft16(r, im);
    __syncthreads();
    const int j0 = ((j - k) << 4) + k;       // (j / ns) * ns * 16 + k
#pragma unroll
    for (int q = (UPPER && pass == 2) ? 8 : 0; q < 16; q++) z[cv_pad(j0 + q * ns)] = make_double2(r[q], im[q]);
    __syncthreads();
  }
}
template <bool UPPER>
__device__ __forceinline__ void fft4096(double (&r)[16], double (&im)[16], double2 *z, const double2 *tw1, const double2 *tw2) {
  fft4096_first(r, im, z);
  fft4096_rest<UPPER>(r, im, z, tw1, tw2);
}

struct ConvArgs {
  const double *user_e;   // [n_flux+1] device
  int n_flux;
  double *out;            // [C][n_flux] device
  double *total;          // [C][NCONV] device (probe; may be null)
  int which;              // xillver table index
  int nz_stride, ne_stride;
  int mode;               // 0 relxill, 1 convolution model (input spectrum in `out`)
  // resolved on the host so that the kernel reads them from the constant bank instead of keeping them in registers
  const int2 *rb_ii;      // rebin map of the xillver table in use (XillDev::rb_ii / rb_dd)
  const double2 *rb_dd;
  int xstride;            // row stride of the zone spectra
  int xc_first, xc_n;     // CG: the zone spectra are on the convolution grid, bins [xc_first, xc_first + xc_n) (xill.cu)
  int renorm3;            // RELXILL_RENORMALIZE=1: scale the spectrum to 1 cts/s/keV/cm2 at 3 keV before the final rebin
};

// two CTAs of 256 threads per SM, 128 registers per thread
// CG: k_xill filed the zone spectra on the convolution grid already, the packing phase is a plain coalesced load
template <bool CG>
__global__ void __launch_bounds__(CONV_NT, 2) k_conv(const VPar *__restrict__ vps, DevTables T, Scratch S, ConvArgs A) {
  extern __shared__ __align__(16) unsigned char smraw[];
  ConvSmem &sm = *reinterpret_cast<ConvSmem *>(smraw);
  const int v = blockIdx.x, t = threadIdx.x;
  double *o = A.out + (size_t) v * A.n_flux;
  const VPar &vp = vps[v];
  if (S.status[v] != ST_OK) {
    for (int j = t; j < A.n_flux; j += CONV_NT) o[j] = 0.0;
    return;
  }
  // the previous run's convolution-grid spectrum of this vector is still valid (same parameters up to z): only the
  // rebin onto the caller's grid is left (the reference's RelxillCache hit, src/Relxill.cpp:296-300,405)
  const bool reuse_all = (A.mode == 0) && A.total && S.reuse && (S.reuse[v] & REUSE_ALL);
  const int nz = reuse_all ? 0 : (A.mode == 0) ? vp.nz : 1;
  const int i1 = T.conv_i1kev, b0 = T.conv_b0, b1 = T.conv_b1;
  const double2 *tw = reinterpret_cast<const double2 *>(T.tw);
  const double2 *cw = reinterpret_cast<const double2 *>(T.conv_w);
  for (int k = t; k <= NCONV / 2; k += CONV_NT) sm.acc[k] = make_double2(0.0, 0.0);
  // tw[m] = exp(-2 pi i m / 4096): the last pass's twiddle of this thread is re-read per transform, the 16 distinct ones
  // of the pass before it sit in shared memory
  if (t < 16) sm.tw1[t] = __ldg(tw + t * 16);
  double bal_prev = 0.0;   // ratio of the two input scales in the last zone that had one (0: none yet)
  // What does not change from zone to zone, per thread: the rotated bin of the line profile that goes with the
  // thread's bin i(u) = t + 256 u is ri(u) = r0 + 256 ((c0 + u) mod 16), and whether i(u) (bits 0-15) and ri(u)
  // (bits 16-31) lie in the normalisation band
  const int r0 = (t + i1) & (CONV_NT - 1), c0 = (t + i1) >> 8;
  unsigned band = 0u;
#pragma unroll
  for (int u = 0; u < CONV_R; u++) {
    const int i = t + u * CONV_NT, ri = r0 + (((c0 + u) & (CONV_R - 1)) << 8);
    if (i >= b0 && i <= b1) band |= 1u << u;
    if (ri >= b0 && ri <= b1) band |= 0x10000u << u;
  }
  const double *rel_v = S.relflux + (size_t) v * A.nz_stride * A.ne_stride + r0;
  const double *xz_v = S.xillz + (size_t) v * A.nz_stride * A.xstride;
  const int *zr_v = S.zrange + (size_t) v * NZMAX * 2;
  for (int z = 0; z < nz; z++) {
    const double *relr = rel_v + (size_t) z * A.ne_stride;
    const double *xz = xz_v + (size_t) z * A.xstride;
    int rjlo = zr_v[2 * z], rjw = zr_v[2 * z + 1] - rjlo;   // bins of the profile that were written: [rjlo, rjlo + rjw]
    if (rjw < 0) { rjlo = 0x40000000; rjw = 0; }
    // ---- pack the zone's spectrum rebinned onto the convolution grid (real part) and its line profile, rotated so
    //      that 1 keV sits at index 0 (imaginary part); band and total sums on the way.  The reference multiplies
    //      both by E_mid/dE before the transform and divides the result by it afterwards (src/Relbase.cpp:93-103,
    //      186-190); on the logarithmic convolution grid that factor is one constant (to 1e-13, checked at load,
    //      tables.cu) and cancels against the normalisation, so it is left out
    double sums[3] = {0.0, 0.0, 0.0};         // all rel, band x, band rel
    double re[CONV_R], im[CONV_R];            // bins t + 256 u: the first FFT pass takes them from here
    if (A.mode == 0 && z + 1 < nz) {
      // the next zone's two rows are pulled into L2 while this zone is transformed: the packing loads are the
      // kernel's only DRAM accesses and there are too few warps to hide their latency (one 128-byte line per thread:
      // 188 lines of the zone spectrum, up to 256 of the written part of the line profile)
      if (t < (CG ? (A.xc_n + 15) >> 4 : 188)) prefetch_l2(xz + A.xstride + t * 16);
      {
        const int nlo = zr_v[2 * z + 2], nhi = zr_v[2 * z + 3];
        const int l = (nlo & ~15) + t * 16;
        if (l <= nhi) prefetch_l2(relr - r0 + A.ne_stride + l);
      }
    }
    if (A.mode == 0) {
      // all loads of the 8 bins are independent: bins outside the table grid carry zero weights, the common
      // spans (1-3 source bins) are branch-free
#pragma unroll
      for (int u = 0; u < CONV_R; u++) {
        const int i = t + u * CONV_NT;
        double f = 0.0;
        if (CG) {
          const int c = i - A.xc_first;
          if ((unsigned) c < (unsigned) A.xc_n) f = xz[c];
        } else {
          const int2 ii = __ldg(A.rb_ii + i);      // (0, 0) with zero weights outside the table grid
          const double2 dd = __ldg(A.rb_dd + i);
          f += xz[ii.x] * dd.x + xz[ii.y] * dd.y;
          if (ii.y - ii.x >= 2) {
            f += xz[ii.x + 1];
            for (int jj = ii.x + 2; jj <= ii.y - 1; jj++) f += xz[jj];
          }
        }
        const int ku = ((c0 + u) & (CONV_R - 1)) << 8;
        const double r = ((unsigned) (r0 + ku - rjlo) <= (unsigned) rjw) ? relr[ku] : 0.0;
        re[u] = f;
        im[u] = r;
        sums[0] += r;
        if (band & (1u << u)) sums[1] += f;
        if (band & (0x10000u << u)) sums[2] += r;
      }
    } else {
#pragma unroll 1
      for (int i = t; i < NCONV; i += CONV_NT) {
        const double f = rebin_bin(T.econv[i], T.econv[i + 1], A.user_e, o, A.n_flux);
        const int ri = (i + i1) & (NCONV - 1);
        const double r = ((unsigned) (ri - rjlo) <= (unsigned) rjw) ? relr[ri - r0] : 0.0;
        sm.z[cv_pad(i)] = make_double2(f, r);
        sums[0] += r;
        if (i >= b0 && i <= b1) sums[1] += f;
        if (ri >= b0 && ri <= b1) sums[2] += r;
      }
#pragma unroll
      for (int u = 0; u < CONV_R; u++) {   // the thread's own values
        const double2 c = sm.z[cv_pad(t + u * CONV_NT)];
        re[u] = c.x;
        im[u] = c.y;
      }
    }
    // Both real inputs ride one complex transform, so they are brought to the same scale first (any factor cancels
    // in the norm).  The exact ratio needs the block-wide sums; from the second zone on the previous zone's ratio
    // is close enough (it only guards the rounding of the smaller input), and the sums are finished after the
    // transform together with the band sum — one block reduction per zone instead of two.
    const bool exact = (bal_prev == 0.0) || vp.renorm;   // uniform over the block
    double rscale = 1.0, s_xill = 0.0, s_rel = 0.0, yscale;
    if (exact) {
      block_sum_n<3>(sums, sm);
      const double srel_all = sums[0];
      if (vp.renorm) rscale = vp.relline_norm / srel_all;                  // renorm_relline_profile (one-zone models)
      const double srel_n = vp.renorm ? srel_all * rscale : srel_all;
      if (srel_n < 1e-12) { __syncthreads(); continue; }                   // src/Relxill.cpp:455-457
      // the scale of the spectrum: its band sum (spectra are non-negative; the band covers the whole table grid)
      const double bal = (fabs(sums[1]) > 0.0 && srel_n > 0.0) ? fabs(sums[1]) / srel_n : 1.0;
      s_xill = sums[1];
      s_rel = vp.renorm ? sums[2] * rscale : sums[2];
      yscale = rscale * bal;
      if (!vp.renorm) bal_prev = bal;
#pragma unroll
      for (int u = 0; u < CONV_R; u++) im[u] *= yscale;
      fft4096_first(re, im, sm.z);
    } else {
      // the first pass goes ahead of the warp reductions of the packing sums: the 16 packed values leave the registers
      // before the shuffles need them (they were being spilled across the reduction)
      yscale = bal_prev;
#pragma unroll
      for (int u = 0; u < CONV_R; u++) im[u] *= yscale;
      fft4096_first(re, im, sm.z);
#pragma unroll
      for (int q = 0; q < 3; q++)
        for (int o = 16; o > 0; o >>= 1) sums[q] += __shfl_xor_sync(0xffffffffu, sums[q], o);
      if ((t & 31) == 0) {
#pragma unroll
        for (int q = 0; q < 3; q++) sm.part[q * (CONV_NT / 32) + (t >> 5)] = sums[q];
      }
    }
    fft4096_rest<true>(re, im, sm.z, sm.tw1, tw);
    // ---- split, product spectrum, band sum of the convolved zone in the frequency domain.  Point k = t + 256 nk
    // (nk < 8) is still in the thread's registers; its partner 4096 - k is point q = 15 - nk of thread 256 - t, in the
    // stored upper half (k = 0 is its own partner; thread 0 also takes k = 2048, its own point q = 8)
    double dot[1] = {0.0};
    double pr_[CONV_R / 2], pi_[CONV_R / 2];
#pragma unroll
    for (int nk = 0; nk < CONV_R / 2; nk++) {
      const int k = t + nk * CONV_NT;
      const int kk = (NCONV - k) & (NCONV - 1);
      const double2 q = (k == 0) ? make_double2(re[0], im[0]) : sm.z[cv_pad(kk)];
      const double a = re[nk], b = im[nk], c = q.x, d = q.y;
      const double Xr = 0.5 * (a + c), Xi = 0.5 * (b - d);
      const double Yr = 0.5 * (b + d), Yi = 0.5 * (c - a);
      const double Pr = Xr * Yr - Xi * Yi, Pi = Xr * Yi + Xi * Yr;
      pr_[nk] = Pr;
      pi_[nk] = Pi;
      const double wgt = (k == 0) ? 1.0 : 2.0;
      const double2 w = __ldg(cw + k);
      dot[0] += wgt * (Pr * w.x + Pi * w.y);
    }
    if (t == 0) {   // the Nyquist bin (thread 0's own point q = 8; both spectra are real there) waits in shared memory
      const double pn = re[CONV_R / 2] * im[CONV_R / 2];
      sm.nyq = pn;
      dot[0] += pn * __ldg(cw + NCONV / 2).x;
    }
    if (exact) {
      block_sum_n<1>(dot, sm);
    } else {
      // the band sum and, from the per-warp parts of the packing phase (visible: the transform's barriers lie in
      // between), the three packing sums: four rows of 16 partials, one thread each
      for (int o = 16; o > 0; o >>= 1) dot[0] += __shfl_xor_sync(0xffffffffu, dot[0], o);
      __syncthreads();
      if ((t & 31) == 0) sm.red[t >> 5] = dot[0];
      __syncthreads();
      if (t < 4) {
        const double *row = (t == 0) ? sm.red : sm.part + (t - 1) * (CONV_NT / 32);
        double a = 0.0;
        for (int i = 0; i < CONV_NT / 32; i++) a += row[i];
        sm.bc[t] = a;
      }
      __syncthreads();
      dot[0] = sm.bc[0];
      const double srel_all = sm.bc[1];
      if (srel_all < 1e-12) { __syncthreads(); continue; }                 // src/Relxill.cpp:455-457
      s_xill = sm.bc[2];
      s_rel = sm.bc[3];
      if (fabs(s_xill) > 0.0) bal_prev = fabs(s_xill) / srel_all;
    }
    const double norm = s_rel * s_xill / dot[0];
#pragma unroll
    for (int nk = 0; nk < CONV_R / 2; nk++) {
      const int k = t + nk * CONV_NT;
      double2 a = sm.acc[k];
      a.x += norm * pr_[nk];
      a.y += norm * pi_[nk];
      sm.acc[k] = a;
    }
    if (t == 0) sm.acc[NCONV / 2].x += norm * sm.nyq;
    __syncthreads();
  }
  // ---- one inverse transform for the whole vector: out = Re(FFT(conj(A)))
  if (!reuse_all) {
    double re[CONV_R], im[CONV_R];
#pragma unroll
    for (int u = 0; u < CONV_R; u++) {   // Hermitian extension of the accumulated half spectrum, conjugated
      const int i = t + u * CONV_NT;
      const int k = (i <= NCONV / 2) ? i : NCONV - i;
      const double2 a = sm.acc[k];
      const double ai = (k == 0 || k == NCONV / 2) ? 0.0 : a.y;
      re[u] = a.x;
      im[u] = (i <= NCONV / 2) ? -ai : ai;
    }
    fft4096<false>(re, im, sm.z, sm.tw1, tw);
  }
  __syncthreads();       // every thread has taken its part of the accumulated spectrum
  double *acc = reinterpret_cast<double *>(sm.acc);   // 4098 doubles
  if (reuse_all) {
    __syncthreads();
    for (int i = t; i < NCONV; i += CONV_NT) acc[i] = A.total[(size_t) v * NCONV + i];
  } else {
    for (int i = t; i < NCONV; i += CONV_NT) acc[i] = sm.z[cv_pad(i)].x;
  }
  __syncthreads();
  if (A.mode == 0 && !reuse_all) {
    // primary spectrum on the convolution grid (cutoff power law here; nthcomp is added by k_prim_nthcomp)
    double refl_scale, prim_scale;
    if (vp.emis_type != EMIS_LP) {
      refl_scale = fabs(vp.refl_frac);
      prim_scale = 1.0;
    } else {
      const double *rf = S.reflfrac + (size_t) v * 8;
      double rfi = vp.refl_frac;
      if (vp.boost) rfi *= rf[0];
      prim_scale = rf[4] / 0.5 * pow(vp.eshift_obs, vp.gam);
      if (vp.beta > 1e-4) prim_scale *= vp.doppler_obs * vp.doppler_obs;
      refl_scale = (fabs(rfi)) / rf[0];
    }
    const double nsrc = S.nsrc[v];
    const bool add_prim = (vp.refl_frac >= 0);
    if (vp.prim_type == PRIM_ECUT) {
      const double ecut = vp.ect * vp.eshift_obs;
      const double ex0 = exp(1.0 / ecut);
      for (int i = t; i < NCONV; i += CONV_NT) {
        const double e0 = T.econv[i], e1 = T.econv[i + 1];
        const double en = 0.5 * (e0 + e1);
        double pr = ex0 * pow(en, -vp.gam) * exp(-en / ecut) * (e1 - e0);
        pr *= nsrc;
        if (vp.emis_type == EMIS_LP) pr *= prim_scale;
        double tot = acc[i] * refl_scale;
        if (add_prim) tot += pr;
        acc[i] = tot;
      }
    } else if (vp.prim_type == PRIM_BB) {   // spec_blackbody (src/Xillspec.cpp:262-269); these flavours have no lamp post
      const double kt4 = pow(vp.ktbb, 4);
      for (int i = t; i < NCONV; i += CONV_NT) {
        const double e0 = T.econv[i], e1 = T.econv[i + 1];
        const double en = 0.5 * (e0 + e1);
        double pr = en * en / (kt4 * (exp(en / vp.ktbb) - 1));
        pr *= (e1 - e0);
        pr *= nsrc;
        double tot = acc[i] * refl_scale;
        if (add_prim) tot += pr;
        acc[i] = tot;
      }
    } else {
      for (int i = t; i < NCONV; i += CONV_NT) acc[i] = acc[i] * refl_scale;  // primary added afterwards
    }
    __syncthreads();
    if (A.total) for (int i = t; i < NCONV; i += CONV_NT) A.total[(size_t) v * NCONV + i] = acc[i];
  }
  __syncthreads();
  // rebin to the caller's grid (shifted by 1+z), src/Relxill.cpp:261-278
  double rn = 1.0;   // renorm_relxill_spectrum_1keV (src/Relxill.cpp:249-259); the Cp models get it in k_prim_nth, after their primary
  if (A.renorm3 && A.mode == 0 && vp.prim_type != PRIM_NTHCOMP) {
    const int i3 = T.conv_i3kev;
    rn = 1.0 / (acc[i3] / (T.econv[i3 + 1] - T.econv[i3]));
  }
  for (int j = t; j < A.n_flux; j += CONV_NT) {
    double elo = A.user_e[j], ehi = A.user_e[j + 1];
    if (A.mode == 0 && vp.z > 0) { elo *= (1 + vp.z); ehi *= (1 + vp.z); }
    double f = rebin_bin(elo, ehi, T.econv, acc, NCONV);
    if (A.mode == 1 && (ehi < 0.01 || elo > 1000.0)) f = 0;   // src/Relbase.cpp:233-246
    o[j] = (A.renorm3 && A.mode == 0) ? f * rn : f;
  }
}

int conv_kernel_init() {
  cudaError_t e = cudaFuncSetAttribute(k_conv<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(ConvSmem));
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(k_conv<false>, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(k_conv<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(ConvSmem));
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(k_conv<true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared);
  return e != cudaSuccess;
}

void launch_conv(const VPar *vps, const DevTables &T, const Scratch &S, long n, const double *user_e, int n_flux,
                 double *out, double *total, int which, int mode, int conv_grid, int renorm3, cudaStream_t st) {
  ConvArgs A;
  A.renorm3 = renorm3;
  A.user_e = user_e; A.n_flux = n_flux; A.out = out; A.total = total; A.which = which;
  A.nz_stride = S.nz_cap; A.ne_stride = S.ne_line_cap; A.mode = mode;
  A.rb_ii = reinterpret_cast<const int2 *>(T.xill[which].rb_ii);
  A.rb_dd = reinterpret_cast<const double2 *>(T.xill[which].rb_dd);
  const bool cg = conv_grid && mode == 0;
  A.xstride = cg ? T.xill[which].xc_stride : T.xill[which].stride;
  A.xc_first = T.xill[which].xc_first;
  A.xc_n = T.xill[which].xc_n;
  if (cg) k_conv<true><<<(unsigned) n, CONV_NT, sizeof(ConvSmem), st>>>(vps, T, S, A);
  else k_conv<false><<<(unsigned) n, CONV_NT, sizeof(ConvSmem), st>>>(vps, T, S, A);
}

}  // namespace rx
