// devutil.cuh — small device helpers shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "common.h"

namespace rx {

#define PI 3.14159265358979323846
#define GFAC_H 5e-3

// ---------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ double lin1d(double f, double lo, double hi) { return f * hi + (1.0 - f) * lo; }

__device__ __forceinline__ double lin2d_f(double f1, double f2, float r11, float r12, float r21, float r22) {
  return (1.0 - f1) * (1.0 - f2) * r11 + (f1) * (1.0 - f2) * r12 + (1.0 - f1) * (f2) * r21 + (f1) * (f2) * r22;
}

// arr ascending: k with arr[k] <= val < arr[k+1], clamped to [0, n-2]  (src/relutility.c:135-171)
template <class T> __device__ __forceinline__ int bsearch_asc(const T *arr, int n, T val) {
  int klo = 0, khi = n - 1;
  while (khi - klo > 1) {
    const int k = (khi + klo) >> 1;
    if (arr[k] > val) khi = k; else klo = k;
  }
  return klo;
}
// arr descending (src/relutility.c:195-211)
__device__ __forceinline__ int bsearch_desc(const double *arr, int n, double val) {
  int klo = 0, khi = n - 1;
  while (khi - klo > 1) {
    const int k = (khi + klo) >> 1;
    if (arr[k] < val) khi = k; else klo = k;
  }
  return klo;
}
// number of entries of ascending arr[0..n) that are <= val
__device__ __forceinline__ int count_le_asc(const double *arr, int n, double val) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int m = (lo + hi) >> 1;
    if (arr[m] <= val) lo = m + 1; else hi = m;
  }
  return lo;
}
// number of entries of descending arr[0..n) that are > val
__device__ __forceinline__ int count_gt_desc(const double *arr, int n, double val) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int m = (lo + hi) >> 1;
    if (arr[m] > val) lo = m + 1; else hi = m;
  }
  return lo;
}

__device__ __forceinline__ double trapez_single(const double *re, int i, int nr) {  // src/relutility.c:233-244
  double dr;
  if (i == 0) dr = 0.5 * (re[i] - re[i + 1]);
  else if (i == nr - 1) dr = 0.5 * (re[i - 1] - re[i]);
  else dr = 0.5 * (re[i - 1] - re[i + 1]);
  return re[i] * dr * PI;
}

__device__ __forceinline__ double doppler_factor(double del, double bet) {  // src/Relphysics.cpp:158-160
  return sqrt(1.0 - bet * bet) / (1.0 + bet * cos(del));
}
__device__ __forceinline__ double relat_abberation(double del, double beta) {  // src/Relphysics.cpp:127-129
  return acos((cos(del) - beta) / (1 - beta * cos(del)));
}
__device__ inline double gi_potential_lp(double r, double a, double h, double bet, double del) {  // src/Relphysics.cpp:163-207
  const double ut_d = ((r * sqrt(r) + a) / (sqrt(r) * sqrt(r * r - 3 * r + 2 * a * sqrt(r))));
  const double ut_h = sqrt((h * h + a * a) / (h * h - 2 * h + a * a));
  const double gi = ut_d / ut_h;
  if (fabs(bet) < 1e-6) return gi;
  const double gam = 1.0 / sqrt(1.0 - bet * bet);
  const double sign = (del > PI / 2) ? -1.0 : 1.0;
  const double delta_eq = h * h - 2 * h + a * a;
  const double sd = sin(del);
  const double hh = (h * h + a * a);
  const double q2 = (sd * sd) * ((hh * hh) / delta_eq) - a * a;
  double beta_fac = sqrt(hh * hh - delta_eq * (q2 + a * a));
  beta_fac = gam * (1.0 + sign * beta_fac / (h * h + a * a) * bet);
  return gi / beta_fac;
}
__device__ __forceinline__ double density_ss73_zone_a(double radius, double rms) {  // src/Relphysics.cpp:123-125
  const double t = (1 - sqrt(rms / radius));
  return pow((radius / rms), (3. / 2)) * (1.0 / (t * t));
}

// fixed-order block reduction (deterministic); all threads must call; result broadcast
template <int NT> __device__ double block_sum(double v, double *red) {
  const int t = threadIdx.x;
  red[t] = v;
  __syncthreads();
#pragma unroll
  for (int s = NT / 2; s > 0; s >>= 1) {
    if (t < s) red[t] += red[t + s];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

// _rebin_spectrum (src/relutility.c:549-601) for one output bin, source spectrum in memory `flu0`
static __device__ double rebin_bin(double elo_out, double ehi_out, const double *__restrict__ e0, const double *flu0, int n0) {
  if (!((e0[0] <= ehi_out) && (e0[n0] >= elo_out))) return 0.0;
  int imin = count_le_asc(e0, n0 + 1, elo_out) - 1;
  if (imin < 0) imin = 0;
  int imax = count_le_asc(e0, n0 + 1, ehi_out);
  if (imax > n0) imax = n0;
  imax -= 1;
  if (imax < 0) imax = 0;
  double elo = elo_out, ehi = ehi_out;
  if (elo < e0[imin]) elo = e0[imin];
  if (ehi > e0[imax + 1]) ehi = e0[imax + 1];
  if (imax == imin) return (ehi - elo) / (e0[imin + 1] - e0[imin]) * flu0[imin];
  const double dmin = (e0[imin + 1] - elo) / (e0[imin + 1] - e0[imin]);
  const double dmax = (ehi - e0[imax]) / (e0[imax + 1] - e0[imax]);
  double f = 0.0;
  f += flu0[imin] * dmin + flu0[imax] * dmax;
  for (int jj = imin + 1; jj <= imax - 1; jj++) f += flu0[jj];
  return f;
}

}  // namespace rx
