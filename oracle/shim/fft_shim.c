/* FFTW3 entry points used by the reference (src/Relbase.cpp), as an exact
 * fp64 iterative radix-2 DFT.  Conventions match FFTW: forward sign -1,
 * unnormalised both ways; r2c writes bins 0..n/2, c2r reads bins 0..n/2 and
 * assumes Hermitian symmetry.  The overall scale cancels in the reference's
 * calcFFTNormFactor, so any correct DFT gives the same spectra to ~1e-15.
 * Test infrastructure, not product code. */
#include "fftw3.h"
#include <math.h>
#include <stdlib.h>

struct oracle_fftw_plan_s {
  int n, inverse;
  double *rbuf;        /* real side */
  fftw_complex *cbuf;  /* complex side */
  double *wr, *wi, *xr, *xi;
};

static void fft_pow2(int n, int sign, double *xr, double *xi, const double *wr, const double *wi) {
  for (int i = 1, j = 0; i < n; i++) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      double t = xr[i]; xr[i] = xr[j]; xr[j] = t;
      t = xi[i]; xi[i] = xi[j]; xi[j] = t;
    }
  }
  for (int len = 2; len <= n; len <<= 1) {
    int half = len >> 1, step = n / len;
    for (int i = 0; i < n; i += len) {
      for (int k = 0; k < half; k++) {
        double c = wr[k * step], s = sign * wi[k * step];
        double ur = xr[i + k], ui = xi[i + k];
        double vr = xr[i + k + half] * c - xi[i + k + half] * s;
        double vi = xr[i + k + half] * s + xi[i + k + half] * c;
        xr[i + k] = ur + vr; xi[i + k] = ui + vi;
        xr[i + k + half] = ur - vr; xi[i + k + half] = ui - vi;
      }
    }
  }
}

static fftw_plan mkplan(int n, int inverse, double *r, fftw_complex *c) {
  if (n < 2 || (n & (n - 1))) return NULL; /* power of two only (reference uses 4096) */
  fftw_plan p = (fftw_plan) calloc(1, sizeof(*p));
  p->n = n; p->inverse = inverse; p->rbuf = r; p->cbuf = c;
  p->wr = (double *) malloc(sizeof(double) * n / 2);
  p->wi = (double *) malloc(sizeof(double) * n / 2);
  p->xr = (double *) malloc(sizeof(double) * n);
  p->xi = (double *) malloc(sizeof(double) * n);
  for (int k = 0; k < n / 2; k++) {
    double ang = -2.0 * M_PI * k / n;
    p->wr[k] = cos(ang);
    p->wi[k] = sin(ang);
  }
  return p;
}

fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags) {
  (void) flags; return mkplan(n, 0, in, out);
}
fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *in, double *out, unsigned flags) {
  (void) flags; return mkplan(n, 1, out, in);
}

void fftw_execute(const fftw_plan p) {
  int n = p->n;
  if (!p->inverse) {
    for (int i = 0; i < n; i++) { p->xr[i] = p->rbuf[i]; p->xi[i] = 0.0; }
    fft_pow2(n, +1, p->xr, p->xi, p->wr, p->wi);
    for (int k = 0; k <= n / 2; k++) { p->cbuf[k][0] = p->xr[k]; p->cbuf[k][1] = p->xi[k]; }
  } else {
    for (int k = 0; k <= n / 2; k++) { p->xr[k] = p->cbuf[k][0]; p->xi[k] = p->cbuf[k][1]; }
    p->xi[0] = 0.0; p->xi[n / 2] = 0.0;
    for (int k = n / 2 + 1; k < n; k++) { p->xr[k] = p->cbuf[n - k][0]; p->xi[k] = -p->cbuf[n - k][1]; }
    fft_pow2(n, -1, p->xr, p->xi, p->wr, p->wi);
    for (int i = 0; i < n; i++) p->rbuf[i] = p->xr[i];
  }
}

void fftw_destroy_plan(fftw_plan p) {
  if (!p) return;
  free(p->wr); free(p->wi); free(p->xr); free(p->xi); free(p);
}

void fftw_free(void *p) { (void) p; /* the reference frees new[]-allocated buffers through this: leave them */ }
