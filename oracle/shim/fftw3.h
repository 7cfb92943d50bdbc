/* fftw3.h shim — declarations only (FFTW3 is not installed here).  The five
 * entry points the reference uses (src/Relbase.cpp:55,153-155,165-167,182)
 * are implemented in oracle/shim/fft_shim.c as an exact fp64 radix-2 DFT with
 * FFTW's conventions: unnormalised, r2c gives n/2+1 half-complex bins, c2r
 * reads n/2+1 bins.  Test infrastructure, not product code. */
#ifndef ORACLE_SHIM_FFTW3_H_
#define ORACLE_SHIM_FFTW3_H_
#ifdef __cplusplus
extern "C" {
#endif
typedef double fftw_complex[2];
typedef struct oracle_fftw_plan_s *fftw_plan;
#define FFTW_ESTIMATE (1U << 6)
fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *in, double *out, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
void fftw_free(void *p);
#ifdef __cplusplus
}
#endif
#endif
