/* relxill_oracle.h — CPU restatement of the relxill spectrum-evaluation hot path.
 *
 * TEST INFRASTRUCTURE.  This is the checker the CUDA path is compared with; it is
 * never linked, loaded or called by anything under relxill_b200/.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Parity status: PINNED — every stage is checked in tests/test_oracle_vs_ref.py against
 * fresh runs of the unmodified reference (oracle/_ref) on the same synthetic tables, and
 * against the golden vectors in tests/golden/ generated from that reference.
 */
#ifndef RELXILL_ORACLE_H_
#define RELXILL_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NR 1000      /* fine radial grid, reference src/common.h:133 */
#define ORC_NG 40        /* g* grid, src/common.h:71 */
#define ORC_NCONV 4096   /* convolution grid, src/Xillspec.h:36 */
#define ORC_NZMAX 50

/* load tables from `dir` (same files the reference reads); returns 0 on success.
 * num_zones_env: value of RELXILL_NUM_RZONES to emulate (0 = unset). */
int orc_init(const char *dir);
void orc_set_num_zones_env(int n);
void orc_set_returnrad_env(int v); /* RELXILL_RETURNRAD_SWITCH (-1 = unset) */

int orc_num_params(const char *model);
int orc_default_params(const char *model, double *out);

/* whole model, same contract as the XSPEC lmod* entry points (reference
 * src/LocalModel.cpp:143-160): energy[n_flux+1], par[npar], flux[n_flux].
 * For convolution models flux is input and output. returns 0 on success. */
int orc_eval_model(const char *model, const double *energy, int n_flux, const double *par, double *flux);

/* stage probes (same meaning as oracle/ref_probe.cpp) */
int orc_syspar(const char *model, const double *par, double *re, double *gmin, double *gmax, double *emis,
               double *del_emit, double *del_inc, double *trff, double *cosne, double *frac);
int orc_relbase(const char *model, const double *par, const double *ener, int n_ener, double *flux);
int orc_relxill_stages(const char *model, const double *par, double *zone, double *zpar, double *corr,
                       double *normch, double *emis2, double *relflux, double *dist, double *xill,
                       int *n_ener_x, int *n_incl, double *conv, double *total);
void orc_conv_grid(double *ener);
void orc_rebin(const double *ener, double *flu, int n, const double *ener0, const double *flu0, int n0);
void orc_fft_conv(const double *fxill, const double *frel, double *fout);
void orc_nthcomp(const double *ener, int n, double gamma, double kte, double z, double *out);
double orc_kerr_rms(double a);
/* corrected_gshift_fluxboost_factor (src/Relreturn_Corona.cpp:39-83) and interp_lin_2d_float (src/relutility.c:53-58) */
double orc_gshift_fluxboost(double xill_gshift_fac, double g, double gamma);
double orc_lin2d_float(double f1, double f2, float r11, float r12, float r21, float r22);

#ifdef __cplusplus
}
#endif
#endif
