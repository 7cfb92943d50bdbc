/* relxill_oracle.c — CPU restatement of the relxill spectrum-evaluation hot path.
 *
 * TEST INFRASTRUCTURE (see relxill_oracle.h): the checker for the CUDA library, never part
 * of the product path.  Plain sequential C that follows the reference's arithmetic —
 * including its float/double mixing and its quirks — so that it agrees with the unmodified
 * reference (oracle/_ref) to rounding.  Every function names the reference code it restates
 * (paths relative to /root/reference).  Parity status: PINNED against fresh runs of oracle/_ref
 * (tests/test_oracle.py::test_oracle_vs_reference_*, whole models and stage by stage) and against the
 * golden vectors under tests/golden/ that were made with it (14 models, 50-zone cases, rows of the
 * BASELINE configurations 2-5).
 *
 * Not restated (out of scope): relxillBB, the alpha model, extended jet, debug file
 * writers, the caches (result-transparent).
 */
#include "relxill_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../relxill_b200/csrc/minifits.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ constants (src/common.h, src/Xillspec.h) */
enum { EMIS_BKN = 1, EMIS_LP = 2 };
enum { PRIM_NONE = 0, PRIM_ECUT = 1, PRIM_NTHCOMP = 2, PRIM_BB = 3 };   /* src/common.h:57-59 */
enum { XT_STD = 0, XT_CP = 1, XT_NS = 2, XT_CO = 3, XT_COUNT = 4 };
enum { T_LINE, T_CONV, T_XILL, T_RELXILL };
enum { ION_CONST = 0, ION_PL = 1, ION_ALPHA = 2 };
#define REL_NA 25
#define REL_NMU 30
#define REL_NRT 100
#define LP_NA 20
#define LP_NH 250
#define LP_NRT 100
#define RR_NR 50
#define RR_NG 20
#define N_COARSE 500
#define GFAC_H 5e-3
#define XP_GAM 0
#define XP_AFE 1
#define XP_LXI 2
#define XP_ECT 3
#define XP_DNS 4
#define XP_KTB 5
#define XP_FRA 6
#define XP_INC 7

/* ------------------------------------------------------------------ small utilities (src/relutility.c) */
static double lin1d(double f, double lo, double hi) { return f * hi + (1.0 - f) * lo; } /* :24-26 */

static int bsearch_f(const float *arr, int n, float val) { /* :135-151 */
  if (n <= 1) return -1;
  int klo = 0, khi = n - 1;
  while (khi - klo > 1) {
    int k = (khi + klo) / 2;
    if (arr[k] > val) khi = k; else klo = k;
  }
  return klo;
}
static int bsearch_d(const double *arr, int n, double val) { /* :155-171 */
  if (n <= 1) return -1;
  int klo = 0, khi = n - 1;
  while (khi - klo > 1) {
    int k = (khi + klo) / 2;
    if (arr[k] > val) khi = k; else klo = k;
  }
  return klo;
}
static int inv_bsearch_d(const double *arr, int n, double val) { /* :195-211, descending array */
  if (n <= 1) return -1;
  int klo = 0, khi = n - 1;
  while (khi - klo > 1) {
    int k = (khi + klo) / 2;
    if (arr[k] < val) khi = k; else klo = k;
  }
  return klo;
}
static void log_grid(double *e, int n, double emin, double emax) { /* :399-405 */
  for (int i = 0; i < n; i++) {
    e[i] = 1.0 * i / (n - 1) * (log(emax) - log(emin)) + log(emin);
    e[i] = exp(e[i]);
  }
}
static double trapez_single(const double *re, int i, int nr) { /* :233-244 */
  double dr;
  if (i == 0) dr = 0.5 * (re[i] - re[i + 1]);
  else if (i == nr - 1) dr = 0.5 * (re[i - 1] - re[i]);
  else dr = 0.5 * (re[i - 1] - re[i + 1]);
  return re[i] * dr * M_PI;
}
static double trapez_single_asc(const double *re, int i, int nr) { /* :246-257 */
  double dr;
  if (i == 0) dr = 0.5 * (re[i + 1] - re[i]);
  else if (i == nr - 1) dr = 0.5 * (re[i] - re[i - 1]);
  else dr = 0.5 * (re[i + 1] - re[i - 1]);
  return re[i] * dr * M_PI;
}

/* flux-conserving rebin, src/relutility.c:549-601 (cursor semantics kept) */
void orc_rebin(const double *ener, double *flu, int nbins, const double *ener0, const double *flu0, int nbins0) {
  int imin = 0, imax = 0;
  for (int ii = 0; ii < nbins; ii++) {
    flu[ii] = 0.0;
    if ((ener0[0] <= ener[ii + 1]) && (ener0[nbins0] >= ener[ii])) {
      while (ener0[imin] <= ener[ii] && imin <= nbins0) imin++;
      if (imin > 0) imin--;
      while (ener0[imax] <= ener[ii + 1] && imax < nbins0) imax++;
      if (imax > 0) imax--;
      double elo = ener[ii], ehi = ener[ii + 1];
      if (elo < ener0[imin]) elo = ener0[imin];
      if (ehi > ener0[imax + 1]) ehi = ener0[imax + 1];
      if (imax == imin) {
        flu[ii] = (ehi - elo) / (ener0[imin + 1] - ener0[imin]) * flu0[imin];
      } else {
        double dmin = (ener0[imin + 1] - elo) / (ener0[imin + 1] - ener0[imin]);
        double dmax = (ehi - ener0[imax]) / (ener0[imax + 1] - ener0[imax]);
        flu[ii] += flu0[imin] * dmin + flu0[imax] * dmax;
        for (int jj = imin + 1; jj <= imax - 1; jj++) flu[ii] += flu0[jj];
      }
    }
  }
}

/* mean of a descending-x profile at ascending points, src/relutility.c:636-663; returns 0 ok */
static int inv_rebin_mean(const double *x0, const double *y0, int n0, const double *xn, double *yn, int nn) {
  if (xn[0] > xn[nn - 1] || x0[nn - 1] > x0[0]) return 1;
  if (xn[0] < x0[n0 - 1] || xn[nn - 1] > x0[0]) return 1;
  int in = nn - 1;
  for (int ii = 0; ii < n0 - 1; ii++) {
    if (x0[ii] > xn[in] && x0[ii + 1] <= xn[in]) {
      double f = (xn[in] - x0[ii + 1]) / (x0[ii] - x0[ii + 1]);
      yn[in] = lin1d(f, y0[ii + 1], y0[ii]);
      in--;
      if (in < 0) break;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ Kerr formulas (src/Relphysics.cpp) */
double orc_kerr_rms(double a) { /* :139-150 */
  double sign = (a < 0) ? -1.0 : 1.0;
  double Z1 = 1.0 + pow(1.0 - a * a, 1.0 / 3.0) * (pow(1.0 + a, 1.0 / 3.0) + pow(1.0 - a, 1.0 / 3.0));
  double Z2 = sqrt((3.0 * a * a) + (Z1 * Z1));
  return 3.0 + Z2 - sign * sqrt((3.0 - Z1) * (3.0 + Z1 + (2 * Z2)));
}
static double kerr_rplus(double a) { return 1 + sqrt(1 - a * a); } /* :153-155 */
static double relat_abberation(double del, double beta) { /* :127-129 */
  return acos((cos(del) - beta) / (1 - beta * cos(del)));
}
static double doppler_factor(double del, double bet) { return sqrt(1.0 - bet * bet) / (1.0 + bet * cos(del)); }
static double gi_potential_lp(double r, double a, double h, double bet, double del) { /* :163-207 */
  double ut_d = ((r * sqrt(r) + a) / (sqrt(r) * sqrt(r * r - 3 * r + 2 * a * sqrt(r))));
  double ut_h = sqrt((h * h + a * a) / (h * h - 2 * h + a * a));
  double gi = ut_d / ut_h;
  if (fabs(bet) < 1e-6) return gi;
  double gam = 1.0 / sqrt(1.0 - bet * bet);
  double sign = (del > M_PI / 2) ? -1.0 : 1.0;
  double delta_eq = h * h - 2 * h + a * a;
  double q2 = (pow(sin(del), 2)) * (pow((h * h + a * a), 2) / delta_eq) - a * a;
  double beta_fac = sqrt(pow((h * h + a * a), 2) - delta_eq * (q2 + a * a));
  beta_fac = gam * (1.0 + sign * beta_fac / (h * h + a * a) * bet);
  return gi / beta_fac;
}
static double density_ss73_zone_a(double radius, double rms) { /* :123-125 */
  return pow((radius / rms), (3. / 2)) * pow((1 - sqrt(rms / radius)), -2);
}

/* ------------------------------------------------------------------ parameters */
typedef struct {
  int model_type, emis_type, prim_type, type;
  double a, incl, emis1, emis2, rbr, rin, rout, lineE, z, height, gamma, beta;
  int limb, num_zones, return_rad, ion_grad_type;
  /* xillver side */
  double gam, afe, lxi, ect, dens, refl_frac, iongrad_index, xincl;
  double ktbb, frac_pl_bb;   /* xillverNS / xillverCO tables */
  int boost;
  int xtab;                  /* xillver table: XT_STD, XT_CP, XT_NS, XT_CO (get_xilltable_id, src/xilltable.c:682-694) */
  /* returning-radiation correction factors (NULL = none), on the zone grid */
  const double *corr_rgrid, *corr_flux, *corr_gshift;
  int corr_nz;
} Par;

enum {
  P_LINEE, P_INDEX1, P_INDEX2, P_RBR, P_A, P_RIN, P_ROUT, P_INCL, P_Z, P_LIMB, P_GAMMA, P_LOGXI, P_LOGN, P_AFE,
  P_ECUT, P_KTE, P_REFLFRAC, P_H, P_BETA, P_IONIDX, P_IONTYPE, P_SWRET, P_SWBOOST, P_KTBB, P_ACO, P_FRAC, P_COUNT
};
typedef struct {
  const char *name;
  int type, irrad, prim, model_type, npar;
  int ids[20];
  double def[20];
} ModelDef;

/* parameter order and defaults: src/modelfiles/lmodel_relxill_public.dat; types: src/ModelDatabase.h:136-165,
 * integer model types: src/ModelDefinition.cpp:35-61 */
static const ModelDef MODELS[] = {
    {"relline", T_LINE, EMIS_BKN, PRIM_NONE, 1, 10,
     {P_LINEE, P_INDEX1, P_INDEX2, P_RBR, P_A, P_INCL, P_RIN, P_ROUT, P_Z, P_LIMB},
     {6.4, 3, 3, 15, 0.998, 30, -1, 400, 0, 0}},
    {"relconv", T_CONV, EMIS_BKN, PRIM_NONE, 11, 8,
     {P_INDEX1, P_INDEX2, P_RBR, P_A, P_INCL, P_RIN, P_ROUT, P_LIMB},
     {3, 3, 15, 0.998, 30, -1, 400, 0}},
    {"relline_lp", T_LINE, EMIS_LP, PRIM_NONE, 2, 10,
     {P_LINEE, P_H, P_A, P_INCL, P_RIN, P_ROUT, P_Z, P_LIMB, P_GAMMA, P_SWRET},
     {6.4, 6, 0.998, 30, -1, 400, 0, 0, 2, 1}},
    {"relconv_lp", T_CONV, EMIS_LP, PRIM_NONE, 12, 9,
     {P_H, P_BETA, P_A, P_INCL, P_RIN, P_ROUT, P_LIMB, P_GAMMA, P_SWRET},
     {6, 0, 0.998, 30, -1, 400, 0, 2, 1}},
    {"relxill", T_RELXILL, EMIS_BKN, PRIM_ECUT, -1, 13,
     {P_INDEX1, P_INDEX2, P_RBR, P_A, P_INCL, P_RIN, P_ROUT, P_Z, P_GAMMA, P_LOGXI, P_AFE, P_ECUT, P_REFLFRAC},
     {3, 3, 15, 0.998, 30, -1, 400, 0, 2, 3.1, 1, 300, 3}},
    {"relxilllp", T_RELXILL, EMIS_LP, PRIM_ECUT, -2, 14,
     {P_H, P_BETA, P_A, P_INCL, P_RIN, P_ROUT, P_Z, P_GAMMA, P_LOGXI, P_AFE, P_ECUT, P_REFLFRAC, P_SWRET, P_SWBOOST},
     {6, 0, 0.998, 30, -1, 400, 0, 2, 3.1, 1, 300, 1, 1, 0}},
    {"xillver", T_XILL, 0, PRIM_ECUT, 0, 7,
     {P_GAMMA, P_AFE, P_ECUT, P_LOGXI, P_Z, P_INCL, P_REFLFRAC},
     {2, 1, 300, 3.1, 0, 30, -1}},
    {"xillverCp", T_XILL, 0, PRIM_NTHCOMP, 100, 8,
     {P_GAMMA, P_AFE, P_KTE, P_LOGXI, P_LOGN, P_Z, P_INCL, P_REFLFRAC},
     {2, 1, 60, 3.1, 15, 0, 30, -1}},
    {"relxillCp", T_RELXILL, EMIS_BKN, PRIM_NTHCOMP, -1, 14,
     {P_INCL, P_A, P_RIN, P_ROUT, P_RBR, P_INDEX1, P_INDEX2, P_Z, P_GAMMA, P_LOGXI, P_LOGN, P_AFE, P_KTE, P_REFLFRAC},
     {30, 0.998, -1, 400, 15, 3, 3, 0, 2, 3.1, 15, 1, 60, 3}},
    {"relxilllpCp", T_RELXILL, EMIS_LP, PRIM_NTHCOMP, -2, 17,
     {P_INCL, P_A, P_RIN, P_ROUT, P_H, P_BETA, P_GAMMA, P_LOGXI, P_LOGN, P_AFE, P_KTE, P_REFLFRAC, P_Z, P_IONIDX,
      P_IONTYPE, P_SWRET, P_SWBOOST},
     {30, 0.998, -1, 400, 6, 0, 2, 3.1, 15, 1, 60, 1, 0, 0, 0, 1, 0}},
    /* neutron-star (blackbody-irradiated) and CO flavours: lmodel_relxill_public.dat:131-153, lmodel_relxill_devel.dat:1-25 */
    {"xillverNS", T_XILL, 0, PRIM_BB, -101, 7,
     {P_KTBB, P_AFE, P_LOGN, P_LOGXI, P_Z, P_INCL, P_REFLFRAC},
     {2, 1, 15, 3.1, 0, 30, -1}},
    {"relxillNS", T_RELXILL, EMIS_BKN, PRIM_BB, -30, 13,
     {P_INDEX1, P_INDEX2, P_RBR, P_A, P_INCL, P_RIN, P_ROUT, P_Z, P_KTBB, P_LOGXI, P_AFE, P_LOGN, P_REFLFRAC},
     {3, 3, 15, 0.998, 30, -1, 400, 0, 2, 3.1, 1, 15, 3}},
    {"xillverCO", T_XILL, 0, PRIM_ECUT, -210, 8,
     {P_GAMMA, P_ACO, P_KTBB, P_FRAC, P_ECUT, P_Z, P_INCL, P_REFLFRAC},
     {2, 5, 0.1, 0.01, 300, 0, 45, -1}},
    {"relxillCO", T_RELXILL, EMIS_BKN, PRIM_ECUT, -200, 14,
     {P_INDEX1, P_INDEX2, P_RBR, P_A, P_INCL, P_RIN, P_ROUT, P_Z, P_GAMMA, P_ACO, P_KTBB, P_FRAC, P_ECUT, P_REFLFRAC},
     {3, 3, 15, 0.998, 30, -1, 400, 0, 2, 5, 0.1, 0.01, 300, 3}},
};
#define N_MODELS ((int) (sizeof(MODELS) / sizeof(MODELS[0])))

/* The reference reads its environment switches on every evaluation; so does the oracle.  orc_set_num_zones_env /
 * orc_set_returnrad_env are programmatic overrides for the tests (0 / -1: no override, the environment decides). */
static int g_env_num_zones = 0;
static int g_env_returnrad = -1;
void orc_set_num_zones_env(int n) { g_env_num_zones = n; }
void orc_set_returnrad_env(int v) { g_env_returnrad = v; }
static int env_is_one(const char *name) { /* is_env_set, src/ModelDefinition.cpp:123-137 */
  const char *env = getenv(name);
  return (env != NULL && (int) strtod(env, NULL) == 1) ? 1 : 0;
}
static int env_num_zones(void) { /* src/relutility.c:509-513 */
  if (g_env_num_zones > 0) return g_env_num_zones;
  const char *env = getenv("RELXILL_NUM_RZONES");
  return env ? (int) atof(env) : 0;
}
static int env_returnrad(void) {
  if (g_env_returnrad >= 0) return g_env_returnrad;
  return getenv("RELXILL_RETURNRAD_SWITCH") ? env_is_one("RELXILL_RETURNRAD_SWITCH") : -1;
}

static const ModelDef *find_model(const char *name) {
  for (int i = 0; i < N_MODELS; i++)
    if (strcmp(MODELS[i].name, name) == 0) return &MODELS[i];
  return NULL;
}
int orc_num_params(const char *model) {
  const ModelDef *m = find_model(model);
  return m ? m->npar : -1;
}
int orc_default_params(const char *model, double *out) {
  const ModelDef *m = find_model(model);
  if (!m) return -1;
  for (int i = 0; i < m->npar; i++) out[i] = m->def[i];
  return m->npar;
}

/* zone count, src/relutility.c:506-544 */
static int get_num_zones(int model_type, int emis_type, int ion_grad_type) {
  int env = env_num_zones();
  if (ion_grad_type != ION_CONST) {
    if (env != 0 && env > 9 && env <= ORC_NZMAX) return env;
    return 25;
  } else if (model_type < 0 && emis_type == EMIS_LP) {
    if (env != 0 && env > 0 && env <= ORC_NZMAX) return env;
    return 10;
  }
  return 1;
}

/* src/ModelDefinition.cpp:168-385 (get_rel_params, get_xill_params, check_parameter_bounds) and the Ecut frame
 * change of src/Relxill.cpp:192-202.  returns 0 ok, 1 invalid parameters */
static int interpret_params(const ModelDef *m, const double *par, Par *p) {
  double v[P_COUNT];
  int has[P_COUNT];
  memset(has, 0, sizeof(has));
  for (int i = 0; i < P_COUNT; i++) v[i] = 0.0;
  for (int i = 0; i < m->npar; i++) { v[m->ids[i]] = par[i]; has[m->ids[i]] = 1; }
  memset(p, 0, sizeof(*p));
  p->type = m->type;
  p->model_type = m->model_type;
  p->emis_type = m->irrad;
  p->prim_type = m->prim;
  /* xillver-side values (get_xill_params) */
  const int is_co = (m->model_type == -200 || m->model_type == -210);   /* is_co_model, src/relutility.c:103-109 */
  const int is_ns = (m->model_type == -30 || m->model_type == -101);    /* is_ns_model, :95-101 */
  p->xtab = is_ns ? XT_NS : is_co ? XT_CO : (m->prim == PRIM_NTHCOMP) ? XT_CP : XT_STD;
  p->afe = is_co ? v[P_ACO] : v[P_AFE];                                  /* src/ModelDefinition.cpp:347-349 */
  p->ktbb = v[P_KTBB];
  p->frac_pl_bb = v[P_FRAC];
  p->xincl = v[P_INCL];
  p->ect = (m->prim == PRIM_NTHCOMP) ? (has[P_KTE] ? v[P_KTE] : 0.0) : (has[P_ECUT] ? v[P_ECUT] : 300.0);
  p->lxi = has[P_LOGXI] ? v[P_LOGXI] : 0.0;
  p->dens = has[P_LOGN] ? v[P_LOGN] : (is_co ? 17.0 : 15.0);           /* :362-363 */
  p->iongrad_index = v[P_IONIDX];
  p->gam = v[P_GAMMA];
  p->refl_frac = v[P_REFLFRAC];
  p->boost = (int) lround(has[P_SWBOOST] ? v[P_SWBOOST] : 0.0);
  p->z = v[P_Z];
  if (m->type == T_XILL) return 0;

  p->a = v[P_A];
  p->incl = v[P_INCL] * M_PI / 180;
  p->rin = v[P_RIN];
  p->rout = v[P_ROUT];
  p->emis1 = v[P_INDEX1];
  p->emis2 = v[P_INDEX2];
  p->rbr = v[P_RBR];
  p->lineE = v[P_LINEE];
  p->gamma = v[P_GAMMA];
  p->height = v[P_H];
  p->beta = v[P_BETA];
  p->limb = (int) lround(v[P_LIMB]);
  { /* get_returnrad_switch :139-149 */
    int def = (m->irrad == EMIS_LP) ? 1 : 0;
    int sw = (env_returnrad() == 1) ? 1 : def;
    p->return_rad = (int) lround(has[P_SWRET] ? v[P_SWRET] : (double) sw);
  }
  /* check_parameter_bounds :168-273 */
  if (p->rin < 0) p->rin = -1.0 * p->rin * orc_kerr_rms(p->a);
  if (p->rout < 0) p->rout = -1.0 * p->rout * orc_kerr_rms(p->a);
  if (p->rbr < 0) p->rbr = -1.0 * p->rbr * orc_kerr_rms(p->a);
  int bad = 0;
  if (p->rout <= p->rin) bad = 1;
  double rms = orc_kerr_rms(p->a);
  if (p->rin < rms) p->rin = rms;
  if (p->a > 0.9982) return 1;
  if (p->a < -1) return 1;
  if (p->incl < 3 * M_PI / 180 || p->incl > 87 * M_PI / 180) return 1;
  if (p->rout <= p->rin) return 1;
  if (p->rout > 1000.0) p->rout = 1000.0;
  if (p->emis_type == EMIS_BKN) {
    if (p->rbr < p->rin) p->rbr = p->rin;
    if (p->rbr > p->rout) p->rbr = p->rout;
  }
  if (p->emis_type == EMIS_LP) {
    if (p->beta < 0) p->beta = 0.0;
    if (p->beta > 0.99) p->beta = 0.99;
    if (p->height < 0) p->height = -1.0 * p->height * kerr_rplus(p->a);
    double h_fac = 1.1, r_event = kerr_rplus(p->a);
    if ((h_fac * r_event - p->height) > 1e-4) p->height = r_event * h_fac;
  }
  if (bad) return 1;
  p->ion_grad_type = (int) lround(has[P_IONTYPE] ? v[P_IONTYPE] : 0.0);
  p->num_zones = get_num_zones(p->model_type, p->emis_type, p->ion_grad_type);
  return 0;
}

static double energy_shift_source_obs(const Par *p) { /* src/Relphysics.cpp:239-255 */
  if (p->emis_type != EMIS_LP) return 1;
  double g_inf_0 = sqrt(1.0 - (2 * p->height / (p->height * p->height + p->a * p->a)));
  if (p->beta < 1e-4) return g_inf_0;
  return g_inf_0 * doppler_factor(M_PI - p->incl, p->beta);
}

/* ------------------------------------------------------------------ tables */
typedef struct {
  float a[REL_NA], mu0[REL_NMU];
  float *r, *gmin, *gmax;                  /* [na][nmu][100] */
  float *trff1, *trff2, *cosne1, *cosne2;  /* [na][nmu][100][40] */
} RelTab;
typedef struct {
  float a[LP_NA];
  float *h;                       /* [na][250] */
  float *rad;                     /* [na][100] */
  float *intens, *del, *del_inc;  /* [na][250][100] */
} LpTab;
typedef struct {
  int npar, nvals[6], pindex[6], n_ener, n_incl;
  float *vals[6];
  float *elo, *ehi, *incl;
  float *data; /* [nrows][n_ener], renormalised as the reference does on load */
  long nrows;
} XillTab;
typedef struct {
  int nspin;
  double *spin;
  double *rlo, *rhi;         /* [nspin][50] */
  double *tf_r, *gmin, *gmax; /* [nspin][50][50] */
  double *frac_g;            /* [nspin][50][50][20] */
} RradTab;

static char g_dir[1024];
static RelTab *g_rel = NULL;
static LpTab *g_lp = NULL;
static XillTab *g_xill[XT_COUNT]; /* index by table id */
static RradTab *g_rr = NULL;
static double g_econv[ORC_NCONV + 1], g_ecoarse[N_COARSE + 1];

static mf_file *open_tab(const char *name) {
  char path[1200];
  snprintf(path, sizeof(path), "%s/%s", g_dir, name);
  mf_file *f = mf_open(path);
  if (!f) fprintf(stderr, "oracle: cannot open %s\n", path);
  return f;
}

static int load_rel(void) { /* layout: src/reltable.c:171-311 */
  mf_file *f = open_tab("rel_table_v0.5a.fits");
  if (!f) return 1;
  RelTab *t = (RelTab *) calloc(1, sizeof(RelTab));
  int h = mf_find_hdu(f, "a");
  mf_read(&f->hdus[h - 1], mf_find_col(&f->hdus[h - 1], "a"), 1, 1, REL_NA, 'f', t->a);
  h = mf_find_hdu(f, "mu0");
  mf_read(&f->hdus[h - 1], mf_find_col(&f->hdus[h - 1], "mu0"), 1, 1, REL_NMU, 'f', t->mu0);
  size_t n1 = (size_t) REL_NA * REL_NMU * REL_NRT, n2 = n1 * ORC_NG;
  t->r = malloc(4 * n1); t->gmin = malloc(4 * n1); t->gmax = malloc(4 * n1);
  t->trff1 = malloc(4 * n2); t->trff2 = malloc(4 * n2); t->cosne1 = malloc(4 * n2); t->cosne2 = malloc(4 * n2);
  for (int ia = 0; ia < REL_NA; ia++)
    for (int im = 0; im < REL_NMU; im++) {
      const mf_hdu *hd = &f->hdus[ia * REL_NMU + im + 4 - 1];
      size_t o1 = ((size_t) ia * REL_NMU + im) * REL_NRT, o2 = o1 * ORC_NG;
      mf_read(hd, mf_find_col(hd, "r"), 1, 1, REL_NRT, 'f', t->r + o1);
      mf_read(hd, mf_find_col(hd, "gmin"), 1, 1, REL_NRT, 'f', t->gmin + o1);
      mf_read(hd, mf_find_col(hd, "gmax"), 1, 1, REL_NRT, 'f', t->gmax + o1);
      mf_read(hd, mf_find_col(hd, "trff1"), 1, 1, REL_NRT * ORC_NG, 'f', t->trff1 + o2);
      mf_read(hd, mf_find_col(hd, "trff2"), 1, 1, REL_NRT * ORC_NG, 'f', t->trff2 + o2);
      mf_read(hd, mf_find_col(hd, "cosne1"), 1, 1, REL_NRT * ORC_NG, 'f', t->cosne1 + o2);
      mf_read(hd, mf_find_col(hd, "cosne2"), 1, 1, REL_NRT * ORC_NG, 'f', t->cosne2 + o2);
    }
  mf_close(f);
  g_rel = t;
  return 0;
}

static int load_lp(void) { /* layout: src/reltable.c:313-448 */
  mf_file *f = open_tab("rel_lp_table_v0.5b.fits");
  if (!f) return 1;
  LpTab *t = (LpTab *) calloc(1, sizeof(LpTab));
  const mf_hdu *hd = &f->hdus[mf_find_hdu(f, "I_h") - 1];
  mf_read(hd, mf_find_col(hd, "a"), 1, 1, LP_NA, 'f', t->a);
  t->h = malloc(4 * LP_NA * LP_NH);
  t->rad = malloc(4 * LP_NA * LP_NRT);
  size_t n3 = (size_t) LP_NA * LP_NH * LP_NRT;
  t->intens = malloc(4 * n3); t->del = malloc(4 * n3); t->del_inc = malloc(4 * n3);
  for (int ia = 0; ia < LP_NA; ia++) {
    mf_read(hd, mf_find_col(hd, "hgrid"), ia + 1, 1, LP_NH, 'f', t->h + ia * LP_NH);
    mf_read(hd, mf_find_col(hd, "r"), ia + 1, 1, LP_NRT, 'f', t->rad + ia * LP_NRT);
    for (int ih = 0; ih < LP_NH; ih++) {
      char nm[32];
      size_t o = ((size_t) ia * LP_NH + ih) * LP_NRT;
      snprintf(nm, sizeof(nm), "h%i", ih + 1);
      mf_read(hd, mf_find_col(hd, nm), ia + 1, 1, LP_NRT, 'f', t->intens + o);
      snprintf(nm, sizeof(nm), "del%i", ih + 1);
      mf_read(hd, mf_find_col(hd, nm), ia + 1, 1, LP_NRT, 'f', t->del + o);
      snprintf(nm, sizeof(nm), "del_inc%i", ih + 1);
      mf_read(hd, mf_find_col(hd, nm), ia + 1, 1, LP_NRT, 'f', t->del_inc + o);
      for (int k = 0; k < LP_NRT; k++) { /* :370-377 */
        t->del[o + k] = fabsf(t->del[o + k]);
        t->del_inc[o + k] = fabsf(t->del_inc[o + k]);
      }
    }
  }
  mf_close(f);
  g_lp = t;
  return 0;
}

static int xill_param_id(const char *name) { /* src/common.h:141-161 */
  if (!strcmp(name, "Gamma")) return XP_GAM;
  if (!strcmp(name, "A_Fe")) return XP_AFE;
  if (!strcmp(name, "logXi")) return XP_LXI;
  if (!strcmp(name, "Ecut") || !strcmp(name, "kTe")) return XP_ECT;
  if (!strcmp(name, "Dens")) return XP_DNS;
  if (!strcmp(name, "kTbb")) return XP_KTB;
  if (!strcmp(name, "A_CO")) return XP_AFE;   /* PARAM_ACO == PARAM_AFE */
  if (!strcmp(name, "Frac")) return XP_FRA;
  if (!strcmp(name, "Incl")) return XP_INC;
  return -1;
}

/* src/xilltable.c:169-276 (axes), :513-564 (rows; the reference loads them lazily — we load all of them
 * eagerly) and :478-511 (renormalisation at load, with its two float roundings) */
static int load_xill(int xtab) {
  static const char *const names[XT_COUNT] = {"xillver-a-Ec5.fits", "xillverCp_v3.4.fits", "xillverNS-2.fits", "xillverCO.fits"};
  mf_file *f = open_tab(names[xtab]);
  if (!f) return 1;
  XillTab *t = (XillTab *) calloc(1, sizeof(XillTab));
  const mf_hdu *hp = &f->hdus[mf_find_hdu(f, "PARAMETERS") - 1];
  t->npar = (int) hp->nrows;
  mf_read(hp, 9, 1, 1, t->npar, 'i', t->nvals);
  t->nrows = 1;
  for (int i = 0; i < t->npar; i++) {
    char nm[16];
    mf_read_str(hp, 1, i + 1, nm, 8);
    t->pindex[i] = xill_param_id(nm);
    t->vals[i] = malloc(4 * t->nvals[i]);
    mf_read(hp, 10, i + 1, 1, t->nvals[i], 'f', t->vals[i]);
    t->nrows *= t->nvals[i];
  }
  t->incl = t->vals[t->npar - 1];
  t->n_incl = t->nvals[t->npar - 1];
  const mf_hdu *he = &f->hdus[mf_find_hdu(f, "ENERGIES") - 1];
  t->n_ener = (int) he->nrows;
  t->elo = malloc(4 * t->n_ener); t->ehi = malloc(4 * t->n_ener);
  mf_read(he, 1, 1, 1, t->n_ener, 'f', t->elo);
  mf_read(he, 2, 1, 1, t->n_ener, 'f', t->ehi);
  const mf_hdu *hs = &f->hdus[mf_find_hdu(f, "SPECTRA") - 1];
  t->data = malloc(4 * (size_t) t->nrows * t->n_ener);
  int ax_lxi = -1, ax_dns = -1;
  for (int i = 0; i < t->npar; i++) {
    if (t->pindex[i] == XP_LXI) ax_lxi = i;
    if (t->pindex[i] == XP_DNS) ax_dns = i;
  }
  for (long row = 0; row < t->nrows; row++) {
    float *spec = t->data + (size_t) row * t->n_ener;
    mf_read(hs, 2, row + 1, 1, t->n_ener, 'f', spec);
    long rem = row;
    int idx[6];
    for (int i = t->npar - 1; i >= 0; i--) { idx[i] = (int) (rem % t->nvals[i]); rem /= t->nvals[i]; }
    /* axes the table does not have take the model's fixed value (getDefaultLogxi / getDefaultDensity,
     * src/xilltable.c:566-573): logxi 0; logN 17 for the CO table, 15 otherwise (src/ModelDefinition.cpp:361-363) */
    double lxi = (ax_lxi >= 0) ? t->vals[ax_lxi][idx[ax_lxi]] : 0.0;
    double dens = (ax_dns >= 0) ? t->vals[ax_dns][idx[ax_dns]] : (xtab == XT_CO ? 17.0 : 15.0);
    for (int k = 0; k < t->n_ener; k++) {
      spec[k] /= pow(10, lxi);
      if (fabs(dens - 15) > 1e-6) spec[k] /= pow(10, dens - 15);
    }
  }
  mf_close(f);
  g_xill[xtab] = t;
  return 0;
}

static int load_rrad(void) { /* layout: src/Relreturn_Table.cpp:289-361 */
  mf_file *f = open_tab("table_returnRad_v20220301.fits");
  if (!f) return 1;
  RradTab *t = (RradTab *) calloc(1, sizeof(RradTab));
  const mf_hdu *hs = &f->hdus[mf_find_hdu(f, "SPIN") - 1];
  t->nspin = (int) hs->nrows;
  t->spin = malloc(8 * t->nspin);
  mf_read(hs, mf_find_col(hs, "a"), 1, 1, t->nspin, 'd', t->spin);
  t->rlo = malloc(8 * t->nspin * RR_NR); t->rhi = malloc(8 * t->nspin * RR_NR);
  size_t n2 = (size_t) RR_NR * RR_NR;
  t->tf_r = malloc(8 * t->nspin * n2); t->gmin = malloc(8 * t->nspin * n2); t->gmax = malloc(8 * t->nspin * n2);
  t->frac_g = malloc(8 * t->nspin * n2 * RR_NG);
  for (int s = 0; s < t->nspin; s++) {
    char nm[32];
    snprintf(nm, sizeof(nm), "FRAC%02i", s + 1);
    const mf_hdu *hd = &f->hdus[mf_find_hdu(f, nm) - 1];
    mf_read(hd, mf_find_col(hd, "rlo"), 1, 1, RR_NR, 'd', t->rlo + s * RR_NR);
    mf_read(hd, mf_find_col(hd, "rhi"), 1, 1, RR_NR, 'd', t->rhi + s * RR_NR);
    mf_read(hd, mf_find_col(hd, "tf_r"), 1, 1, n2, 'd', t->tf_r + s * n2);
    mf_read(hd, mf_find_col(hd, "gmin"), 1, 1, n2, 'd', t->gmin + s * n2);
    mf_read(hd, mf_find_col(hd, "gmax"), 1, 1, n2, 'd', t->gmax + s * n2);
    mf_read(hd, mf_find_col(hd, "frac_g"), 1, 1, n2 * RR_NG, 'd', t->frac_g + s * n2 * RR_NG);
  }
  mf_close(f);
  g_rr = t;
  return 0;
}

int orc_init(const char *dir) {
  strncpy(g_dir, dir, sizeof(g_dir) - 1);
  log_grid(g_econv, ORC_NCONV + 1, 0.00035, 2000.0); /* src/Xillspec.cpp:44-54 */
  log_grid(g_ecoarse, N_COARSE + 1, 0.1, 1000.0);    /* src/Xillspec.cpp:166-177 */
  return 0;
}
void orc_conv_grid(double *ener) { memcpy(ener, g_econv, sizeof(g_econv)); }

/* ------------------------------------------------------------------ system parameters */
typedef struct {
  double re[ORC_NR], gmin[ORC_NR], gmax[ORC_NR];
  double *trff, *cosne; /* [nr][ng][2] */
  double emis[ORC_NR], del_emit[ORC_NR], del_inc[ORC_NR];
  double refl_frac, f_bh, f_ad, f_inf, f_inf_rest;
  int limb;
} SysPar;

static double lin2d_f(double f1, double f2, float r11, float r12, float r21, float r22) { /* relutility.c:53-58 */
  return (1.0 - f1) * (1.0 - f2) * r11 + (f1) * (1.0 - f2) * r12 + (1.0 - f1) * (f2) * r21 + (f1) * (f2) * r22;
}

/* src/Relprofile.cpp:141-303 */
static int interpol_reltable(double a, double incl, double rin, double rout, SysPar *sp) {
  if (!g_rel && load_rel()) return 1;
  const RelTab *t = g_rel;
  double mu0 = cos(incl);
  int ia = bsearch_f(t->a, REL_NA, (float) a);
  int im = bsearch_f(t->mu0, REL_NMU, (float) mu0);
  float ifac_a = ((float) a - t->a[ia]) / (t->a[ia + 1] - t->a[ia]);
  float ifac_mu = ((float) mu0 - t->mu0[im]) / (t->mu0[im + 1] - t->mu0[im]);
  size_t o[4] = {((size_t) ia * REL_NMU + im) * REL_NRT, ((size_t) (ia + 1) * REL_NMU + im) * REL_NRT,
                 ((size_t) ia * REL_NMU + im + 1) * REL_NRT, ((size_t) (ia + 1) * REL_NMU + im + 1) * REL_NRT};
  double rt[REL_NRT], gmin_t[REL_NRT], gmax_t[REL_NRT];
  static double trff_t[REL_NRT][ORC_NG][2], cosne_t[REL_NRT][ORC_NG][2];
  for (int i = 0; i < REL_NRT; i++) rt[i] = lin1d(ifac_a, t->r[o[0] + i], t->r[o[1] + i]);
  double rms = orc_kerr_rms(a);
  if ((rt[REL_NRT - 1] > rms) && ((rt[REL_NRT - 1] - rms) / rt[REL_NRT - 1] < 1e-3)) rt[REL_NRT - 1] = rms;
  int ind_rmin = inv_bsearch_d(rt, REL_NRT, rin);
  int ind_rmax = inv_bsearch_d(rt, REL_NRT, rout);
  for (int i = 0; i < REL_NRT; i++) { /* all radii are interpolated (the guard at :211 is always true) */
    gmin_t[i] = lin2d_f(ifac_a, ifac_mu, t->gmin[o[0] + i], t->gmin[o[1] + i], t->gmin[o[2] + i], t->gmin[o[3] + i]);
    gmax_t[i] = lin2d_f(ifac_a, ifac_mu, t->gmax[o[0] + i], t->gmax[o[1] + i], t->gmax[o[2] + i], t->gmax[o[3] + i]);
    for (int j = 0; j < ORC_NG; j++) {
      size_t q = (size_t) i * ORC_NG + j;
#define Q4(arr) arr[o[0] * ORC_NG + q], arr[o[1] * ORC_NG + q], arr[o[2] * ORC_NG + q], arr[o[3] * ORC_NG + q]
      trff_t[i][j][0] = lin2d_f(ifac_a, ifac_mu, Q4(t->trff1));
      trff_t[i][j][1] = lin2d_f(ifac_a, ifac_mu, Q4(t->trff2));
      cosne_t[i][j][0] = lin2d_f(ifac_a, ifac_mu, Q4(t->cosne1));
      cosne_t[i][j][1] = lin2d_f(ifac_a, ifac_mu, Q4(t->cosne2));
#undef Q4
    }
  }
  { /* fine grid, relutility.c:666-677 */
    double r1 = 1.0 / sqrt(rout), r2 = 1.0 / sqrt(rin);
    for (int i = 0; i < ORC_NR; i++) {
      sp->re[i] = ((double) (i)) * (r2 - r1) / (ORC_NR - 1) + r1;
      sp->re[i] = pow(1.0 / (sp->re[i]), 2);
    }
  }
  if (rt[ind_rmax] < 1000.0 && rt[ind_rmax] * 1.01 > 1000.0) rt[ind_rmax] = 1000.0;
  int it = ind_rmin;
  for (int i = ORC_NR - 1; i >= 0; i--) {
    while (sp->re[i] >= rt[it]) {
      it--;
      if (it < 0) {
        if (sp->re[i] - 1000.0 <= 1e-6) { it = 0; break; }
        return 2;
      }
    }
    double fr = (sp->re[i] - rt[it + 1]) / (rt[it] - rt[it + 1]);
    if (fr > 1.0 && it > 0) return 2;
    for (int j = 0; j < ORC_NG; j++)
      for (int k = 0; k < 2; k++) {
        sp->trff[((size_t) i * ORC_NG + j) * 2 + k] = lin1d(fr, trff_t[it + 1][j][k], trff_t[it][j][k]);
        sp->cosne[((size_t) i * ORC_NG + j) * 2 + k] = lin1d(fr, cosne_t[it + 1][j][k], cosne_t[it][j][k]);
      }
    sp->gmin[i] = lin1d(fr, gmin_t[it + 1], gmin_t[it]);
    sp->gmax[i] = lin1d(fr, gmax_t[it + 1], gmax_t[it]);
  }
  return 0;
}

/* src/Rellp.cpp:91-109 */
static void norm_emis_profile(const double *re, int nr, double *emis) {
  double integ = 0.0;
  for (int i = 0; i < nr; i++) {
    double da = (re[1] < re[0]) ? trapez_single(re, i, nr) * 2 : trapez_single_asc(re, i, nr) * 2;
    integ += emis[i] * da;
  }
  for (int i = 0; i < nr; i++) emis[i] /= integ;
}

static void ipol_factor_f(float value, const float *arr, int n, int *ind, double *ifac) { /* relutility.c:419-423 */
  *ind = bsearch_f(arr, n, value);
  *ifac = (value - arr[*ind]) / (arr[*ind + 1] - arr[*ind]);
}

/* log/linear radial re-grid of an ascending-radius profile onto the descending fine grid, src/Rellp.cpp:120-175 */
static int rebin_emis_on_grid(const double *re, int nr, double *emis, double *del_emit, double *del_inc,
                              const double *rt, int nt, const double *et, const double *det, const double *dit) {
  int kk = bsearch_d(rt, nt, re[nr - 1]);
  for (int i = nr - 1; i >= 0; i--) {
    while (re[i] >= rt[kk + 1]) {
      kk++;
      if (kk >= nt - 1) {
        if (re[i] - 1000.0 <= 1e-6) { kk = nt - 2; break; }
        return 1;
      }
    }
    double f;
    if (det[kk] / M_PI * 180.0 <= 75.0) f = (re[i] - rt[kk]) / (rt[kk + 1] - rt[kk]); /* relutility.c:407-417 */
    else f = (log(re[i]) - log(rt[kk])) / (log(rt[kk + 1]) - log(rt[kk]));
    emis[i] = exp(f * log(et[kk + 1]) + (1.0 - f) * log(et[kk]));
    del_emit[i] = lin1d(f, det[kk], det[kk + 1]);
    del_inc[i] = lin1d(f, dit[kk], dit[kk + 1]);
  }
  return 0;
}

/* lamp-post emissivity: src/Rellp.cpp:36-89 (refl. fraction), :190-285 */
static int emis_lamp_post(const Par *p, SysPar *sp) {
  if (!g_lp && load_lp()) return 1;
  const LpTab *t = g_lp;
  int ia;
  double fa;
  ipol_factor_f((float) p->a, t->a, LP_NA, &ia, &fa);
  double rad[LP_NRT], et[LP_NRT], det[LP_NRT], dit[LP_NRT];
  for (int i = 0; i < LP_NRT; i++) rad[i] = lin1d(fa, t->rad[ia * LP_NRT + i], t->rad[(ia + 1) * LP_NRT + i]);
  int ih[2];
  double fh[2];
  for (int s = 0; s < 2; s++) ipol_factor_f((float) p->height, t->h + (ia + s) * LP_NH, LP_NH, &ih[s], &fh[s]);
  size_t o0 = ((size_t) ia * LP_NH + ih[0]) * LP_NRT, o1 = ((size_t) (ia + 1) * LP_NH + ih[1]) * LP_NRT;
  for (int i = 0; i < LP_NRT; i++) {
    et[i] = (1.0 - fa) * lin1d(fh[0], t->intens[o0 + i], t->intens[o0 + LP_NRT + i])
            + (fa) * lin1d(fh[1], t->intens[o1 + i], t->intens[o1 + LP_NRT + i]);
    det[i] = (1.0 - fa) * lin1d(fh[0], t->del[o0 + i], t->del[o0 + LP_NRT + i])
             + (fa) * lin1d(fh[1], t->del[o1 + i], t->del[o1 + LP_NRT + i]);
    dit[i] = (1.0 - fa) * lin1d(fh[0], t->del_inc[o0 + i], t->del_inc[o0 + LP_NRT + i])
             + (fa) * lin1d(fh[1], t->del_inc[o1 + i], t->del_inc[o1 + LP_NRT + i]);
  }
  if (rebin_emis_on_grid(sp->re, ORC_NR, sp->emis, sp->del_emit, sp->del_inc, rad, LP_NRT, et, det, dit)) return 2;
  { /* calc_refl_frac */
    double del_ad_max = det[LP_NRT - 1];
    double del_bh = sp->del_emit[inv_bsearch_d(sp->re, ORC_NR, p->rin)];
    double del_ad = sp->del_emit[inv_bsearch_d(sp->re, ORC_NR, p->rout)];
    if (del_ad_max < M_PI / 2.0) del_ad_max = M_PI / 2.0;
    if (p->beta > 1e-6) {
      del_bh = relat_abberation(del_bh, -1. * p->beta);
      del_ad = relat_abberation(del_ad, -1. * p->beta);
    }
    sp->f_bh = 0.5 * (1.0 - cos(del_bh));
    sp->f_ad = 0.5 * (cos(del_bh) - cos(del_ad));
    sp->f_inf_rest = 0.5 * (1.0 + cos(del_ad_max));
    if (p->beta > 1e-6) sp->f_inf = 0.5 * (1.0 + cos(relat_abberation(del_ad_max, -1. * p->beta)));
    else sp->f_inf = sp->f_inf_rest;
    sp->refl_frac = sp->f_ad / sp->f_inf;
  }
  for (int i = 0; i < ORC_NR; i++) { /* flux boost source->disk, Relphysics.cpp:301-311 */
    double boost = pow(gi_potential_lp(sp->re[i], p->a, p->height, p->beta, sp->del_emit[i]), p->gamma);
    if (p->beta > 1e-6) boost *= pow(doppler_factor(sp->del_emit[i], p->beta), 2);
    sp->emis[i] *= boost;
  }
  return 0;
}

/* returning radiation, src/Relreturn_Corona.cpp:39-83 */
static double gshift_fluxboost(double xill_gshift_fac, double g, double gamma) {
  double g0 = 2. / 3;
  double corr;
  if (xill_gshift_fac < 1) {
    double a = (xill_gshift_fac / g0 - 1) / (g0 - 1);
    double b = 1 - a;
    corr = (g >= 1) ? 1. / g * (1. / g * a + b) : g * (g * a + b);
  } else {
    double alin = (xill_gshift_fac - 1) / (g0 - 1);
    double blin = 1 - alin;
    corr = (g >= 1) ? (1. / g * alin + blin) : (g * alin + blin);
  }
  double fb = pow(g, gamma) * corr;
  if (g < 1 && fb > 1) fb = 1;
  if (fb < 0) fb = 0;
  return fb;
}

/* exported for the reference's known-answer tests of these two helpers (tests/test_oracle.py) */
double orc_gshift_fluxboost(double xill_gshift_fac, double g, double gamma) { return gshift_fluxboost(xill_gshift_fac, g, gamma); }
double orc_lin2d_float(double f1, double f2, float r11, float r12, float r21, float r22) { return lin2d_f(f1, f2, r11, r12, r21, r22); }

/* src/Relreturn_Corona.cpp:100-171,263-321 with the table handling of src/Relreturn_Table.cpp:396-611 */
static int add_returnrad_emis(const Par *p, SysPar *sp) {
  if (!g_rr && load_rrad()) return 1;
  const RradTab *t = g_rr;
  double rlo_e = sp->re[ORC_NR - 1], rhi_e = sp->re[0];
  /* spin: next table spin >= a (no interpolation) */
  int is = bsearch_d(t->spin, t->nspin, p->a);
  if (t->spin[is] < p->a) is++;
  if (is >= t->nspin) return 2;
  const double *rlo_t = t->rlo + is * RR_NR, *rhi_t = t->rhi + is * RR_NR;
  size_t n2 = (size_t) RR_NR * RR_NR;
  const double *tf_t = t->tf_r + is * n2, *gmin_t = t->gmin + is * n2, *gmax_t = t->gmax + is * n2;
  const double *fg_t = t->frac_g + is * n2 * RR_NG;
  /* radial trimming, allocate_radial_grid */
  int klo = bsearch_d(rlo_t, RR_NR, rlo_e);
  int khi = bsearch_d(rhi_t, RR_NR, rhi_e);
  if (fabs(rhi_e - rhi_t[RR_NR - 1]) < 1e-6) khi = RR_NR - 1; else khi++;
  if (fabs(rlo_e - rlo_t[0]) < 1e-6) klo = 0;
  int nrad = (khi + 1) - klo;
  if (nrad <= 0 || nrad > RR_NR) return 3;
  int irad[RR_NR];
  double rlo[RR_NR], rhi[RR_NR], rad[RR_NR];
  for (int i = 0; i < nrad; i++) {
    irad[i] = klo + i;
    rlo[i] = rlo_t[irad[i]];
    rhi[i] = rhi_t[irad[i]];
  }
  rlo[0] = rlo_e;
  rhi[nrad - 1] = rhi_e;
  for (int i = 0; i < nrad; i++) rad[i] = 0.5 * (rlo[i] + rhi[i]);
  /* tf_r trimmed + area correction of the partially covered edge rings, get_interpolated_tfr */
  static double tfr[RR_NR][RR_NR];
  for (int i = 0; i < nrad; i++)
    for (int j = 0; j < nrad; j++) tfr[i][j] = tf_t[irad[i] * RR_NR + irad[j]];
  double corr_area[2];
  for (int s = 0; s < 2; s++) {
    int idx = s == 0 ? 0 : nrad - 1;
    double rlo_tab = rlo_t[irad[idx]];
    if (irad[idx] == 0 && rlo_tab > orc_kerr_rms(p->a)) rlo_tab = orc_kerr_rms(p->a);
    double rhi_tab = rhi_t[irad[idx]];
    double area_table = 0.5 * (rlo_tab + rhi_tab) * (rhi_tab - rlo_tab);
    double area_model = 0.5 * (rlo[idx] + rhi[idx]) * (rhi[idx] - rlo[idx]);
    corr_area[s] = area_model / area_table;
  }
  for (int i = 0; i < nrad - 1; i++) tfr[i][0] *= corr_area[0];
  for (int i = 0; i < nrad - 2; i++) tfr[i][nrad - 1] *= corr_area[1];
  /* emissivity on the table's radial zones */
  double emis_in[RR_NR];
  for (int i = 0; i < nrad; i++) emis_in[i] = 0.0;
  if (inv_rebin_mean(sp->re, sp->emis, ORC_NR, rad, emis_in, nrad)) return 4;
  /* correction factors re-gridded by zone membership of the ring centre (no interpolation) */
  double cflux[RR_NR], cgsh[RR_NR];
  int have_corr = (p->corr_flux != NULL);
  if (have_corr) {
    for (int i = 0; i < nrad; i++) {
      int ind = bsearch_d(p->corr_rgrid, p->corr_nz + 1, rad[i]);
      if (rad[i] < p->corr_rgrid[0]) ind = 0;
      else if (rad[i] > p->corr_rgrid[p->corr_nz]) ind = p->corr_nz;
      cflux[i] = p->corr_flux[ind];
      cgsh[i] = p->corr_gshift[ind];
    }
  }
  double emis_ret[RR_NR], zero[RR_NR];
  for (int io = 0; io < nrad; io++) {
    double sum = 0.0;
    for (int ie = 0; ie < nrad; ie++) {
      double cg = have_corr ? cgsh[ie] : 1.0;
      size_t q = (size_t) irad[io] * RR_NR + irad[ie];
      double ez = 0.0;
      for (int jj = 0; jj < RR_NG; jj++) {
        double g = ((jj + 0.5) / RR_NG) * (gmax_t[q] - gmin_t[q]) + gmin_t[q];
        double e1 = fg_t[q * RR_NG + jj];
        if (fabs(cg - 1) > 1e-3) e1 *= gshift_fluxboost(cg, g, p->gamma) / g;
        else if (fabs(g - 1) > 1e-3) e1 *= pow(g, p->gamma - 1);
        ez += e1;
      }
      sum += ez * tfr[io][ie] * emis_in[ie];
    }
    emis_ret[io] = sum;
    if (have_corr) emis_ret[io] *= cflux[io];
    zero[io] = 0.0;
  }
  /* back onto the fine grid and add */
  static double er[ORC_NR], d1[ORC_NR], d2[ORC_NR];
  if (nrad < 2) return 5;
  if (rebin_emis_on_grid(sp->re, ORC_NR, er, d1, d2, rad, nrad, emis_ret, zero, zero)) return 6;
  for (int i = 0; i < ORC_NR; i++) {
    if (p->return_rad == 1) sp->emis[i] += er[i];
    else if (p->return_rad == -1 || p->return_rad == 2) sp->emis[i] = er[i];
    else return 7;
  }
  return 0;
}

/* src/Relprofile.cpp:310-358 + src/Rellp.cpp:518-564 */
static int system_parameters(const Par *p, SysPar *sp) {
  int rc = interpol_reltable(p->a, p->incl, p->rin, p->rout, sp);
  if (rc) return rc;
  sp->limb = p->limb;
  sp->refl_frac = sp->f_bh = sp->f_ad = sp->f_inf = sp->f_inf_rest = 0.0;
  if (p->emis_type == EMIS_BKN) { /* Rellp.cpp:421-437 */
    for (int i = 0; i < ORC_NR; i++) {
      double alpha = p->emis1;
      if (sp->re[i] > p->rbr) alpha = p->emis2;
      sp->emis[i] = pow(sp->re[i] / p->rbr, -alpha);
      sp->del_emit[i] = -1.0;
      sp->del_inc[i] = -1.0;
    }
    norm_emis_profile(sp->re, ORC_NR, sp->emis);
  } else {
    rc = emis_lamp_post(p, sp);
    if (rc) return 10 + rc;
  }
  if (abs(p->return_rad) > 1e-6) {
    rc = add_returnrad_emis(p, sp);
    if (rc) return 20 + rc;
  }
  return 0;
}

static SysPar *new_syspar(void) {
  SysPar *sp = (SysPar *) calloc(1, sizeof(SysPar));
  sp->trff = (double *) malloc(sizeof(double) * ORC_NR * ORC_NG * 2);
  sp->cosne = (double *) malloc(sizeof(double) * ORC_NR * ORC_NG * 2);
  return sp;
}
static void free_syspar(SysPar *sp) {
  if (!sp) return;
  free(sp->trff); free(sp->cosne); free(sp);
}

/* ------------------------------------------------------------------ relline profile (src/Relprofile.cpp:489-958) */
typedef struct {
  double re, gmin, gmax, del_g, emis;
  const double *trff, *cosne; /* [ng][2] of this radius */
  int limb, save_g_ind;
  double gstar[ORC_NG];
} Relb;

static double relb_func(double eg, int k, Relb *s) { /* :489-521 */
  double egstar = (eg - s->gmin) * s->del_g;
  if (!((egstar >= s->gstar[s->save_g_ind]) && (egstar < s->gstar[s->save_g_ind + 1])))
    s->save_g_ind = bsearch_d(s->gstar, ORC_NG, egstar);
  int ind = s->save_g_ind;
  double inte = (egstar - s->gstar[ind]) / (s->gstar[ind + 1] - s->gstar[ind]);
  double inte1 = 1.0 - inte;
  double ftrf = inte * s->trff[ind * 2 + k] + inte1 * s->trff[(ind + 1) * 2 + k];
  double val = pow(eg, 3) / ((s->gmax - s->gmin) * sqrt(egstar - egstar * egstar)) * ftrf * s->emis;
  if (s->limb == 0) return val;
  double fmu0 = inte * s->cosne[ind * 2 + k] + inte1 * s->cosne[(ind + 1) * 2 + k];
  double limb = 1.0;
  if (s->limb == 1) limb = (1.0 + 2.06 * fmu0);
  else if (s->limb == 2) limb = log(1.0 + 1.0 / fmu0);
  return val * limb;
}

static double romberg(double a, double b, int k, Relb *s) { /* :524-579 */
  const double prec = 0.02;
  double obtprec = 1.0;
  double t[7][7];
  int niter = 0;
  double r = relb_func(a, k, s);
  double rb = relb_func(b, k, s);
  double ta = (r + rb) / 2.0;
  double pas = b - a;
  t[0][0] = ta * pas;
  while ((obtprec > prec) && (niter <= 5)) {
    niter++;
    pas = pas / 2.0;
    double sum = ta;
    for (int ii = 1; ii <= pow(2, niter) - 1; ii++) sum += relb_func(a + pas * ii, k, s);
    t[0][niter] = sum * pas;
    r = 1.0;
    for (int ii = 1; ii <= niter; ii++) {
      r *= 4.0;
      int jj = niter - ii;
      t[ii][jj] = (r * t[ii - 1][jj + 1] - t[ii - 1][jj]) / (r - 1.0);
    }
    obtprec = fabs(t[niter][0] - t[niter - 1][0]) / t[niter][0];
  }
  return t[niter][0];
}

static double gstar2ener(double g, double gmin, double gmax, double ener) { return (g * (gmax - gmin) + gmin) * ener; }

static double int_edge(double blo, double bhi, double h, Relb *s, double line_energy) { /* :585-621 */
  double hex, lo, hi;
  if (blo <= 0.5) { hex = h; lo = blo; hi = bhi; }
  else { hex = 1.0 - h; lo = 1.0 - bhi; hi = 1.0 - blo; }
  double norm = 0.0;
  for (int k = 0; k < 2; k++) norm = norm + relb_func(gstar2ener(hex, s->gmin, s->gmax, line_energy), k, s);
  norm = norm * sqrt(h);
  return 2 * norm * (sqrt(hi) - sqrt(lo)) * line_energy * (s->gmax - s->gmin);
}

static double int_romb(double lo, double hi, Relb *s, double line_energy) { /* :628-647 */
  double flu = 0.0;
  if (lo >= line_energy * 0.95) {
    for (int k = 0; k < 2; k++) flu += romberg(lo, hi, k, s);
  } else {
    for (int k = 0; k < 2; k++) flu += relb_func((hi + lo) / 2.0, k, s) * (hi - lo);
  }
  return flu;
}

static double integ_relline_bin(Relb *s, double rlo0, double rhi0) { /* :650-726 */
  double line_ener = 1.0, flu = 0.0;
  double gblo = (rlo0 / line_ener - s->gmin) * s->del_g;
  if (gblo < 0.0) gblo = 0.0; else if (gblo > 1.0) gblo = 1.0;
  double gbhi = (rhi0 / line_ener - s->gmin) * s->del_g;
  if (gbhi < 0.0) gbhi = 0.0; else if (gbhi > 1.0) gbhi = 1.0;
  if (gbhi == 0) return 0.0;
  double rlo = rlo0, rhi = rhi0, hlo, hhi;
  if (gblo <= GFAC_H) {
    hlo = gblo;
    hhi = GFAC_H;
    rlo = gstar2ener(GFAC_H, s->gmin, s->gmax, line_ener);
    if (gbhi <= GFAC_H) { hhi = gbhi; rlo = -1.0; }
    flu = flu + int_edge(hlo, hhi, GFAC_H, s, line_ener);
  }
  if (gbhi >= (1.0 - GFAC_H)) {
    hhi = gbhi;
    hlo = 1.0 - GFAC_H;
    rhi = gstar2ener(1 - GFAC_H, s->gmin, s->gmax, line_ener);
    if (gblo >= (1.0 - GFAC_H)) { hlo = gblo; rhi = -1.0; }
    flu = flu + int_edge(hlo, hhi, GFAC_H, s, line_ener);
  }
  if ((rhi >= 0) && (rlo >= 0)) flu = flu + int_romb(rlo, rhi, s, line_ener);
  return flu;
}

/* calc_relline_profile (:835-958) + renorm_relline_profile (:749-796).
 * flux[nz][n_ener]; dist[nz][n_incl] (NULL = no angular distribution). returns 0 ok */
static int relline_profile(const Par *p, const SysPar *sp, const double *ener, int n_ener, const double *rgrid, int nz,
                           double *flux, double *dist, int n_incl) {
  Relb s;
  for (int i = 0; i < ORC_NG; i++) s.gstar[i] = GFAC_H + (1.0 - 2 * GFAC_H) / (ORC_NG - 1) * ((float) (i));
  double d_gstar[ORC_NG];
  for (int i = 0; i < ORC_NG; i++)
    d_gstar[i] = (i == 0 || i == ORC_NG - 1) ? 0.5 * (s.gstar[1] - s.gstar[0]) + GFAC_H : s.gstar[1] - s.gstar[0];
  for (int i = 0; i < nz * n_ener; i++) flux[i] = 0.0;
  if (dist) for (int i = 0; i < nz * n_incl; i++) dist[i] = 0.0;
  for (int ii = 0; ii < ORC_NR; ii++) {
    double egmin = sp->gmin[ii], egmax = sp->gmax[ii];
    if ((egmax > ener[0]) && (egmin < ener[n_ener])) {
      if (egmin < ener[0]) egmin = ener[0];
      if (egmax > ener[n_ener]) egmax = ener[n_ener];
      int ielo = bsearch_d(ener, n_ener + 1, egmin);
      int iehi = bsearch_d(ener, n_ener + 1, egmax);
      int izone = bsearch_d(rgrid, nz + 1, sp->re[ii]);
      s.re = sp->re[ii]; s.gmin = sp->gmin[ii]; s.gmax = sp->gmax[ii];
      s.del_g = 1. / (s.gmax - s.gmin);
      s.emis = sp->emis[ii];
      s.trff = sp->trff + (size_t) ii * ORC_NG * 2;
      s.cosne = sp->cosne + (size_t) ii * ORC_NG * 2;
      s.limb = sp->limb;
      s.save_g_ind = 0;
      double weight = trapez_single(sp->re, ii, ORC_NR) / 2;
      for (int jj = ielo; jj <= iehi; jj++) {
        double tmp = integ_relline_bin(&s, ener[jj], ener[jj + 1]);
        flux[izone * n_ener + jj] += tmp * weight;
      }
      if (dist) {
        for (int jj = 0; jj < ORC_NG; jj++) {
          double g = s.gstar[jj] * (s.gmax - s.gmin) + s.gmin;
          for (int kk = 0; kk < 2; kk++) {
            int imu = ((int) (n_incl * (1 - s.cosne[jj * 2 + kk]) + 1)) - 1; /* get_cosne_bin :799-801 */
            double tmp = s.re * pow(2 * M_PI * g * s.re, 2) / sqrt(s.gstar[jj] - s.gstar[jj] * s.gstar[jj])
                         * s.trff[jj * 2 + kk] * s.emis * weight * d_gstar[jj];
            if (tmp != tmp) return 1;
            if (imu < 0 || imu >= n_incl) return 2;
            dist[izone * n_incl + imu] += tmp;
          }
        }
      }
    }
  }
  /* renorm */
  double sum = 0.0;
  for (int i = 0; i < nz; i++)
    for (int j = 0; j < n_ener; j++) {
      flux[i * n_ener + j] /= 0.5 * (ener[j] + ener[j + 1]);
      sum += flux[i * n_ener + j];
    }
  int renorm; /* do_renorm_model, relutility.c:603-623; do_not_normalize_relline, :386-396 */
  const int phys_norm = env_is_one("RELLINE_PHYSICAL_NORM");
  if (p->model_type < 0) renorm = (p->emis_type == EMIS_LP || phys_norm) ? 0 : 1;
  else renorm = phys_norm ? 0 : 1;
  if (renorm) {
    double norm = 1;
    if (p->model_type < 0 && p->emis_type == EMIS_BKN) norm = 0.5 * cos((p->incl * 180.0 / M_PI) * M_PI / 180);
    for (int i = 0; i < nz * n_ener; i++) flux[i] *= norm / sum;
  }
  if (dist) {
    for (int i = 0; i < nz; i++) {
      double s2 = 0.0;
      for (int j = 0; j < n_incl; j++) s2 += dist[i * n_incl + j];
      if (!(s2 > 1e-8)) return 3;
      for (int j = 0; j < n_incl; j++) dist[i * n_incl + j] /= s2;
    }
  }
  return 0;
}

/* radial zone grid, src/IonGradient.cpp:457-492 (= relutility.c:265-308) */
static void zone_grid(double rmin, double rmax, int nz, double h, double *rgrid) {
  if (nz == 1) { rgrid[0] = rmin; rgrid[1] = rmax; return; }
  double r_transition = rmin;
  int indr = 0;
  if (h > rmin) {
    r_transition = h;
    log_grid(rgrid, nz + 1, rmin, rmax);
    indr = bsearch_d(rgrid, nz + 1, r_transition);
    r_transition = rgrid[indr];
  }
  if (indr < nz) {
    double rlo = r_transition, rhi = rmax;
    for (int i = indr; i < nz + 1; i++) {
      rgrid[i] = 1.0 * (i - indr) / (nz - indr) * (1.0 / rhi - 1.0 / rlo) + 1.0 / rlo;
      rgrid[i] = fabs(1.0 / rgrid[i]);
    }
  }
}

/* ------------------------------------------------------------------ nthcomp (src/donthcomp.c, diskbb seed branch) */
static double mcd_value(double et) { /* f_mcdint__ :49-108 */
  static const double gc[3] = {.078196667, -1.066202, 1.192418};
  static const double gw[3] = {.5207874, .513457, .4077983};
  static const double gn[3] = {.3728691, .039775528, .037766505};
  static const double res[98] = {
      9.6198382e-4, .0010901181, .0012310012, .0013841352, .0015481583, .0017210036, .0018988943, .002076939,
      .0022484281, .0024049483, .0025366202, .0026316255, .0026774985, .0026613059, .0025708784, .0023962965,
      .002130655, .0017725174, .0013268656, 8.0657672e-4, 2.3337584e-4, -3.6291778e-4, -9.4443569e-4,
      -.0014678875, -.0018873741, -.0021588493, -.0022448371, -.0021198179, -.0017754602, -.0012246034,
      -5.0414167e-4, 3.2507078e-4, .0011811065, .0019673402, .0025827094, .0029342526, .0029517083, .0026012166,
      .0018959062, 9.0128649e-4, -2.6757144e-4, -.0014567885, -.002492855, -.0032079776, -.0034678637,
      -.0031988217, -.0024080969, -.001193624, 2.6134145e-4, .0017117758, .0028906898, .0035614435, .0035711778,
      .0028921374, .0016385898, 4.9857464e-5, -.0015572671, -.0028578151, -.0035924212, -.0036253044,
      -.002975086, -.0018044436, -3.7796664e-4, .0010076215, .0020937327, .0027090854, .0028031667, .0024276576,
      .0017175597, 8.1030795e-4, -1.2592304e-4, -9.4888491e-4, -.0015544816, -.0018831972, -.0019203142,
      -.0016905849, -.0012487737, -6.6789911e-4, -2.7079461e-5, 5.9931935e-4, .0011499748, .0015816521,
      .0018709224, .0020129966, .0020184702, .0019089181, .0017122289, .001458377, .0011760717, 8.9046768e-4,
      6.2190822e-4, 3.8553762e-4, 1.9155022e-4, 4.5837109e-5, -4.9177834e-5, -9.3670762e-5, -8.9622968e-5,
      -4.01538532e-5};
  const double log10e = 0.43429448190325182765;
  double loget = log10e * log(et);
  double c45 = .001;
  double pos = (loget - log10e * log(c45)) / .06 + 1;
  int j = (int) pos;
  double resfact;
  if (j < 1) resfact = res[0];
  else if (j >= 98) resfact = res[97];
  else { pos -= j; resfact = res[j - 1] * (1. - pos) + res[j] * pos; }
  double gaufact = 1.;
  for (j = 1; j <= 3; ++j) {
    double z = (loget - gc[j - 1]) / gw[j - 1];
    gaufact += gn[j - 1] * exp(-z * z / 2.);
  }
  double d1 = et / .001;
  return pow(d1, -.66666666666666663) * 193.21556 * (pow(et, 1.663753) * .52876731 + 1.) * exp(-et) * gaufact
         * (resfact + 1.);
}

/* Kompaneets solution with disk-blackbody seed: f_thdscompton__ :467-648 + f_thermlc__ :200-301.
 * x[0..jmax] (1-based in the original -> stored 0-based here as x[j-1]); sptot[j-1] */
static void thdscompton(double tempbb, double theta, double gamma, double *x, int *jmax_out, double *sptot) {
  const double log10e = 0.43429448190325182765;
  static double c2[900], bet[900], rel[900], dphesc[900], dphdot[900];
  static double ear[5001], photar[5000];
  double d1 = gamma + .5;
  double tautom = sqrt(3. / (theta * (d1 * d1 - 2.25)) + 2.25) - 1.5;
  for (int j = 0; j < 900; j++) { dphesc[j] = dphdot[j] = rel[j] = bet[j] = c2[j] = 0.; sptot[j] = 0.; }
  double delta = .02;
  double deltal = delta * log(10.);
  double xmin = tempbb * 1e-4, xmax = theta * 40.;
  d1 = xmax / xmin;
  int jmax = (int) (log10e * log(d1) / delta) + 1;
  if (jmax > 899) jmax = 899;
  for (int j = 1; j <= jmax + 1; j++) x[j - 1] = xmin * pow(10., (j - 1) * delta);
  for (int j = 1; j <= jmax; j++) {
    double w = x[j - 1];
    double w1 = sqrt(x[j - 1] * x[j]);
    c2[j - 1] = pow(w1, 4.) / (w1 * 4.6 + 1. + w1 * 1.1 * w1);
    if (w <= .05) {
      rel[j - 1] = 1 - w * 2 + w * 26 * w / 5;
    } else {
      double z1 = (w + 1) / (w * (w * w));
      double z2 = w * 2 + 1;
      double z3 = log(z2);
      double z4 = w * 2 * (w + 1) / z2;
      double z5 = z3 / 2 / w;
      double z6 = (w * 3 + 1) / z2 / z2;
      rel[j - 1] = (z1 * (z4 - z3) + z5 - z6) * .75;
    }
  }
  d1 = tempbb * 50. / xmin;
  int jmaxth = (int) (log10e * log(d1) / delta);
  if (jmaxth > 900) jmaxth = 900;
  if (jmaxth > jmax) jmaxth = jmax;
  for (int j = 1; j <= jmaxth - 1; j++) ear[j - 1] = sqrt(x[j - 1] * x[j]) * 511.;
  double tin = tempbb * 511.;
  int ne = jmaxth - 2;
  { /* f_xsdskb__ :142-197: 5-point Gauss integration of the multicolour disk spectrum */
    static const double gw5[5] = {.236926885, .47862867, .568888888, .47862867, .236926885};
    static const double gx5[5] = {-.906179846, -.53846931, 0., .53846931, .906179846};
    for (int i = 1; i <= ne; i++) {
      double xn = (ear[i] - ear[i - 1]) / 2.f;
      double ph = 0.f;
      double xh = xn + ear[i - 1];
      for (int j = 0; j < 5; j++) {
        double e = xn * gx5[j] + xh;
        double flux;
        if (tin == 0.) flux = 0.;
        else { double et = e / tin; flux = mcd_value(et) * tin * tin * 1. / 361.; }
        ph += gw5[j] * flux;
      }
      photar[i - 1] = ph * xn;
    }
  }
  for (int j = 1; j <= ne; j++) dphdot[j] = photar[j - 1] * 511. / (ear[j] - ear[j - 1]);
  jmaxth = ne + 1;
  dphdot[0] = dphdot[1];
  d1 = .1 / xmin;
  int jnr = (int) (log10e * log(d1) / delta + 1);
  if (jnr > jmax - 1) jnr = jmax - 1;
  d1 = 1. / xmin;
  int jrel = (int) (log10e * log(d1) / delta + 1);
  if (jrel > jmax) jrel = jmax;
  double xnr = x[jnr - 1], xr = x[jrel - 1];
  for (int j = 1; j <= jnr - 1; j++) {
    double taukn = tautom * rel[j - 1];
    bet[j - 1] = 1 / tautom / (taukn / 3 + 1);
  }
  for (int j = jnr; j <= jrel; j++) {
    double taukn = tautom * rel[j - 1];
    double arg = (x[j - 1] - xnr) / (xr - xnr);
    double flz = 1 - arg;
    bet[j - 1] = 1 / tautom / (taukn / 3 * flz + 1);
  }
  for (int j = jrel + 1; j <= jmax; j++) bet[j - 1] = 1 / tautom;
  { /* f_thermlc__: tridiagonal solve; arrays 1-based in the original (a[j-1] etc.) */
    static double a[900], b[900], c[900], d[900], g[900], u[900], gam[900], alp[900];
    double c20 = tautom / deltal;
    for (int j = 2; j <= jmax - 1; j++) {
      double w1 = sqrt(x[j - 1] * x[j]);
      double w2 = sqrt(x[j - 2] * x[j - 1]);
      a[j - 1] = -c20 * c2[j - 1] * (theta / deltal / w1 + .5);
      double t1 = -c20 * c2[j - 1] * (.5 - theta / deltal / w1);
      double t2 = c20 * c2[j - 2] * (theta / deltal / w2 + .5);
      double t3 = pow(x[j - 1], 3.) * (tautom * bet[j - 1]);
      b[j - 1] = t1 + t2 + t3;
      c[j - 1] = c20 * c2[j - 2] * (.5 - theta / deltal / w2);
      d[j - 1] = x[j - 1] * dphdot[j - 1];
    }
    double x32 = sqrt(x[0] * x[1]);
    double aa = (theta / deltal / x32 + .5) / (theta / deltal / x32 - .5);
    u[jmax - 1] = 0.;
    alp[1] = b[1] + c[1] * aa;
    gam[1] = a[1] / alp[1];
    for (int j = 3; j <= jmax - 1; j++) {
      alp[j - 1] = b[j - 1] - c[j - 1] * gam[j - 2];
      gam[j - 1] = a[j - 1] / alp[j - 1];
    }
    g[1] = d[1] / alp[1];
    for (int j = 3; j <= jmax - 2; j++) g[j - 1] = (d[j - 1] - c[j - 1] * g[j - 2]) / alp[j - 1];
    g[jmax - 2] = (d[jmax - 2] - a[jmax - 2] * u[jmax - 1] - c[jmax - 2] * g[jmax - 3]) / alp[jmax - 2];
    u[jmax - 2] = g[jmax - 2];
    for (int j = 3; j <= jmax - 1; j++) {
      int jj = jmax + 1 - j;
      u[jj - 1] = g[jj - 1] - gam[jj - 1] * u[jj];
    }
    u[0] = aa * u[1];
    for (int j = 1; j <= jmax; j++) dphesc[j - 1] = x[j - 1] * x[j - 1] * u[j - 1] * bet[j - 1] * tautom;
  }
  for (int j = 1; j <= jmax - 1; j++) sptot[j - 1] = dphesc[j - 1] * (x[j - 1] * x[j - 1]);
  *jmax_out = jmax;
}

/* c_donthcomp :651-793 with param = (gamma, kTe, kTbb=0.05, inp_type=1, z) (relutility.c:625-632) */
void orc_nthcomp(const double *ear, int ne, double gamma, double kte, double z_red, double *photar) {
  static double xth[901], spt[901];
  int nth;
  thdscompton(0.05 / 511., kte / 511., gamma, xth, &nth, spt);
  double xn = (z_red + 1) / 511.;
  double normfac;
  { /* f_spp__ :651-681 evaluated at y = 1/xn (stateless form of its static cursor) */
    double xx = 1 / (1 / xn);
    int ih = 2;
    while (ih < nth && xx > xth[ih - 1]) ++ih;
    int il = ih - 1;
    double v = spt[il - 1] + (spt[ih - 1] - spt[il - 1]) * (xx - xth[il - 1]) / (xth[ih - 1] - xth[il - 1]);
    normfac = 1 / v;
  }
  double *prim = (double *) malloc(sizeof(double) * (ne + 1));
  for (int i = 0; i <= ne; i++) prim[i] = 0.;
  int j = 1;
  for (int i = 0; i <= ne; i++) {
    while (j <= nth && xth[j - 1] * 511. < ear[i] * (z_red + 1)) ++j;
    if (j <= nth) {
      if (j > 1) {
        int jl = j - 1;
        prim[i] = spt[jl - 1] + (ear[i] / 511. * (z_red + 1) - xth[jl - 1]) * (spt[jl] - spt[jl - 1]) / (xth[jl] - xth[jl - 1]);
      } else {
        prim[i] = spt[0];
      }
    }
  }
  for (int i = 1; i <= ne; i++)
    photar[i - 1] = (prim[i] / pow(ear[i], 2.) + prim[i - 1] / pow(ear[i - 1], 2.)) * .5 * (ear[i] - ear[i - 1]) * normfac;
  free(prim);
}

/* ------------------------------------------------------------------ primary spectrum + xillver normalisation */
typedef struct { double gam, afe, lxi, ect, dens; int prim_type; int xtab; double ktbb, frac; } XPar;

/* src/Xillspec.cpp:215-300 */
static void primary_spectrum(double *out, const double *ener, int n, const XPar *x, double shift) {
  if (x->prim_type == PRIM_ECUT) {
    double ecut = x->ect * shift;
    for (int i = 0; i < n; i++) {
      double en = 0.5 * (ener[i] + ener[i + 1]);
      out[i] = exp(1.0 / ecut) * pow(en, -x->gam) * exp(-en / ecut) * (ener[i + 1] - ener[i]);
    }
  } else if (x->prim_type == PRIM_BB) { /* spec_blackbody, src/Xillspec.cpp:262-269 (not shifted) */
    for (int i = 0; i < n; i++) {
      double en = 0.5 * (ener[i] + ener[i + 1]);
      out[i] = en * en / (pow(x->ktbb, 4) * (exp(en / x->ktbb) - 1));
      out[i] *= (ener[i + 1] - ener[i]);
    }
  } else {
    orc_nthcomp(ener, n, x->gam, x->ect, 1 / shift - 1, out);
  }
}
/* src/Xillspec.cpp:179-205 */
static double norm_wrt_xillver(const double *flux, const double *ener, int n) {
  double keV2erg = 1.602177e-09;
  double sum_pl = 0.0;
  for (int i = 0; i < n; i++)
    if (ener[i] >= 0.1 && ener[i] <= 1000.0) sum_pl += flux[i] * 0.5 * (ener[i] + ener[i + 1]) * 1e20 * keV2erg;
  return sum_pl / (1e15 / 4.0 / M_PI);
}
static double norm_factor_source(const XPar *x) { /* PrimarySource.h:279-293 */
  double ps[N_COARSE];
  primary_spectrum(ps, g_ecoarse, N_COARSE, x, 1.0);
  return 1. / norm_wrt_xillver(ps, g_ecoarse, N_COARSE);
}

/* ------------------------------------------------------------------ xillver table interpolation */
/* src/xilltable.c:297-323, :812-876, :999-1019, :1054-1181 — all inclinations, no interpolation over Incl.
 * flu[n_incl][n_ener] */
static int xillver_spectra(const XPar *x, double *flu) {
  if (!g_xill[x->xtab] && load_xill(x->xtab)) return 1;
  const XillTab *t = g_xill[x->xtab];
  float inp[8];
  inp[XP_GAM] = (float) x->gam; inp[XP_AFE] = (float) x->afe; inp[XP_LXI] = (float) x->lxi;
  inp[XP_ECT] = (float) x->ect; inp[XP_DNS] = (float) x->dens; inp[XP_INC] = 0.f;
  inp[XP_KTB] = (float) x->ktbb; inp[XP_FRA] = (float) x->frac;
  int ind[6];
  double fac[6];
  int ax_ect = -1;
  for (int i = 0; i < t->npar; i++) {
    ind[i] = bsearch_f(t->vals[i], t->nvals[i], inp[t->pindex[i]]);
    if (ind[i] < 0) ind[i] = 0; else if (ind[i] > t->nvals[i] - 2) ind[i] = t->nvals[i] - 2;
    if (t->pindex[i] == XP_ECT) ax_ect = i;
  }
  for (int i = 0; i < t->npar; i++) {
    int pind = t->pindex[i];
    if (pind != XP_INC) {
      float lo = t->vals[i][0], hi = t->vals[i][t->nvals[i] - 1];
      if (inp[pind] < lo) inp[pind] = lo; else if (inp[pind] > hi) inp[pind] = hi;
    }
    fac[i] = (inp[pind] - t->vals[i][ind[i]]) / (t->vals[i][ind[i] + 1] - t->vals[i][ind[i]]);
  }
  if (ax_ect >= 0) { /* ensure_ecut_within_boundarys (indexes param_vals by the global id 3: same axis here) */
    if (x->ect <= t->vals[3][0]) fac[ax_ect] = 0.0;
    if (x->ect >= t->vals[3][t->nvals[3] - 1]) fac[ax_ect] = 1.0;
  }
  int n_ener = t->n_ener, n_incl = t->n_incl;
  int off = (t->npar == 6) ? 1 : 0;
  const int *n = t->nvals;
  double f1 = fac[off], f2 = fac[off + 1], f3 = fac[off + 2], f4 = fac[off + 3];
  double w[16];
  /* weight/term order of interp_5d_tab_incl: 0000 1000 0100 0010 1100 1010 0110 1110 0001 1001 0101 0011 1101 1011 0111 1111 */
  static const int bits[16][4] = {{0,0,0,0},{1,0,0,0},{0,1,0,0},{0,0,1,0},{1,1,0,0},{1,0,1,0},{0,1,1,0},{1,1,1,0},
                                  {0,0,0,1},{1,0,0,1},{0,1,0,1},{0,0,1,1},{1,1,0,1},{1,0,1,1},{0,1,1,1},{1,1,1,1}};
  for (int c = 0; c < 16; c++)
    w[c] = (bits[c][0] ? f1 : (1.0 - f1)) * (bits[c][1] ? f2 : (1.0 - f2)) * (bits[c][2] ? f3 : (1.0 - f3))
           * (bits[c][3] ? f4 : (1 - f4));
  double *s2 = (t->npar == 6) ? (double *) malloc(sizeof(double) * n_ener) : NULL;
  for (int mm = 0; mm < n_incl; mm++) {
    for (int half = 0; half < (t->npar == 6 ? 2 : 1); half++) {
      const float *dat[16];
      for (int c = 0; c < 16; c++) {
        long row;
        int i1 = ind[off] + bits[c][0], i2 = ind[off + 1] + bits[c][1], i3 = ind[off + 2] + bits[c][2],
            i4 = ind[off + 3] + bits[c][3];
        if (t->npar == 5) row = ((((long) i1 * n[1] + i2) * n[2] + i3) * n[3] + i4) * n[4] + mm;
        else row = (((((long) (ind[0] + half) * n[1] + i1) * n[2] + i2) * n[3] + i3) * n[4] + i4) * n[5] + mm;
        dat[c] = t->data + (size_t) row * n_ener;
      }
      double *dst = (half == 0) ? flu + (size_t) mm * n_ener : s2;
      for (int e = 0; e < n_ener; e++) {
        double v = w[0] * (double) dat[0][e];
        for (int c = 1; c < 16; c++) v += w[c] * (double) dat[c][e];
        dst[e] = v;
      }
    }
    if (t->npar == 6) {
      double *s1 = flu + (size_t) mm * n_ener;
      for (int e = 0; e < n_ener; e++) s1[e] = lin1d(fac[0], s1[e], s2[e]);
    }
  }
  free(s2);
  return 0;
}

/* Standalone xillver models: interpolation over ALL axes including the inclination, interp_5d_tab
 * (src/xilltable.c:878-996) and interp_6d_tab (:1022-1044), bracket/factor code of interp_xill_table
 * (:1090-1181).  flu[n_ener] */
static int xillver_spectrum_incl(const XPar *x, double incl_deg, double *flu) {
  if (!g_xill[x->xtab] && load_xill(x->xtab)) return 1;
  const XillTab *t = g_xill[x->xtab];
  float inp[8];
  inp[XP_GAM] = (float) x->gam; inp[XP_AFE] = (float) x->afe; inp[XP_LXI] = (float) x->lxi;
  inp[XP_ECT] = (float) x->ect; inp[XP_DNS] = (float) x->dens; inp[XP_INC] = (float) incl_deg;
  inp[XP_KTB] = (float) x->ktbb; inp[XP_FRA] = (float) x->frac;
  int ind[6];
  double fac[6];
  int ax_ect = -1;
  for (int i = 0; i < t->npar; i++) {
    ind[i] = bsearch_f(t->vals[i], t->nvals[i], inp[t->pindex[i]]);
    if (ind[i] < 0) ind[i] = 0; else if (ind[i] > t->nvals[i] - 2) ind[i] = t->nvals[i] - 2;
    if (t->pindex[i] == XP_ECT) ax_ect = i;
  }
  for (int i = 0; i < t->npar; i++) {
    int pind = t->pindex[i];
    float lo = t->vals[i][0], hi = t->vals[i][t->nvals[i] - 1];
    if (inp[pind] < lo) inp[pind] = lo; else if (inp[pind] > hi) inp[pind] = hi;
    fac[i] = (inp[pind] - t->vals[i][ind[i]]) / (t->vals[i][ind[i] + 1] - t->vals[i][ind[i]]);
  }
  if (ax_ect >= 0) {
    if (x->ect <= t->vals[3][0]) fac[ax_ect] = 0.0;
    if (x->ect >= t->vals[3][t->nvals[3] - 1]) fac[ax_ect] = 1.0;
  }
  int n_ener = t->n_ener;
  int off = (t->npar == 6) ? 1 : 0;
  const int *n = t->nvals;
  double f[5] = {fac[off], fac[off + 1], fac[off + 2], fac[off + 3], fac[off + 4]};
  /* term order of interp_5d_tab: the 16 patterns of interp_5d_tab_incl with the 5th bit 0, then with it 1 */
  static const int b16[16][4] = {{0,0,0,0},{1,0,0,0},{0,1,0,0},{0,0,1,0},{1,1,0,0},{1,0,1,0},{0,1,1,0},{1,1,1,0},
                                 {0,0,0,1},{1,0,0,1},{0,1,0,1},{0,0,1,1},{1,1,0,1},{1,0,1,1},{0,1,1,1},{1,1,1,1}};
  double w[32];
  for (int h5 = 0; h5 < 2; h5++)
    for (int c = 0; c < 16; c++)
      w[h5 * 16 + c] = (b16[c][0] ? f[0] : (1.0 - f[0])) * (b16[c][1] ? f[1] : (1.0 - f[1])) * (b16[c][2] ? f[2] : (1.0 - f[2]))
                       * (b16[c][3] ? f[3] : (1 - f[3])) * (h5 ? f[4] : (1 - f[4]));
  double *s2 = (t->npar == 6) ? (double *) malloc(sizeof(double) * n_ener) : NULL;
  for (int half = 0; half < (t->npar == 6 ? 2 : 1); half++) {
    const float *dat[32];
    for (int h5 = 0; h5 < 2; h5++)
      for (int c = 0; c < 16; c++) {
        long row;
        int i1 = ind[off] + b16[c][0], i2 = ind[off + 1] + b16[c][1], i3 = ind[off + 2] + b16[c][2],
            i4 = ind[off + 3] + b16[c][3], i5 = ind[off + 4] + h5;
        if (t->npar == 5) row = ((((long) i1 * n[1] + i2) * n[2] + i3) * n[3] + i4) * n[4] + i5;
        else row = (((((long) (ind[0] + half) * n[1] + i1) * n[2] + i2) * n[3] + i3) * n[4] + i4) * n[5] + i5;
        dat[h5 * 16 + c] = t->data + (size_t) row * n_ener;
      }
    double *dst = (half == 0) ? flu : s2;
    for (int e = 0; e < n_ener; e++) {
      double v = w[0] * (double) dat[0][e];
      for (int c = 1; c < 32; c++) v += w[c] * (double) dat[c][e];
      dst[e] = v;
    }
  }
  if (t->npar == 6) {
    for (int e = 0; e < n_ener; e++) flu[e] = lin1d(fac[0], flu[e], s2[e]);
    free(s2);
  }
  return 0;
}

static void xill_energy_grid(int xtab, double *ener) { /* xilltable.c:1105-1109 */
  const XillTab *t = g_xill[xtab];
  for (int i = 0; i < t->n_ener; i++) ener[i] = t->elo[i];
  ener[t->n_ener] = t->ehi[t->n_ener - 1];
}

/* returning-radiation correction factors of one zone, src/Xillspec.cpp:109-126, :344-362, :457-526 */
static void fluxcorr_factors(const double *flu, int n_incl, int n_ener, const double *ener, const float *incl,
                             const XPar *xz, double *fac_flux, double *fac_gshift) {
  double *avg = (double *) calloc(n_ener, sizeof(double));
  for (int i = 0; i < n_incl; i++) {
    double incl_deg = ((double) incl[i]) * 180 / M_PI;
    double f = 0.5 * cos(incl_deg * M_PI / 180) / n_incl;
    for (int j = 0; j < n_ener; j++) avg[j] += f * flu[(size_t) i * n_ener + j];
  }
  double direct[N_COARSE];
  primary_spectrum(direct, g_ecoarse, N_COARSE, xz, 1.0);
  double nf = norm_factor_source(xz);
  for (int i = 0; i < N_COARSE; i++) direct[i] *= nf;
  double s1 = 0.0, s2 = 0.0;
  for (int i = 0; i < n_ener; i++)
    if (ener[i] >= 0.1 && ener[i + 1] <= 1000) s1 += avg[i] * 0.5 * (ener[i] + ener[i + 1]);
  for (int i = 0; i < N_COARSE; i++)
    if (g_ecoarse[i] >= 0.1 && g_ecoarse[i + 1] <= 1000) s2 += direct[i] * 0.5 * (g_ecoarse[i] + g_ecoarse[i + 1]);
  *fac_flux = s1 / s2;
  const double gref = 2. / 3.;
  double *ez = (double *) malloc(sizeof(double) * (n_ener + 1));
  double *fz = (double *) malloc(sizeof(double) * n_ener);
  for (int i = 0; i <= n_ener; i++) ez[i] = ener[i] / gref;
  orc_rebin(ez, fz, n_ener, ener, avg, n_ener);
  for (int i = 0; i < n_ener; i++) fz[i] *= gref;
  double p1 = 0.0, p2 = 0.0;
  for (int i = 0; i < n_ener; i++)
    if (ener[i] >= 0.15 && ener[i + 1] <= 500.0) { p1 += avg[i]; p2 += fz[i]; }
  *fac_gshift = (p1 / p2) / pow(1.5, xz->gam);
  free(avg); free(ez); free(fz);
}

/* ------------------------------------------------------------------ FFT convolution (src/Relbase.cpp:119-230) */
static void fft_c(int n, int sign, double *xr, double *xi) {
  for (int i = 1, j = 0; i < n; i++) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { double t = xr[i]; xr[i] = xr[j]; xr[j] = t; t = xi[i]; xi[i] = xi[j]; xi[j] = t; }
  }
  for (int len = 2; len <= n; len <<= 1) {
    int half = len >> 1;
    for (int k = 0; k < half; k++) {
      double ang = sign * 2.0 * M_PI * k / len;
      double c = cos(ang), s = sin(ang);
      for (int i = k; i < n; i += len) {
        double vr = xr[i + half] * c - xi[i + half] * s, vi = xr[i + half] * s + xi[i + half] * c;
        xr[i + half] = xr[i] - vr; xi[i + half] = xi[i] - vi;
        xr[i] += vr; xi[i] += vi;
      }
    }
  }
}

void orc_fft_conv(const double *fxill, const double *frel, double *fout) {
  const int n = ORC_NCONV;
  const double *ener = g_econv;
  static double xr[ORC_NCONV], xi[ORC_NCONV], yr[ORC_NCONV], yi[ORC_NCONV], cf[ORC_NCONV];
  for (int i = 0; i < n; i++) cf[i] = 0.5 * (ener[i] + ener[i + 1]) / (ener[i + 1] - ener[i]);
  int i1 = bsearch_d(ener, n + 1, 1.0);
  for (int i = 0; i < n; i++) { xr[i] = fxill[i] * cf[i]; xi[i] = 0.0; yi[i] = 0.0; }
  for (int i = 0; i < n; i++) yr[(i - i1 + n) % n] = frel[i] * cf[i];
  fft_c(n, -1, xr, xi);
  fft_c(n, -1, yr, yi);
  for (int i = 0; i < n; i++) {
    double re = xr[i] * yr[i] - xi[i] * yi[i], im = xr[i] * yi[i] + xi[i] * yr[i];
    xr[i] = re; xi[i] = im;
  }
  fft_c(n, +1, xr, xi);
  for (int i = 0; i < n; i++) fout[i] = xr[i] / cf[i];
  double s_rel = 0.0, s_xill = 0.0, s_conv = 0.0; /* calcFFTNormFactor :199-213 */
  for (int j = 0; j < n; j++)
    if (ener[j] >= 0.01 && ener[j + 1] < 1000.0) { s_xill += fxill[j]; s_rel += frel[j]; s_conv += fout[j]; }
  double norm = s_rel * s_xill / s_conv;
  for (int i = 0; i < n; i++) fout[i] *= norm;
}

/* ------------------------------------------------------------------ relxill pipeline (src/Relxill.cpp:289-413) */
typedef struct {
  int nz, n_ener_x, n_incl;
  double zone[ORC_NZMAX + 1], lxi[ORC_NZMAX], dens[ORC_NZMAX], ect[ORC_NZMAX], eshift[ORC_NZMAX];
  double corr_flux[ORC_NZMAX], corr_gshift[ORC_NZMAX], normch[ORC_NZMAX];
  double emis2[ORC_NR];
  double *relflux, *dist, *xill; /* [nz][4096], [nz][n_incl], [nz][n_ener_x] */
  double conv[ORC_NCONV], total[ORC_NCONV];
} Stages;

static void free_stages(Stages *s) { free(s->relflux); free(s->dist); free(s->xill); }

static int relxill_pipeline(Par *p, Stages *st) {
  int rc;
  if (p->emis_type == EMIS_LP && p->prim_type == PRIM_ECUT) p->ect /= energy_shift_source_obs(p);
  XPar src = {p->gam, p->afe, p->lxi, p->ect, p->dens, p->prim_type, p->xtab, p->ktbb, p->frac_pl_bb};
  double shift_obs = (p->emis_type == EMIS_LP) ? energy_shift_source_obs(p) : 1.0;
  SysPar *sp = new_syspar();
  p->corr_flux = NULL;
  if ((rc = system_parameters(p, sp))) { free_syspar(sp); return 100 + rc; }
  double refl_frac_pred = sp->refl_frac, f_inf_rest = sp->f_inf_rest;

  /* zones + ionisation gradient: src/IonGradient.cpp:340-390 */
  int nz = p->num_zones;
  st->nz = nz;
  zone_grid(p->rin, p->rout, nz, p->height, st->zone);
  double rmean[ORC_NZMAX], del_emit[ORC_NZMAX];
  for (int i = 0; i < nz; i++) rmean[i] = 0.5 * (st->zone[i] + st->zone[i + 1]);
  if (inv_rebin_mean(sp->re, sp->del_emit, ORC_NR, rmean, del_emit, nz)) { free_syspar(sp); return 201; }
  for (int i = 0; i < nz; i++)
    st->eshift[i] = (p->emis_type == EMIS_LP) ? gi_potential_lp(rmean[i], p->a, p->height, p->beta, del_emit[i]) : 1.0;
  if (p->ion_grad_type == ION_PL) { /* :174-182 */
    for (int i = 0; i < nz; i++) {
      st->lxi[i] = (exp(src.lxi)) * pow((rmean[i] / rmean[0]), -1.0 * p->iongrad_index);
      st->lxi[i] = log(st->lxi[i]);
      st->dens[i] = src.dens;
    }
  } else if (p->ion_grad_type == ION_ALPHA) { /* :128-171 */
    double rin = st->zone[0];
    double irr[ORC_NZMAX], dinc[ORC_NZMAX];
    if (inv_rebin_mean(sp->re, sp->emis, ORC_NR, rmean, irr, nz)) { free_syspar(sp); return 202; }
    if (inv_rebin_mean(sp->re, sp->del_inc, ORC_NR, rmean, dinc, nz)) { free_syspar(sp); return 203; }
    double rad_lxi = pow((11. / 9.), 2) * rin;
    int kk = inv_bsearch_d(sp->re, ORC_NR, rad_lxi);
    double interp = (rad_lxi - sp->re[kk + 1]) / (sp->re[kk] - sp->re[kk + 1]);
    double e_at = lin1d(interp, sp->emis[kk + 1], sp->emis[kk]);
    double di_at = lin1d(interp, sp->del_inc[kk + 1], sp->del_inc[kk]);
    double lxi_max = log10(4.0 * M_PI * e_at / density_ss73_zone_a(rad_lxi, rin) * (cos(M_PI / 4) / cos(di_at)));
    double fac_lxi_norm = src.lxi - lxi_max;
    double density_min = density_ss73_zone_a((25. / 9.) * rin, rin);
    const char *cd_env = getenv("RELXILL_CONSTANT_DENSITY"); /* constantDiskDensity(), src/relutility.c:372-382 */
    int const_dens = (cd_env != NULL && (int) strtod(cd_env, NULL) == 1);
    for (int i = 0; i < nz; i++) {
      double dn = const_dens ? 1.0 : density_ss73_zone_a(rmean[i], rin) / density_min; /* :155-158 */
      st->dens[i] = log10(dn) + src.dens;
      st->lxi[i] = log10(4.0 * M_PI * irr[i] / dn * (cos(M_PI / 4) / cos(dinc[i])));
      st->lxi[i] += fac_lxi_norm;
    }
  } else {
    for (int i = 0; i < nz; i++) { st->lxi[i] = src.lxi; st->dens[i] = src.dens; }
  }
  for (int i = 0; i < nz; i++) {
    if (st->lxi[i] < 0.0) st->lxi[i] = 0.0; else if (st->lxi[i] > 4.7) st->lxi[i] = 4.7;
    st->ect[i] = src.ect * st->eshift[i];
  }

  /* xillver spectra per zone */
  if (!g_xill[p->xtab] && load_xill(p->xtab)) { free_syspar(sp); return 301; }
  const XillTab *xt = g_xill[p->xtab];
  int nex = xt->n_ener, ni = xt->n_incl;
  st->n_ener_x = nex; st->n_incl = ni;
  double *ex = (double *) malloc(sizeof(double) * (nex + 1));
  xill_energy_grid(p->xtab, ex);
  double *flu = (double *) malloc(sizeof(double) * (size_t) nz * ni * nex);
  for (int i = 0; i < nz; i++) {
    XPar xz = {src.gam, src.afe, st->lxi[i], st->ect[i], st->dens[i], p->prim_type, p->xtab, p->ktbb, p->frac_pl_bb};
    xillver_spectra(&xz, flu + (size_t) i * ni * nex);
  }
  /* returning-radiation correction factors + second system-parameter pass (Relxill.cpp:337-344) */
  for (int i = 0; i < nz; i++) st->corr_flux[i] = st->corr_gshift[i] = 0.0;
  if (p->return_rad != 0 && p->a > 0.0) {
    for (int i = 0; i < nz; i++) {
      XPar xz = {src.gam, src.afe, st->lxi[i], st->ect[i], st->dens[i], p->prim_type, p->xtab, p->ktbb, p->frac_pl_bb};
      fluxcorr_factors(flu + (size_t) i * ni * nex, ni, nex, ex, xt->incl, &xz, &st->corr_flux[i], &st->corr_gshift[i]);
    }
    p->corr_rgrid = st->zone; p->corr_flux = st->corr_flux; p->corr_gshift = st->corr_gshift; p->corr_nz = nz;
    if ((rc = system_parameters(p, sp))) { free(ex); free(flu); free_syspar(sp); return 400 + rc; }
    p->corr_flux = NULL;
  }
  memcpy(st->emis2, sp->emis, sizeof(st->emis2));

  /* relline profile on the convolution grid with angular distribution */
  st->relflux = (double *) malloc(sizeof(double) * (size_t) nz * ORC_NCONV);
  st->dist = (double *) malloc(sizeof(double) * (size_t) nz * ni);
  st->xill = (double *) malloc(sizeof(double) * (size_t) nz * nex);
  if ((rc = relline_profile(p, sp, g_econv, ORC_NCONV, st->zone, nz, st->relflux, st->dist, ni))) {
    free(ex); free(flu); free_syspar(sp); return 500 + rc;
  }
  /* angle weighting (Xillspec.cpp:557-573) and normalisation change (Xillspec.cpp:408-440) */
  double nsrc = norm_factor_source(&src);
  for (int i = 0; i < nz; i++) {
    double *xz = st->xill + (size_t) i * nex;
    for (int e = 0; e < nex; e++) xz[e] = 0.0;
    for (int m = 0; m < ni; m++)
      for (int e = 0; e < nex; e++) xz[e] += st->dist[i * ni + m] * flu[((size_t) i * ni + m) * nex + e];
    XPar xd = src;
    xd.ect = src.ect * st->eshift[i];
    st->normch[i] = norm_factor_source(&xd) / nsrc;
    for (int e = 0; e < nex; e++) xz[e] /= st->normch[i];
  }
  /* convolution over zones (Relxill.cpp:432-482) */
  static double rebinned[ORC_NCONV], cout[ORC_NCONV];
  for (int e = 0; e < ORC_NCONV; e++) st->conv[e] = 0.0;
  for (int i = 0; i < nz; i++) {
    double s = 0.0;
    for (int e = 0; e < ORC_NCONV; e++) s += st->relflux[(size_t) i * ORC_NCONV + e];
    if (s < 1e-12) continue;
    orc_rebin(g_econv, rebinned, ORC_NCONV, ex, st->xill + (size_t) i * nex, nex);
    orc_fft_conv(rebinned, st->relflux + (size_t) i * ORC_NCONV, cout);
    for (int e = 0; e < ORC_NCONV; e++) st->conv[e] += cout[e];
  }
  /* primary spectrum (PrimarySource.cpp:66-125) */
  static double prim[ORC_NCONV];
  primary_spectrum(prim, g_econv, ORC_NCONV, &src, shift_obs);
  for (int e = 0; e < ORC_NCONV; e++) { prim[e] *= nsrc; st->total[e] = st->conv[e]; }
  if (p->emis_type != EMIS_LP) {
    for (int e = 0; e < ORC_NCONV; e++) st->total[e] *= fabs(p->refl_frac);
  } else {
    double rf = p->refl_frac;
    if (p->boost) rf *= refl_frac_pred;
    double pfac = f_inf_rest / 0.5 * pow(shift_obs, src.gam);
    if (p->beta > 1e-4) pfac *= pow(doppler_factor(M_PI - p->incl, p->beta), 2);
    double nfr = (fabs(rf)) / refl_frac_pred;
    for (int e = 0; e < ORC_NCONV; e++) { st->total[e] *= nfr; prim[e] *= pfac; }
  }
  if (p->refl_frac >= 0)
    for (int e = 0; e < ORC_NCONV; e++) st->total[e] += prim[e];
  free(ex); free(flu); free_syspar(sp);
  return 0;
}

/* ------------------------------------------------------------------ model entry points */
static void shifted_grid(const double *energy, int n_flux, double z, double *out) { /* XspecSpectrum.h:70-76 */
  for (int i = 0; i <= n_flux; i++) out[i] = energy[i];
  if (z > 0) for (int i = 0; i <= n_flux; i++) out[i] *= (1 + z);
}

int orc_eval_model(const char *model, const double *energy, int n_flux, const double *par, double *flux) {
  const ModelDef *m = find_model(model);
  if (!m) return -1;
  Par p;
  if (interpret_params(m, par, &p)) return 1;
  double *e = (double *) malloc(sizeof(double) * (n_flux + 1));
  shifted_grid(energy, n_flux, p.z, e);
  int rc = 0;
  if (m->type == T_LINE) { /* LocalModel.cpp:32-51 */
    for (int i = 0; i <= n_flux; i++) e[i] /= p.lineE;
    SysPar *sp = new_syspar();
    rc = system_parameters(&p, sp);
    if (!rc) {
      double rgrid[2] = {p.rin, p.rout};
      rc = relline_profile(&p, sp, e, n_flux, rgrid, 1, flux, NULL, 0);
    }
    free_syspar(sp);
  } else if (m->type == T_RELXILL) {
    Stages st;
    memset(&st, 0, sizeof(st));
    rc = relxill_pipeline(&p, &st);
    if (!rc && env_is_one("RELXILL_RENORMALIZE")) { /* renorm_relxill_spectrum_1keV, Relxill.cpp:241-259: 1 cts/s/keV/cm2 at 3 keV */
      int klo = 0, khi = ORC_NCONV - 1; /* binary_search(energy, num_flux_bins, 3.0), relutility.c:155-171 */
      while (khi - klo > 1) {
        int k = (khi + klo) / 2;
        if (g_econv[k] > 3.0) khi = k; else klo = k;
      }
      double dE = g_econv[klo + 1] - g_econv[klo];
      double nf = 1.0 / (st.total[klo] / dE);
      for (int i = 0; i < ORC_NCONV; i++) st.total[i] *= nf;
    }
    if (!rc) orc_rebin(e, flux, n_flux, g_econv, st.total, ORC_NCONV); /* Relxill.cpp:261-278 */
    free_stages(&st);
  } else if (m->type == T_CONV) { /* LocalModel.cpp:82-98, Relbase.cpp:257-289 */
    double s = 0.0;
    for (int i = 0; i < n_flux; i++) s += flux[i];
    if (s <= 0.0) { free(e); return 2; }
    SysPar *sp = new_syspar();
    rc = system_parameters(&p, sp);
    static double prof[ORC_NCONV], rb[ORC_NCONV], co[ORC_NCONV];
    if (!rc) {
      double rgrid[2] = {p.rin, p.rout};
      rc = relline_profile(&p, sp, g_econv, ORC_NCONV, rgrid, 1, prof, NULL, 0);
    }
    if (!rc) {
      orc_rebin(g_econv, rb, ORC_NCONV, e, flux, n_flux);
      orc_fft_conv(rb, prof, co);
      orc_rebin(e, flux, n_flux, g_econv, co, ORC_NCONV);
      for (int i = 0; i < n_flux; i++)
        if (e[i + 1] < 0.01 || e[i] > 1000.0) flux[i] = 0;
    }
    free_syspar(sp);
  } else { /* xillver_model, src/LocalModel.cpp:104-130, + add_primary_component, src/Relbase.cpp:294-351 */
    XPar src = {p.gam, p.afe, p.lxi, p.ect, p.dens, p.prim_type, p.xtab, p.ktbb, p.frac_pl_bb};
    if (!g_xill[p.xtab] && load_xill(p.xtab)) { free(e); return 3; }
    int nex = g_xill[p.xtab]->n_ener;
    double *ex = (double *) malloc(sizeof(double) * (nex + 1));
    double *fx = (double *) malloc(sizeof(double) * nex);
    xill_energy_grid(p.xtab, ex);
    rc = xillver_spectrum_incl(&src, p.xincl, fx);
    if (!rc) {
      double nf = 0.5 * cos(p.xincl * M_PI / 180); /* norm_xillver_spec, src/Xillspec.cpp:528-545 */
      for (int i = 0; i < nex; i++) fx[i] *= nf;
      orc_rebin(e, flux, n_flux, ex, fx, nex);
      double *pl = (double *) malloc(sizeof(double) * n_flux);
      primary_spectrum(pl, e, n_flux, &src, 1.0);
      double nsrc = norm_factor_source(&src);
      for (int i = 0; i < n_flux; i++) pl[i] *= nsrc;
      for (int i = 0; i < n_flux; i++) flux[i] *= fabs(p.refl_frac);
      if (p.refl_frac >= 0)
        for (int i = 0; i < n_flux; i++) flux[i] += pl[i];
      free(pl);
    }
    free(ex); free(fx);
  }
  free(e);
  return rc;
}

/* ------------------------------------------------------------------ stage probes */
int orc_syspar(const char *model, const double *par, double *re, double *gmin, double *gmax, double *emis,
               double *del_emit, double *del_inc, double *trff, double *cosne, double *frac) {
  const ModelDef *m = find_model(model);
  Par p;
  if (!m || interpret_params(m, par, &p)) return 1;
  SysPar *sp = new_syspar();
  int rc = system_parameters(&p, sp);
  if (!rc) {
    memcpy(re, sp->re, sizeof(sp->re)); memcpy(gmin, sp->gmin, sizeof(sp->gmin)); memcpy(gmax, sp->gmax, sizeof(sp->gmax));
    memcpy(emis, sp->emis, sizeof(sp->emis)); memcpy(del_emit, sp->del_emit, sizeof(sp->del_emit));
    memcpy(del_inc, sp->del_inc, sizeof(sp->del_inc));
    memcpy(trff, sp->trff, sizeof(double) * ORC_NR * ORC_NG * 2);
    memcpy(cosne, sp->cosne, sizeof(double) * ORC_NR * ORC_NG * 2);
    frac[0] = sp->refl_frac; frac[1] = sp->f_bh; frac[2] = sp->f_ad; frac[3] = sp->f_inf; frac[4] = sp->f_inf_rest;
  }
  free_syspar(sp);
  return rc;
}

int orc_relbase(const char *model, const double *par, const double *ener, int n_ener, double *flux) {
  const ModelDef *m = find_model(model);
  Par p;
  if (!m || interpret_params(m, par, &p)) return 1;
  SysPar *sp = new_syspar();
  int rc = system_parameters(&p, sp);
  if (!rc) {
    double rgrid[ORC_NZMAX + 1];
    zone_grid(p.rin, p.rout, p.num_zones, p.height, rgrid);
    rc = relline_profile(&p, sp, ener, n_ener, rgrid, 1, flux, NULL, 0);
  }
  free_syspar(sp);
  return rc;
}

int orc_relxill_stages(const char *model, const double *par, double *zone, double *zpar, double *corr,
                       double *normch, double *emis2, double *relflux, double *dist, double *xill,
                       int *n_ener_x, int *n_incl, double *conv, double *total) {
  const ModelDef *m = find_model(model);
  Par p;
  if (!m || m->type != T_RELXILL || interpret_params(m, par, &p)) return -1;
  Stages st;
  memset(&st, 0, sizeof(st));
  int rc = relxill_pipeline(&p, &st);
  if (rc) { free_stages(&st); return -rc; }
  int nz = st.nz;
  for (int i = 0; i <= nz; i++) zone[i] = st.zone[i];
  for (int i = 0; i < nz; i++) {
    zpar[i * 4] = st.lxi[i]; zpar[i * 4 + 1] = st.dens[i]; zpar[i * 4 + 2] = st.ect[i]; zpar[i * 4 + 3] = st.eshift[i];
    corr[2 * i] = st.corr_flux[i]; corr[2 * i + 1] = st.corr_gshift[i];
    normch[i] = st.normch[i];
  }
  memcpy(emis2, st.emis2, sizeof(st.emis2));
  memcpy(relflux, st.relflux, sizeof(double) * (size_t) nz * ORC_NCONV);
  memcpy(dist, st.dist, sizeof(double) * (size_t) nz * st.n_incl);
  memcpy(xill, st.xill, sizeof(double) * (size_t) nz * st.n_ener_x);
  *n_ener_x = st.n_ener_x; *n_incl = st.n_incl;
  memcpy(conv, st.conv, sizeof(st.conv));
  memcpy(total, st.total, sizeof(st.total));
  free_stages(&st);
  return nz;
}
