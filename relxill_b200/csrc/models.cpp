// models.cpp — model database and host-side parameter interpretation.
//
// Mirrors the behaviour of the reference's parameter layer for the models on the hot path:
//   parameter order/defaults: src/modelfiles/lmodel_relxill_public.dat
//   model -> (type, irradiation, primary spectrum): src/ModelDatabase.h:136-165
//   get_rel_params / get_xill_params / check_parameter_bounds: src/ModelDefinition.cpp:168-385
//   zone count: src/relutility.c:506-544;  zone grid: src/IonGradient.cpp:457-492
// Everything here is cheap scalar work; it runs on the host so that the values that drive
// discrete decisions on the device (ISCO, zone edges) have the host libm's bit patterns.
#include "models.h"

#include <cmath>
#include <cstring>

namespace rx {

static const ModelDef MODELS[] = {
    {"relline", "lmodrelline", T_LINE, EMIS_BKN, PRIM_NONE, 1, 10,
     {P_LINEE, P_INDEX1, P_INDEX2, P_RBR, P_A, P_INCL, P_RIN, P_ROUT, P_Z, P_LIMB},
     {6.4, 3., 3., 15.0, 0.998, 30., -1., 400., 0., 0.}},
    {"relconv", "lmodrelconv", T_CONV, EMIS_BKN, PRIM_NONE, 11, 8,
     {P_INDEX1, P_INDEX2, P_RBR, P_A, P_INCL, P_RIN, P_ROUT, P_LIMB},
     {3., 3., 15.0, 0.998, 30., -1., 400., 0.}},
    {"relline_lp", "lmodrellinelp", T_LINE, EMIS_LP, PRIM_NONE, 2, 10,
     {P_LINEE, P_H, P_A, P_INCL, P_RIN, P_ROUT, P_Z, P_LIMB, P_GAMMA, P_SWITCH_RETURNRAD},
     {6.4, 6.0, 0.998, 30., -1., 400., 0., 0., 2., 1.}},
    {"relconv_lp", "lmodrelconvlp", T_CONV, EMIS_LP, PRIM_NONE, 12, 9,
     {P_H, P_BETA, P_A, P_INCL, P_RIN, P_ROUT, P_LIMB, P_GAMMA, P_SWITCH_RETURNRAD},
     {6.0, 0.0, 0.998, 30., -1., 400., 0., 2., 1.}},
    {"relxill", "lmodrelxill", T_RELXILL, EMIS_BKN, PRIM_ECUT, -1, 13,
     {P_INDEX1, P_INDEX2, P_RBR, P_A, P_INCL, P_RIN, P_ROUT, P_Z, P_GAMMA, P_LOGXI, P_AFE, P_ECUT, P_REFLFRAC},
     {3., 3., 15.0, 0.998, 30., -1., 400., 0., 2., 3.1, 1., 300., 3.}},
    {"relxilllp", "lmodrelxilllp", T_RELXILL, EMIS_LP, PRIM_ECUT, -2, 14,
     {P_H, P_BETA, P_A, P_INCL, P_RIN, P_ROUT, P_Z, P_GAMMA, P_LOGXI, P_AFE, P_ECUT, P_REFLFRAC,
      P_SWITCH_RETURNRAD, P_SWITCH_REFLFRAC_BOOST},
     {6.0, 0.0, 0.998, 30., -1., 400., 0., 2., 3.1, 1., 300., 1.0, 1., 0.}},
    {"xillver", "lmodxillver", T_XILL, 0, PRIM_ECUT, 0, 7,
     {P_GAMMA, P_AFE, P_ECUT, P_LOGXI, P_Z, P_INCL, P_REFLFRAC},
     {2., 1., 300., 3.1, 0., 30., -1.}},
    {"xillverCp", "lmodxillverdensnthcomp", T_XILL, 0, PRIM_NTHCOMP, 100, 8,
     {P_GAMMA, P_AFE, P_KTE, P_LOGXI, P_LOGN, P_Z, P_INCL, P_REFLFRAC},
     {2., 1., 60., 3.1, 15., 0., 30., -1.}},
    {"relxillCp", "lmodrelxilldensnthcomp", T_RELXILL, EMIS_BKN, PRIM_NTHCOMP, -1, 14,
     {P_INCL, P_A, P_RIN, P_ROUT, P_RBR, P_INDEX1, P_INDEX2, P_Z, P_GAMMA, P_LOGXI, P_LOGN, P_AFE, P_KTE,
      P_REFLFRAC},
     {30., 0.998, -1., 400., 15.0, 3., 3., 0., 2., 3.1, 15., 1., 60., 3.}},
    {"relxilllpCp", "lmodrelxilllpdensnthcomp", T_RELXILL, EMIS_LP, PRIM_NTHCOMP, -2, 17,
     {P_INCL, P_A, P_RIN, P_ROUT, P_H, P_BETA, P_GAMMA, P_LOGXI, P_LOGN, P_AFE, P_KTE, P_REFLFRAC, P_Z,
      P_IONGRAD_INDEX, P_IONGRAD_TYPE, P_SWITCH_RETURNRAD, P_SWITCH_REFLFRAC_BOOST},
     {30., 0.998, -1., 400., 6.0, 0.0, 2., 3.1, 15., 1., 60., 1., 0., 0.0, 0., 1., 0.}},
    // neutron-star (blackbody-irradiated) and CO flavours: lmodel_relxill_public.dat:131-153, lmodel_relxill_devel.dat:1-25
    {"xillverNS", "lmodxillverns", T_XILL, 0, PRIM_BB, -101, 7,
     {P_KTBB, P_AFE, P_LOGN, P_LOGXI, P_Z, P_INCL, P_REFLFRAC},
     {2., 1., 15., 3.1, 0., 30., -1.}},
    {"relxillNS", "lmodrelxillns", T_RELXILL, EMIS_BKN, PRIM_BB, -30, 13,
     {P_INDEX1, P_INDEX2, P_RBR, P_A, P_INCL, P_RIN, P_ROUT, P_Z, P_KTBB, P_LOGXI, P_AFE, P_LOGN, P_REFLFRAC},
     {3., 3., 15.0, 0.998, 30., -1., 400., 0., 2., 3.1, 1., 15., 3.}},
    {"xillverCO", "lmodxillverco", T_XILL, 0, PRIM_ECUT, -210, 8,
     {P_GAMMA, P_ACO, P_KTBB, P_FRAC_PL_BB, P_ECUT, P_Z, P_INCL, P_REFLFRAC},
     {2., 5., 0.1, 0.01, 300., 0., 45., -1.}},
    {"relxillCO", "lmodrelxillco", T_RELXILL, EMIS_BKN, PRIM_ECUT, -200, 14,
     {P_INDEX1, P_INDEX2, P_RBR, P_A, P_INCL, P_RIN, P_ROUT, P_Z, P_GAMMA, P_ACO, P_KTBB, P_FRAC_PL_BB, P_ECUT,
      P_REFLFRAC},
     {3., 3., 15.0, 0.998, 30., -1., 400., 0., 2., 5., 0.1, 0.01, 300., 3.}},
};

int num_models() { return (int) (sizeof(MODELS) / sizeof(MODELS[0])); }
const ModelDef *model_at(int i) { return &MODELS[i]; }
const ModelDef *find_model(const char *name) {
  for (int i = 0; i < num_models(); i++)
    if (std::strcmp(MODELS[i].name, name) == 0 || std::strcmp(MODELS[i].symbol, name) == 0) return &MODELS[i];
  return nullptr;
}

static bool is_ns_model(int model_type) { return model_type == -30 || model_type == -101; }    // src/relutility.c:95-101
static bool is_co_model(int model_type) { return model_type == -200 || model_type == -210; }  // :103-109
int model_xtab(const ModelDef &m) {
  if (m.type != T_XILL && m.type != T_RELXILL) return XT_NONE;
  if (is_ns_model(m.model_type)) return XT_NS;
  if (is_co_model(m.model_type)) return XT_CO;
  return (m.prim == PRIM_NTHCOMP) ? XT_CP : XT_STD;
}

double kerr_rms(double a) {
  const double sign = (a < 0) ? -1.0 : 1.0;
  const double Z1 = 1.0 + std::pow(1.0 - a * a, 1.0 / 3.0) * (std::pow(1.0 + a, 1.0 / 3.0) + std::pow(1.0 - a, 1.0 / 3.0));
  const double Z2 = std::sqrt((3.0 * a * a) + (Z1 * Z1));
  return 3.0 + Z2 - sign * std::sqrt((3.0 - Z1) * (3.0 + Z1 + (2 * Z2)));
}
double kerr_rplus(double a) { return 1 + std::sqrt(1 - a * a); }

static int zone_count(int model_type, int emis_type, int ion_grad_type, int env) {
  if (ion_grad_type != ION_CONST) {
    if (env != 0 && env > 9 && env <= NZMAX) return env;
    return 25;
  }
  if (model_type < 0 && emis_type == EMIS_LP) {
    if (env != 0 && env > 0 && env <= NZMAX) return env;
    return 10;
  }
  return 1;
}

static int lower_index(const double *arr, int n, double val) {  // arr[k] <= val < arr[k+1], clamped
  int klo = 0, khi = n - 1;
  while (khi - klo > 1) {
    const int k = (khi + klo) / 2;
    if (arr[k] > val) khi = k; else klo = k;
  }
  return klo;
}

static void zone_grid(double rmin, double rmax, int nz, double h, double *rgrid) {
  if (nz == 1) {
    rgrid[0] = rmin;
    rgrid[1] = rmax;
    return;
  }
  double r_transition = rmin;
  int indr = 0;
  if (h > rmin) {
    r_transition = h;
    const double log_rmax = std::log(rmax), log_rmin = std::log(rmin);   // same bits as evaluating them per node
    for (int i = 0; i <= nz; i++) {
      rgrid[i] = 1.0 * i / (nz) * (log_rmax - log_rmin) + log_rmin;
      rgrid[i] = std::exp(rgrid[i]);
    }
    indr = lower_index(rgrid, nz + 1, r_transition);
    r_transition = rgrid[indr];
  }
  if (indr < nz) {
    const double rlo = r_transition, rhi = rmax;
    for (int i = indr; i < nz + 1; i++) {
      rgrid[i] = 1.0 * (i - indr) / (nz - indr) * (1.0 / rhi - 1.0 / rlo) + 1.0 / rlo;
      rgrid[i] = std::fabs(1.0 / rgrid[i]);
    }
  }
}

void interpret_params(const ModelDef &m, const double *par, const HostConfig &cfg, const double *rr_spins,
                      int rr_nspin, VPar &vp) {
  double v[P_COUNT];
  bool has[P_COUNT];
  for (int i = 0; i < P_COUNT; i++) { v[i] = 0.0; has[i] = false; }
  for (int i = 0; i < m.npar; i++) { v[m.ids[i]] = par[i]; has[m.ids[i]] = true; }
  std::memset(&vp, 0, sizeof(vp));
  vp.type = m.type;
  vp.model_type = m.model_type;
  vp.emis_type = m.irrad;
  vp.prim_type = m.prim;
  vp.status = ST_OK;
  vp.rr_spin = -1;
  vp.const_density = cfg.env_const_density;

  // xillver-side parameters
  vp.xtab = model_xtab(m);
  vp.afe = is_co_model(m.model_type) ? v[P_ACO] : v[P_AFE];   // src/ModelDefinition.cpp:347-349
  vp.ktbb = v[P_KTBB];
  vp.frac_pl_bb = v[P_FRAC_PL_BB];
  vp.ect = (m.prim == PRIM_NTHCOMP) ? (has[P_KTE] ? v[P_KTE] : 0.0) : (has[P_ECUT] ? v[P_ECUT] : 300.0);
  vp.lxi = has[P_LOGXI] ? v[P_LOGXI] : 0.0;
  vp.dens = has[P_LOGN] ? v[P_LOGN] : (is_co_model(m.model_type) ? 17.0 : 15.0);   // :362-363
  vp.iongrad_index = v[P_IONGRAD_INDEX];
  vp.gam = v[P_GAMMA];
  vp.refl_frac = v[P_REFLFRAC];
  vp.boost = (int) std::lround(has[P_SWITCH_REFLFRAC_BOOST] ? v[P_SWITCH_REFLFRAC_BOOST] : 0.0);

  vp.xincl = v[P_INCL];
  vp.z = v[P_Z];
  vp.eshift_obs = 1.0;
  vp.doppler_obs = 1.0;
  if (m.type == T_XILL) {  // no relativistic parameters (get_rel_params returns nullptr, src/ModelDefinition.cpp:281-283)
    vp.nz = 0;
    return;
  }

  // relativistic parameters
  vp.a = v[P_A];
  vp.incl = v[P_INCL] * M_PI / 180;
  vp.rin = v[P_RIN];
  vp.rout = v[P_ROUT];
  vp.emis1 = v[P_INDEX1];
  vp.emis2 = v[P_INDEX2];
  vp.rbr = v[P_RBR];
  vp.lineE = v[P_LINEE];
  vp.gamma = v[P_GAMMA];
  vp.height = v[P_H];
  vp.z = v[P_Z];
  vp.beta = v[P_BETA];
  vp.limb = (int) std::lround(v[P_LIMB]);
  {
    const int def = (m.irrad == EMIS_LP) ? 1 : 0;
    const int sw = (cfg.env_returnrad == 1) ? 1 : def;
    vp.return_rad = (int) std::lround(has[P_SWITCH_RETURNRAD] ? v[P_SWITCH_RETURNRAD] : (double) sw);
  }
  // do_renorm_model (src/relutility.c:603-623)
  if (m.model_type < 0) vp.renorm = (m.irrad == EMIS_LP || cfg.env_phys_norm) ? 0 : 1;
  else vp.renorm = cfg.env_phys_norm ? 0 : 1;

  // check_parameter_bounds
  bool bad = false;
  if (vp.rin < 0) vp.rin = -1.0 * vp.rin * kerr_rms(vp.a);
  if (vp.rout < 0) vp.rout = -1.0 * vp.rout * kerr_rms(vp.a);
  if (vp.rbr < 0) vp.rbr = -1.0 * vp.rbr * kerr_rms(vp.a);
  if (vp.rout <= vp.rin) bad = true;
  const double rms = kerr_rms(vp.a);
  if (vp.rin < rms) vp.rin = rms;
  if (vp.a > 0.9982 || vp.a < -1) bad = true;
  if (!(vp.a == vp.a)) bad = true;
  if (vp.incl < 3 * M_PI / 180 || vp.incl > 87 * M_PI / 180 || !(vp.incl == vp.incl)) bad = true;
  if (vp.rout <= vp.rin) bad = true;
  if (bad) {
    vp.status = ST_BAD_PARAM;
    return;
  }
  if (vp.rout > 1000.0) vp.rout = 1000.0;
  if (vp.emis_type == EMIS_BKN) {
    if (vp.rbr < vp.rin) vp.rbr = vp.rin;
    if (vp.rbr > vp.rout) vp.rbr = vp.rout;
  }
  if (vp.emis_type == EMIS_LP) {
    if (vp.beta < 0) vp.beta = 0.0;
    if (vp.beta > 0.99) vp.beta = 0.99;
    if (vp.height < 0) vp.height = -1.0 * vp.height * kerr_rplus(vp.a);
    const double h_fac = 1.1, r_event = kerr_rplus(vp.a);
    if ((h_fac * r_event - vp.height) > 1e-4) vp.height = r_event * h_fac;
  }
  vp.rms = rms;
  vp.ion_grad_type = (int) std::lround(has[P_IONGRAD_TYPE] ? v[P_IONGRAD_TYPE] : 0.0);
  if (vp.ion_grad_type < 0 || vp.ion_grad_type > 2) {
    vp.status = ST_BAD_PARAM;
    return;
  }
  vp.nz = zone_count(vp.model_type, vp.emis_type, vp.ion_grad_type, cfg.num_zones());

  // energy shift source -> observer (src/Relphysics.cpp:217-255)
  vp.eshift_obs = 1.0;
  vp.doppler_obs = 1.0;
  if (vp.emis_type == EMIS_LP) {
    const double g_inf_0 = std::sqrt(1.0 - (2 * vp.height / (vp.height * vp.height + vp.a * vp.a)));
    const double dop = std::sqrt(1.0 - vp.beta * vp.beta) / (1.0 + vp.beta * std::cos(M_PI - vp.incl));
    vp.eshift_obs = (vp.beta < 1e-4) ? g_inf_0 : g_inf_0 * dop;
    vp.doppler_obs = dop;
  }
  // cutoff energy is given in the observer frame for the lamp post: move it to the source frame
  // (src/Relxill.cpp:196-200)
  if (vp.type == T_RELXILL && vp.emis_type == EMIS_LP && vp.prim_type == PRIM_ECUT) vp.ect /= vp.eshift_obs;

  vp.relline_norm = 1.0;
  if (vp.model_type < 0 && vp.emis_type == EMIS_BKN) vp.relline_norm = 0.5 * std::cos((vp.incl * 180.0 / M_PI) * M_PI / 180);

  zone_grid(vp.rin, vp.rout, vp.nz, vp.height, vp.zone);

  // returning radiation: pick the next table spin >= a (src/Relreturn_Table.cpp:396-413)
  if (vp.return_rad != 0) {
    if (vp.return_rad != 1 && vp.return_rad != -1 && vp.return_rad != 2) {
      vp.status = ST_RRAD;
      return;
    }
    if (rr_spins == nullptr || rr_nspin < 1) {
      vp.status = ST_RRAD;
      return;
    }
    int k = (rr_nspin > 1) ? lower_index(rr_spins, rr_nspin, vp.a) : 0;
    if (rr_spins[k] < vp.a) k++;
    if (k >= rr_nspin) {
      vp.status = ST_RRAD;
      return;
    }
    vp.rr_spin = k;
  }
  vp.do_corr = (vp.type == T_RELXILL && vp.return_rad != 0 && vp.a > 0.0) ? 1 : 0;
}

int reusable_state(const VPar &p, const VPar &q) {
  if (p.type != q.type || p.model_type != q.model_type || p.status != q.status || q.status != ST_OK) return 0;
  if (p.type == T_XILL) return 0;
  bool rel = p.emis_type == q.emis_type && p.a == q.a && p.incl == q.incl && p.emis1 == q.emis1 && p.emis2 == q.emis2 &&
             p.rbr == q.rbr && p.rin == q.rin && p.rout == q.rout && p.lineE == q.lineE && p.height == q.height &&
             p.gamma == q.gamma && p.beta == q.beta && p.rms == q.rms && p.limb == q.limb && p.nz == q.nz &&
             p.return_rad == q.return_rad && p.rr_spin == q.rr_spin && p.do_corr == q.do_corr &&
             p.eshift_obs == q.eshift_obs && p.doppler_obs == q.doppler_obs;
  if (rel && p.type == T_LINE) rel = (p.z == q.z);   // the line models integrate on the caller's grid shifted by z
  for (int i = 0; rel && i <= q.nz; i++) rel = (p.zone[i] == q.zone[i]);
  if (!rel) return 0;
  int out = (p.do_corr || q.do_corr) ? 0 : REUSE_REL;
  if (p.type == T_RELXILL) {
    VPar a = p, b = q;   // interpret_params zeroes the struct first, so the padding compares equal
    a.z = b.z = 0.0;
    if (std::memcmp(&a, &b, sizeof(VPar)) == 0) out |= REUSE_ALL | REUSE_REL;
  }
  return out;
}

}  // namespace rx
