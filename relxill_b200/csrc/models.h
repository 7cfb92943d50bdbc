// models.h — host-side mirror of the reference's model database and parameter layer
// (src/ModelDatabase.h, src/ModelDefinition.{h,cpp}, src/modelfiles/lmodel_relxill_public.dat).
#pragma once
#include "common.h"

namespace rx {

enum XPar {
  P_LINEE, P_INDEX1, P_INDEX2, P_RBR, P_A, P_RIN, P_ROUT, P_INCL, P_Z, P_LIMB, P_GAMMA, P_LOGXI, P_LOGN, P_AFE,
  P_ECUT, P_KTE, P_REFLFRAC, P_H, P_BETA, P_IONGRAD_INDEX, P_IONGRAD_TYPE, P_SWITCH_RETURNRAD,
  P_SWITCH_REFLFRAC_BOOST, P_KTBB, P_ACO, P_FRAC_PL_BB, P_COUNT
};

struct ModelDef {
  const char *name;      // XSPEC name (lmodel.dat column 1)
  const char *symbol;    // C symbol after "c_" (lmodel.dat column 5)
  int type, irrad, prim; // T_Model / T_Irrad / T_PrimSpec of src/ModelDatabase.h:136-165
  int model_type;        // integer model type, src/ModelDefinition.cpp:35-61
  int npar;
  int ids[20];
  double def[20];
};

// The reference reads its environment switches on EVERY evaluation (get_num_zones, src/relutility.c:506-544;
// do_not_normalize_relline, :386-396; constantDiskDensity, :372-382; get_returnrad_switch, src/ModelDefinition.cpp:139-149;
// do_renorm_relxill, src/Relxill.cpp:241-247), so api.cu refreshes this struct at the top of every call that interprets
// parameters (read_call_env).
struct HostConfig {
  int env_num_zones = 0;       // RELXILL_NUM_RZONES (0 = unset)
  int override_num_zones = 0;  // relxill_b200_set_num_zones(n > 0): takes precedence over the environment variable
  int env_returnrad = -1;      // RELXILL_RETURNRAD_SWITCH (-1 = unset)
  int env_phys_norm = 0;       // RELLINE_PHYSICAL_NORM
  int env_const_density = 0;   // RELXILL_CONSTANT_DENSITY
  int env_renorm_relxill = 0;  // RELXILL_RENORMALIZE: relxill spectra rescaled to 1 cts/s/keV/cm2 at 3 keV before the final rebin
  int num_zones() const { return override_num_zones > 0 ? override_num_zones : env_num_zones; }
};

const ModelDef *find_model(const char *name);
// xillver table a model reads (XT_NONE for the line / convolution models)
int model_xtab(const ModelDef &m);
int num_models();
const ModelDef *model_at(int i);

// reference kerr_rms / kerr_rplus (src/Relphysics.cpp:139-155)
double kerr_rms(double a);
double kerr_rplus(double a);

// Fill a VPar from one raw parameter vector.  `rr_spins` = spin axis of the returning-radiation
// table (may be null if that table is not loaded).  Sets vp.status (ST_OK / ST_BAD_PARAM).
void interpret_params(const ModelDef &m, const double *par, const HostConfig &cfg, const double *rr_spins,
                      int rr_nspin, VPar &vp);

// What of a vector's device-resident state survives a parameter change — the counterpart of the reference's
// did_rel_param_change / did_xill_param_change logic (CachingStatus, src/Relxill.cpp:316; comp_rel_param and
// comp_xill_param, src/Relbase.cpp:353-470):
//   REUSE_ALL  every interpreted value that reaches the convolution-grid spectrum is unchanged (z may differ for the
//              relxill flavours: it only enters the final rebin);
//   REUSE_REL  the values read by the relativistic half (k_syspar, k_fine, k_dist, k_line) are unchanged and the
//              emissivity does not depend on the xillver side (no returning-radiation correction factors).
// Returns a combination of REUSE_REL / REUSE_ALL (0: recompute everything).
int reusable_state(const VPar &prev, const VPar &now);

}  // namespace rx
