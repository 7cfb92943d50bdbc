/* ref_probe.cpp — extern "C" probes compiled INTO oracle/_ref/librelxill_ref.so next to the
 * unmodified reference objects.  They only call the reference's own functions (the same
 * way its unit tests do, test/unit/common-functions.cpp:46-111) so that Python tests can
 *   (1) evaluate any local model with exceptions turned into a status code, and
 *   (2) read the reference's intermediate products of the relxill pipeline
 *       (src/Relxill.cpp:289-413) stage by stage.
 * Test infrastructure only: nothing in relxill_b200/ links or loads this.
 */
#include "LocalModel.h"
#include "Relxill.h"
#include "Relbase.h"
#include "Relprofile.h"
#include "IonGradient.h"
#include "PrimarySource.h"
#include "Xillspec.h"
#include "XspecSpectrum.h"

#include <cstring>
#include <string>

extern "C" {

/* full model through the reference's public path; returns 0 on success */
int ref_eval_model(const char *xspec_name, const double *energy, int n_flux, const double *par, double *flux) {
  try {
    ModelName name = ModelDatabase::instance().model_name(std::string(xspec_name));
    LocalModel lm{par, name};
    XspecSpectrum spec{energy, flux, static_cast<size_t>(n_flux)};
    lm.eval_model(spec);
  } catch (std::exception &e) {
    return 1;
  }
  return 0;
}

int ref_num_params(const char *xspec_name) {
  try {
    ModelName name = ModelDatabase::instance().model_name(std::string(xspec_name));
    return static_cast<int>(ModelDatabase::instance().param_list(name).num_params());
  } catch (std::exception &e) {
    return -1;
  }
}

int ref_default_params(const char *xspec_name, double *out) {
  try {
    ModelName name = ModelDatabase::instance().model_name(std::string(xspec_name));
    auto v = ModelDatabase::instance().get_default_values_array(name);
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return static_cast<int>(v.size());
  } catch (std::exception &e) {
    return -1;
  }
}

/* interpreted parameters: rel[0..11] = a, incl, emis1, emis2, rbr, rin, rout, lineE, z, height, gamma, beta;
 * irel[0..5] = model_type, emis_type, limb, num_zones, return_rad, ion_grad_type */
int ref_rel_params(const char *xspec_name, const double *par, double *rel, int *irel) {
  try {
    ModelName name = ModelDatabase::instance().model_name(std::string(xspec_name));
    LocalModel lm{par, name};
    relParam *p = lm.get_rel_params();
    if (p == nullptr) return 2;
    double r[12] = {p->a, p->incl, p->emis1, p->emis2, p->rbr, p->rin, p->rout, p->lineE, p->z, p->height,
                    p->gamma, p->beta};
    memcpy(rel, r, sizeof(r));
    int ir[6] = {p->model_type, p->emis_type, p->limb, p->num_zones, p->return_rad, p->ion_grad_type};
    memcpy(irel, ir, sizeof(ir));
    delete p;
  } catch (std::exception &e) {
    return 1;
  }
  return 0;
}

/* system parameters (src/Relprofile.cpp:310-358): arrays of N_FRAD(=1000) radii;
 * trff/cosne are [nr][ng][2] flattened; frac = {refl_frac,f_bh,f_ad,f_inf,f_inf_rest} (LP only) */
int ref_syspar(const char *xspec_name, const double *par, double *re, double *gmin, double *gmax, double *emis,
               double *del_emit, double *del_inc, double *trff, double *cosne, double *frac) {
  try {
    ModelName name = ModelDatabase::instance().model_name(std::string(xspec_name));
    LocalModel lm{par, name};
    relParam *p = lm.get_rel_params();
    int status = EXIT_SUCCESS;
    RelSysPar *sp = get_system_parameters(p, &status);
    if (status != EXIT_SUCCESS || sp == nullptr) { delete p; return 3; }
    for (int i = 0; i < sp->nr; i++) {
      re[i] = sp->re[i];
      gmin[i] = sp->gmin[i];
      gmax[i] = sp->gmax[i];
      emis[i] = sp->emis->emis[i];
      del_emit[i] = sp->emis->del_emit[i];
      del_inc[i] = sp->emis->del_inc[i];
      for (int j = 0; j < sp->ng; j++)
        for (int k = 0; k < 2; k++) {
          trff[(i * sp->ng + j) * 2 + k] = sp->trff[i][j][k];
          cosne[(i * sp->ng + j) * 2 + k] = sp->cosne[i][j][k];
        }
    }
    if (sp->emis->photon_fate_fractions != nullptr) {
      lpReflFrac *f = sp->emis->photon_fate_fractions;
      frac[0] = f->refl_frac; frac[1] = f->f_bh; frac[2] = f->f_ad; frac[3] = f->f_inf; frac[4] = f->f_inf_rest;
    }
    delete p;
  } catch (std::exception &e) {
    return 1;
  }
  return 0;
}

/* line profile on an arbitrary grid through relbase() (src/Relbase.cpp:524-538): one zone, no angular
 * distribution; `ener` is the grid the line is computed on (already divided by lineE) */
int ref_relbase(const char *xspec_name, const double *par, const double *ener, int n_ener, double *flux) {
  try {
    ModelName name = ModelDatabase::instance().model_name(std::string(xspec_name));
    LocalModel lm{par, name};
    relParam *p = lm.get_rel_params();
    int status = EXIT_SUCCESS;
    std::vector<double> e(ener, ener + n_ener + 1);
    relline_spec_multizone *spec = relbase(e.data(), n_ener, p, &status);
    if (status != EXIT_SUCCESS || spec == nullptr) { delete p; return 3; }
    for (int i = 0; i < n_ener; i++) flux[i] = spec->flux[0][i];
    delete p;
  } catch (std::exception &e) {
    return 1;
  }
  return 0;
}

/* The stages of relxill_kernel (src/Relxill.cpp:302-398) replayed with the reference's own functions, keeping
 * the intermediates.  Output buffers (caller-allocated, sized for N_ZONES_MAX=50 zones):
 *   zone[(nz+1)] radial grid; zpar[nz*4] = lxi, dens, ect, energy shift per zone;
 *   corr[nz*2] = corrfac_flux, corrfac_gshift (zeros if not computed); normch[nz];
 *   emis2[1000] emissivity of the second get_system_parameters call;
 *   relflux[nz*4096], dist[nz*n_incl]; xill[nz*n_ener_x] angle-weighted and divided by norm change;
 *   conv[4096] = sum of convolved zones before the primary is added; total[4096] after add_primary_spectrum.
 * returns number of zones (>0) or a negative error. */
int ref_relxill_stages(const char *xspec_name, const double *par, double *zone, double *zpar, double *corr,
                       double *normch, double *emis2, double *relflux, double *dist, double *xill,
                       int *n_ener_x, int *n_incl, double *conv, double *total) {
  try {
    ModelName name = ModelDatabase::instance().model_name(std::string(xspec_name));
    LocalModel lm{par, name};
    const ModelDefinition &params = lm.get_model_params();
    int status = EXIT_SUCCESS;

    relParam *rel_param = get_rel_params(params);
    xillParam *xill_param = get_xill_params(params);
    if (rel_param->emis_type == EMIS_TYPE_LP && xill_param->prim_type == PRIM_SPEC_ECUT) {
      xill_param->ect /= energy_shift_source_obs(rel_param);
    }
    specCache *spec_cache = init_global_specCache(&status);
    RelSysPar *sys_par = get_system_parameters(rel_param, &status);
    if (status != EXIT_SUCCESS) return -3;
    auto primary_source = PrimarySource(params, sys_par);

    IonGradient ion_gradient{RadialGrid(rel_param->rin, rel_param->rout, rel_param->num_zones, rel_param->height),
                             rel_param->ion_grad_type, xill_param->iongrad_index};
    ion_gradient.calculate_gradient(*(sys_par->emis), primary_source.source_parameters);
    const int nz = ion_gradient.nzones();
    auto xill_param_zone = ion_gradient.get_xill_param_zone(primary_source.source_parameters.xilltab_param());
    for (int i = 0; i <= nz; i++) zone[i] = ion_gradient.radial_grid.radius[i];
    for (int i = 0; i < nz; i++) {
      zpar[i * 4 + 0] = xill_param_zone[i]->lxi;
      zpar[i * 4 + 1] = xill_param_zone[i]->dens;
      zpar[i * 4 + 2] = xill_param_zone[i]->ect;
      zpar[i * 4 + 3] = ion_gradient.m_energy_shift_source_disk[i];
    }

    std::vector<xillSpec *> xspec(nz, nullptr);
    for (int i = 0; i < nz; i++) {
      xspec[i] = get_xillver_spectra_table(xill_param_zone[i], &status);
      if (status != EXIT_SUCCESS) return -4;
    }
    *n_ener_x = xspec[0]->n_ener;
    *n_incl = xspec[0]->n_incl;

    for (int i = 0; i < nz; i++) { corr[2 * i] = 0.0; corr[2 * i + 1] = 0.0; }
    rel_param->rrad_corr_factors =
        (rel_param->return_rad != 0 && rel_param->a > SPIN_MIN_RRAD_CALC_CORRFAC)
        ? calc_rrad_corr_factors(xspec.data(), ion_gradient.radial_grid, xill_param_zone, &status)
        : nullptr;
    if (rel_param->rrad_corr_factors != nullptr) {
      for (int i = 0; i < nz; i++) {
        corr[2 * i] = rel_param->rrad_corr_factors->corrfac_flux[i];
        corr[2 * i + 1] = rel_param->rrad_corr_factors->corrfac_gshift[i];
      }
    }
    sys_par = get_system_parameters(rel_param, &status);
    if (status != EXIT_SUCCESS) return -5;
    for (int i = 0; i < sys_par->nr; i++) emis2[i] = sys_par->emis->emis[i];

    xillTable *xill_tab = nullptr;
    get_init_xillver_table(&xill_tab, xill_param->model_type, xill_param->prim_type, &status);
    RelxillSpec relxill_spec;
    relline_spec_multizone *rel_profile =
        relbase_profile(relxill_spec.energy(), static_cast<int>(relxill_spec.num_flux_bins), rel_param, sys_par,
                        xill_tab, ion_gradient.radial_grid.radius.data(), nz, &status);
    if (status != EXIT_SUCCESS) return -6;
    const int ne = rel_profile->n_ener;
    for (int i = 0; i < nz; i++) {
      for (int j = 0; j < ne; j++) relflux[i * ne + j] = rel_profile->flux[i][j];
      for (int j = 0; j < *n_incl; j++) dist[i * (*n_incl) + j] = rel_profile->rel_cosne->dist[i][j];
    }

    auto zones_spec = SpectrumZones(xspec[0]->ener, xspec[0]->n_ener, nz);
    for (int i = 0; i < nz; i++) {
      calc_xillver_angdep(zones_spec.flux[i], xspec[i], rel_profile->rel_cosne->dist[i], &status);
    }
    double *nc = calc_xillver_normalization_change_source_to_disk(
        ion_gradient.m_energy_shift_source_disk, nz, primary_source.source_parameters.xilltab_param());
    for (int i = 0; i < nz; i++) {
      normch[i] = nc[i];
      for (int j = 0; j < zones_spec.num_flux_bins; j++) {
        zones_spec.flux[i][j] /= nc[i];
        xill[i * zones_spec.num_flux_bins + j] = zones_spec.flux[i][j];
      }
    }
    delete[] nc;

    /* convolution: same loop as relxill_convolution_multizone (src/Relxill.cpp:432-482), always recomputing */
    RelxillSpec rebinned, conv_out;
    for (int j = 0; j < ne; j++) relxill_spec.flux[j] = 0.0;
    for (int i = 0; i < nz; i++) {
      if (calcSum(rel_profile->flux[i], ne) < 1e-12) continue;
      _rebin_spectrum(rebinned.energy(), rebinned.flux, ne, zones_spec.energy(), zones_spec.flux[i],
                      zones_spec.num_flux_bins);
      convolveSpectrumFFTNormalized(rebinned.energy(), rebinned.flux, rel_profile->flux[i], conv_out.flux, ne, 1, 1,
                                    i, spec_cache, &status);
      for (int j = 0; j < ne; j++) relxill_spec.flux[j] += conv_out.flux[j];
    }
    for (int j = 0; j < ne; j++) conv[j] = relxill_spec.flux[j];
    primary_source.add_primary_spectrum(relxill_spec);
    for (int j = 0; j < ne; j++) total[j] = relxill_spec.flux[j];

    for (int i = 0; i < nz; i++) {
      free_xill_spec(xspec[i]);
      delete xill_param_zone[i];
    }
    delete[] xill_param_zone;
    free_rrad_corr_factors(&(rel_param->rrad_corr_factors));
    delete rel_param;
    delete xill_param;
    return nz;
  } catch (std::exception &e) {
    return -1;
  }
}

/* fixed grids (bit patterns produced by this host's libm, as the reference would) */
void ref_conv_grid(double *ener /*4097*/) {
  EnerGrid *g = get_relxill_conv_energy_grid();
  for (int i = 0; i <= g->nbins; i++) ener[i] = g->ener[i];
}

void ref_rebin(const double *ener, double *flu, int n, const double *ener0, const double *flu0, int n0) {
  _rebin_spectrum(ener, flu, n, ener0, flu0, n0);
}

void ref_fft_conv(const double *fxill, const double *frel, double *fout) {
  int status = EXIT_SUCCESS;
  specCache *c = init_global_specCache(&status);
  EnerGrid *g = get_relxill_conv_energy_grid();
  convolveSpectrumFFTNormalized(g->ener, fxill, frel, fout, g->nbins, 1, 1, 0, c, &status);
}

void ref_nthcomp(const double *ener, int n, double gamma, double kte, double z, double *out) {
  double prm[5];
  get_nthcomp_param(prm, gamma, kte, z);
  c_donthcomp(ener, n, prm, out);
}

double ref_kerr_rms(double a) { return kerr_rms(a); }

}  // extern "C"
