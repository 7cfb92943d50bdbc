// common.h — data layout shared by the host runtime and the sm_100a kernels.
#pragma once
#include <cstddef>
#include <cstdint>

namespace rx {

// sizes fixed by the table formats / the reference (src/common.h:69-81,127-136; src/Xillspec.h:28-38)
constexpr int NR = 1000;       // fine radial grid
constexpr int NG = 40;         // g* grid
constexpr double CONV_EMIN = 0.00035, CONV_EMAX = 2000.0;   // convolution grid, src/Xillspec.h:36-38
constexpr int NCONV = 4096;    // convolution grid bins
constexpr int LINE_PARTS = 64; // partial rows of the line profile per vector at most (zones x runs, line.cu)
// runs the radii of a zone are cut into for the line profile (line.cu), a function of the vector's zone count alone
#ifdef __CUDACC__
__host__ __device__
#endif
inline int line_parts(int nz) { return nz == 1 ? 8 : (nz <= 8 ? 4 : 1); }
constexpr int NCOARSE = 500;   // xillver-normalisation grid bins
constexpr int NZMAX = 50;      // radial zones
constexpr int REL_NA = 25, REL_NMU = 30, REL_NRT = 100;
constexpr int LP_NA = 20, LP_NH = 250, LP_NRT = 100;
constexpr int RR_NR = 50, RR_NG = 20;
constexpr int MAX_INCL = 16;
constexpr int NTH_MAX = 900;   // nthcomp photon grid (src/donthcomp.c)
constexpr int NTH_SOL = 64;    // Kompaneets solves per vector: one per zone + one for the source (<= NZMAX + 1)

enum { EMIS_BKN = 1, EMIS_LP = 2 };
enum { PRIM_NONE = 0, PRIM_ECUT = 1, PRIM_NTHCOMP = 2, PRIM_BB = 3 };   // src/common.h:57-59
// xillver table flavours (get_xilltable_id, src/xilltable.c:682-694)
enum { XT_NONE = -1, XT_STD = 0, XT_CP = 1, XT_NS = 2, XT_CO = 3, XT_COUNT = 4 };
enum { T_LINE = 0, T_CONV = 1, T_XILL = 2, T_RELXILL = 3 };
enum { ION_CONST = 0, ION_PL = 1, ION_ALPHA = 2 };

enum { REUSE_REL = 1, REUSE_ALL = 2 };

// per-vector status codes (0 = ok)
enum {
  ST_OK = 0,
  ST_BAD_PARAM = 1,      // rejected by the parameter checks (reference: ParamInputException)
  ST_TABLE_RANGE = 2,    // radial grid outside the rel table
  ST_RRAD = 3,           // returning-radiation setup failed
  ST_ZONE = 4,           // degenerate zone grid
  ST_NAN = 5,            // NaN in the angular distribution
  ST_CONV_INPUT = 6      // convolution model with non-positive input flux
};

// One parameter vector after host-side interpretation (reference get_rel_params/get_xill_params/
// check_parameter_bounds, src/ModelDefinition.cpp:168-385).  Values that feed discrete decisions
// (zone grid, ISCO) are computed on the host with the host libm so they carry the same bits as
// the reference's.
struct VPar {
  double a, incl, emis1, emis2, rbr, rin, rout, lineE, z, height, gamma, beta;
  double gam, afe, lxi, ect, dens, refl_frac, iongrad_index;
  double ktbb, frac_pl_bb;      // blackbody temperature / power-law fraction axes of the NS and CO tables
  double eshift_obs;            // energy shift source -> observer (1 unless lamp post)
  double doppler_obs;           // doppler_factor_source_obs (lamp post, beta > 1e-4), else 1
  double rms;                   // ISCO
  double relline_norm;          // 0.5 cos(incl) for relxill+BKN, else 1
  double xincl;                 // inclination in degrees (standalone xillver models interpolate over it)
  double zone[NZMAX + 1];       // radial zone grid
  int model_type, emis_type, prim_type, type;
  int limb, nz, return_rad, ion_grad_type, boost;
  int renorm;                   // do_renorm_model
  int do_corr;                  // returning-radiation correction factors are computed
  int rr_spin;                  // index into the returning-radiation table (spin >= a), -1 = none
  int status;
  int const_density;            // RELXILL_CONSTANT_DENSITY=1: alpha-disk ionisation gradient with constant density
  int xtab;                     // xillver table of this model (XT_*)
};

struct XillDev {
  int npar;               // 5 or 6
  int nvals[6];
  int pindex[6];          // global parameter id of each table axis (0 gam,1 afe|a_co,2 lxi,3 ect/kte,4 dens,5 kTbb,6 frac,7 incl)
  const float *vals[6];   // device pointers to axis values
  int n_ener, n_incl, stride;  // stride = padded row length in floats
  long nnodes;            // rows / n_incl
  const float *data;      // [nnodes][n_incl][stride], renormalised like the reference does at load
  const double *ener;     // [n_ener+1] bin edges (float table values promoted)
  const float *incl;      // [n_incl] degrees
  // per-node scalars for the returning-radiation correction factors (linear functionals of the spectra)
  const double *node_ef, *node_p1, *node_p2;
  // fixed rebin map xillver grid -> convolution grid (same imin/imax/weights as _rebin_spectrum)
  const int *rb_ii;      // [NCONV][2] (imin, imax) of the rebin onto the convolution grid; (0, 0) with zero weights outside the source grid
  const double *rb_dd;   // [NCONV][2] (dmin, dmax) partial-overlap fractions of the first and last source bin
  // Convolution-grid copy of the table for the relxill models: every row rebinned at load onto the convolution bins
  // that overlap the table grid, [xc_first, xc_first + xc_n), in fp64, rows of xc_stride doubles (null: not built;
  // k_conv then rebins the zone spectra itself).  The rebin is linear, so it commutes with the blend (xill.cu)
  int xc_first, xc_n, xc_stride;
  const double *datac;   // [nnodes][n_incl][xc_stride]
};

struct DevTables {
  // relline table (src/reltable.h:24-48); r/gmin/gmax [na][nmu][100], tc [na][nmu][100][40] float4
  const float *rel_a, *rel_mu0, *rel_r, *rel_gmin, *rel_gmax;
  const float *rel_tc;  // float4 {trff1, trff2, cosne1, cosne2}
  // lamp-post table (src/reltable.h:51-69)
  const float *lp_a, *lp_h, *lp_rad, *lp_int, *lp_del, *lp_dinc;
  // returning radiation (src/Relreturn_Table.h:23-84) + ln of the g grid (precomputed)
  int rr_nspin;
  const double *rr_spin, *rr_rlo, *rr_rhi, *rr_tf, *rr_gmin, *rr_gmax;
  const double *rr_fgl;     // [nspin][RR_NG][RR_NR*RR_NR][2] {frac_g, ln g}
  // xillver tables, indexed by XT_STD (cutoff power law), XT_CP (nthcomp), XT_NS, XT_CO
  XillDev xill[XT_COUNT];
  // fixed grids
  const double *econv;      // [NCONV+1]
  const double *conv_cf;    // [NCONV] E_mid / dE
  int conv_b0, conv_b1;     // the normalisation band as a bin interval: 0.01 <= E_lo and E_hi < 1000 for b0 <= i <= b1
  int conv_i1kev;
  int conv_i3kev;           // convolution bin holding 3 keV (RELXILL_RENORMALIZE, src/Relxill.cpp:249-259)
  const double *ecoarse;    // [NCOARSE+1]
  const unsigned char *coarse_m1, *coarse_m2;  // masks of the two band conditions on the coarse grid
  const double *gstar, *d_gstar;  // [NG]
  const double *gstar_w;          // [NG] dg* / sqrt(g* - g*^2)
  const double *tw;               // FFT twiddles exp(-2 pi i m / NCONV), m < NCONV, interleaved (re, im)
  const double *conv_w;     // [NCONV/2+1][2] (re, im) DFT of the band mask (frequency-domain band sum of a convolution)
  // nthcomp: arrays that depend only on the photon grid (kT_bb is fixed at 0.05 keV)
  const double *nth_x, *nth_c2, *nth_rel, *nth_x3, *nth_w, *nth_dphdot;
  const double *nth_rw, *nth_xd, *nth_x4;   // 1 / w, x dphdot, x^4 per node
  const double *nth_w1, *nth_w2;   // weights of the two coarse-grid band integrals on the photon-grid nodes
  int nth_ih1;                     // f_spp__ bracket at 1 keV, z = 0
  double nth_xx1;
  int nth_jnr, nth_jrel, nth_jmaxth;
  double nth_xmin, nth_deltal;
};

// per-chunk device scratch (struct of arrays, vector-major).  Every array is listed once in SCRATCH_FIELDS with its
// element type and the number of elements per parameter vector (nzc / nec / nxs = the zone, line-grid and xillver-row
// capacities of the arena): api.cu allocates from the list and offsets from it when a piece of a batch works on its own
// slice of the arena, so the two cannot drift apart.
//   X(type, name, elements per vector)            NTH_X: allocated only for the nthcomp (Cp) models
//                                                 FINE_X: only when a batch files the fine transfer functions / emission
//                                                 angles (test probes, limb darkening): 1.3 MB per vector
#define SCRATCH_FIELDS(X, NTH_X, FINE_X)                                                                                         \
  X(double, re, NR) X(double, gmin, NR) X(double, gmax, NR) X(double, emis, NR)     /* fine radial grid quantities */     \
  X(double, del_emit, NR) X(double, del_inc, NR) X(double, fr, NR)                                                        \
  X(int, it, NR) X(int, izone, NR)                                                                                        \
  X(int, zfirst, NZMAX + 1)        /* first fine-grid index with zone < z */                                              \
  X(int, brk_i, 2)                 /* rel-table bracket (spin, mu0) */                                                    \
  X(double, brk_f, 2)              /* its interpolation factors */                                                        \
  X(double, reflfrac, 8)                                                                                                  \
  FINE_X(double, trff, (size_t) NR * NG * 2) FINE_X(double, cosne, (size_t) NR * NG * 2)   /* [NR][NG][2]: probes, limb */ \
  X(double, relrow, (size_t) REL_NRT * NG * 4)   /* table rows interpolated in (a, mu0): trff1,2, cosne1,2 */             \
  X(double, eshift, NZMAX) X(double, zlxi, NZMAX) X(double, zdens, NZMAX) X(double, zect, NZMAX)                          \
  X(double, normch, NZMAX) X(double, corr_flux, NZMAX) X(double, corr_gshift, NZMAX)                                      \
  X(double, nsrc, 1)               /* source normalisation factor */                                                      \
  X(int, xrow, NZMAX * 32)         /* node index of each corner */                                                        \
  X(double, xw, NZMAX * 32)        /* corner weights */                                                                   \
  X(int, xkey, NZMAX * 32)         /* rest-corner node offset per (zone, slot) */                                         \
  X(int, xga_off, 4)               /* node offsets of the (Gamma, A_Fe) corners */                                        \
  X(double, xga_w, 4)              /* their weights */                                                                    \
  X(double, xwsort, NZMAX * 32)    /* weights in xkey order */                                                            \
  X(int, xn, 1)                    /* number of rest corners per zone */                                                  \
  X(double, relflux, (size_t) nzc * nec)   /* [nz_cap][ne_line_cap] (valid inside zrange only) */                         \
  X(int, zrange, NZMAX * 2)        /* first/last bin written per zone (-1: none) */                                       \
  X(int, zrpart, LINE_PARTS * 2)   /* the same per partial row of a split zone (k_line -> k_linemerge) */                 \
  X(double, dist, NZMAX * MAX_INCL)                                                                                       \
  X(double, distpart, NR * 10)     /* per-radius parts of dist (k_fine -> k_dist) */                                      \
  X(double, xillz, (size_t) nzc * nxs)     /* zone spectra: rows of XillDev::xc_stride (convolution grid) or ::stride */  \
  X(int, status, 1)                                                                                                       \
  X(double, total, NCONV)          /* convolution-grid spectrum of the vector (state cache, probes) */                    \
  NTH_X(double, nth_spt, NTH_MAX)  /* E F_E solution of the SOURCE (the zones' are consumed inside k_nth) */              \
  NTH_X(int, nth_jmax, NTH_SOL) NTH_X(double, nth_nfac, NTH_SOL) NTH_X(double, nth_s2, NTH_SOL)

struct Scratch {
  long cap;          // vectors
  int nz_cap;        // zones per vector allocated
  int ne_line_cap;   // bins of the line-profile grid allocated
  int nex_stride;    // xillver row stride
#define RX_DECL(type, name, count) type *name;
  SCRATCH_FIELDS(RX_DECL, RX_DECL, RX_DECL)
#undef RX_DECL
  // Re-use of the previous run's device-resident state (api.cu: the arena still holds this batch): per vector,
  // REUSE_REL = the relativistic half (k_syspar, k_fine, k_dist, k_line outputs) is still valid, REUSE_ALL = the
  // whole convolution-grid spectrum is.  Null when nothing can be re-used.
  const unsigned char *reuse;                                 // [cap]
};

}  // namespace rx
