// kernels.h — launchers of the sm_100a kernels (kernels.cu, nthcomp.cu).
#pragma once
#include <cuda_runtime.h>

#include "common.h"

namespace rx {

int kernels_init();  // opt-in shared-memory sizes; returns 0 on success

void launch_syspar(const VPar *vps, const DevTables &T, const Scratch &S, long n, int pass, cudaStream_t st);
void launch_zone(const VPar *vps, const DevTables &T, const Scratch &S, long n, cudaStream_t st);
// n_incl > 0: also file the per-radius parts of the emission-angle distribution (relxill models)
void launch_fine(const VPar *vps, const DevTables &T, const Scratch &S, long n, int n_incl, double e_first,
                 double e_last, int store_cosne, int store_trff, cudaStream_t st);
void launch_dist(const VPar *vps, const DevTables &T, const Scratch &S, long n, int n_incl, cudaStream_t st);
// nz_min, nz_max: zone counts in the batch.  Vectors with few zones have their radii cut into line_parts(nz) runs, summed
// by a second kernel: the arena needs line_rows(nz_min, nz_max) profile rows per vector
int line_rows(int nz_min, int nz_max);
int line_launches(int nz_min, int nz_max);
void launch_line(const VPar *vps, const DevTables &T, const Scratch &S, long n, const double *egrid, int n_ener,
                 int grid_mode, int nz_min, int nz_max, cudaStream_t st);
int line_max_bins();  // largest energy grid the line kernel handles in one pass
void launch_linefinish(const VPar *vps, const Scratch &S, long n, int n_ener, double *out, cudaStream_t st);
// conv_grid != 0: zone spectra are filed (k_xill) and read (k_conv) on the convolution grid, see xill.cu
void launch_xill(const VPar *vps, const DevTables &T, const Scratch &S, long n, int which, int conv_grid, cudaStream_t st);
void xill_force_generic(int on);   // test hook: use k_xill's any-table instantiation (run-time row strides) on standard tables
void launch_conv(const VPar *vps, const DevTables &T, const Scratch &S, long n, const double *user_e, int n_flux,
                 double *out, double *total, int which, int mode, int conv_grid, int renorm3, cudaStream_t st);

void launch_xillver(const VPar *vps, const DevTables &T, const Scratch &S, long n, int which, const double *user_e,
                    int n_flux, double *out, int stride, cudaStream_t st);
void launch_xillver_prim_nth(const VPar *vps, const DevTables &T, const Scratch &S, long n, const double *user_e, int n_flux,
                             double *out, cudaStream_t st);
void launch_nth(const VPar *vps, const DevTables &T, const Scratch &S, long n, int nz_max, cudaStream_t st);   // 2 kernels
void launch_prim_nth(const VPar *vps, const DevTables &T, const Scratch &S, long n, double *total, const double *user_e,
                     int n_flux, double *out, int renorm3, cudaStream_t st);

}  // namespace rx
