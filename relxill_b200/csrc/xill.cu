// xill.cu — xillver table gather/blend kernel (the table-reading stage of the pipeline).
// Compiled with FMA contraction on (build.py): pure weighted sums, no discrete decisions.
//
// Replaces interp_5d_tab_incl / interp_6d_tab_incl for every inclination (src/xilltable.c:812-876,
// 999-1019) fused with calc_xillver_angdep (src/Xillspec.cpp:557-573) and the division by the
// normalisation change (src/Relxill.cpp:380-385):
//   xill[z][e] = sum_m dist[z][m] * sum_c w_c(z) * tab[node_c(z)][m][e] / normch[z].
//
// What makes this cheap on the GPU:
//  * factorised weights: the multilinear weight is a product over the table axes and Gamma / A_Fe do not
//    change between zones, so each "rest" corner (logXi, Ecut|kTe[, Dens]) is first contracted over the 4
//    (Gamma, A_Fe) nodes, and a zone blends 4 (5-D) or 8 (6-D) such contracted rows instead of 16 / 32;
//  * register-resident corners: a thread owns one energy bin and walks the zones in order; the contracted
//    rows H[slot][inclination] of the current bracket live in registers.  k_zone files every corner under
//    the slot given by the parities of its node indices, so when the bracket of the next zone moves by one
//    node only the slots whose node changed are re-read (neighbouring zones share most corners: constant-xi
//    lamp-post zones differ only in the Ecut node) — table traffic is the distinct rows, not zones x 16;
//  * the per-zone blend is ordered as sum_m dist[z][m] * (sum_s w_s H[s][m]): (slots + 1) FMAs per
//    inclination, and only slots + inclinations coefficients to fetch per zone (broadcast shared loads).
// 6-D tables (8 slots) split the inclinations over the two halves of a warp (two lanes per energy bin, one
// shuffle-add per zone) to keep the same register footprint.
// One CTA per (vector, tile of energy bins); table rows stream in as coalesced float loads, the zone spectra
// leave as coalesced double stores.
//  * CG (convolution-grid) variant.  The reference rebins every zone spectrum from the table grid onto the
//    convolution grid (_rebin_spectrum, src/relutility.c:549-601, per zone in src/Relxill.cpp:461-463).  That map
//    is linear and fixed, so it commutes with the blend: tables.cu applies it ONCE PER TABLE ROW at load and keeps
//    the rebinned rows in fp64 (XillDev::datac; fp64 so that nothing is rounded that the reference does not
//    round).  This kernel then blends straight on the convolution grid: same code, a thread owns a convolution
//    bin, the rows are doubles (no float -> double conversions, which bound the corner refresh on the table
//    grid), only the xc_n (about 2520 of 4096) bins that overlap the table exist, and k_conv reads its input with
//    plain coalesced loads — no rebin map, no dependent gathers.
#include <cuda_runtime.h>

#include "common.h"
#include "devutil.cuh"
#include "kernels.h"

namespace rx {

constexpr int XL_NT = 128;
constexpr int XL_NI = 10;   // inclination nodes handled (all xillver tables have 10; fewer are padded with zero weight)

// NS rest corners per zone, SPLIT lanes per energy bin; ST > 0: the table has XL_NI inclinations and rows of ST floats,
// so the loads of a corner refresh address [pointer + immediate] (the refresh is a third of the kernel's instructions
// when every row offset is 64-bit arithmetic on run-time strides); ST = 0: any table
template <bool CG> struct RowT { typedef float type; };
template <> struct RowT<true> { typedef double type; };

// ST > 0: row length of the table in use as a constant (3008 floats on the table grid, the xc_stride of the 2999-bin
// xillver grid on the convolution grid)
template <int NS, int SPLIT, int ST, bool CG>
__global__ void __launch_bounds__(XL_NT, 4) k_xill(const VPar *__restrict__ vps, DevTables T, Scratch S, int which,
                                                   int nz_stride) {
  typedef typename RowT<CG>::type row_t;
  constexpr int NIT = XL_NI / SPLIT;          // inclinations per lane
  constexpr int EPC = XL_NT / SPLIT;          // energy bins per CTA
  __shared__ __align__(16) double s_w[NZMAX * NS];
  __shared__ __align__(16) double s_dist[NZMAX * XL_NI];
  __shared__ double s_rn[NZMAX];
  __shared__ int s_off[NZMAX * NS];
  // tile-major launch order (blockIdx.x = vector): the CTAs resident on an SM at any time work on the SAME energy tile of
  // different vectors, so the table rows that vectors share (MCMC walkers sit in the same table cell) hit in L1
  const int v = blockIdx.x, tile = blockIdx.y, t = threadIdx.x;
  if (S.status[v] != ST_OK) return;
  if (S.reuse && (S.reuse[v] & REUSE_ALL)) return;
  const VPar &vp = vps[v];
  const int nz = vp.nz;
  const XillDev &X = T.xill[which];
  const int ni = ST ? XL_NI : X.n_incl;
  const int st = ST ? ST : (CG ? X.xc_stride : X.stride), ne = CG ? X.xc_n : X.n_ener;
  const row_t *xdata = CG ? reinterpret_cast<const row_t *>(X.datac) : reinterpret_cast<const row_t *>(X.data);
  // lane -> (energy bin, inclination part): with SPLIT = 2 the two halves of a warp share 16 bins
  const int lane = t & 31, warp = t >> 5;
  const int part = (SPLIT == 2) ? (lane >> 4) : 0;
  const int e = tile * EPC + ((SPLIT == 2) ? (warp * 16 + (lane & 15)) : t);
  const bool live = e < ne;
  for (int q = t; q < nz * NS; q += XL_NT) {
    const int z = q / NS, s = q - z * NS;
    s_off[q] = S.xkey[((size_t) v * NZMAX + z) * 8 + s];
    s_w[q] = S.xwsort[((size_t) v * NZMAX + z) * 8 + s];
  }
  for (int q = t; q < nz * XL_NI; q += XL_NT) {
    const int z = q / XL_NI, m = q - z * XL_NI;
    s_dist[q] = (m < ni) ? S.dist[((size_t) v * NZMAX + z) * MAX_INCL + m] : 0.0;
  }
  for (int z = t; z < nz; z += XL_NT) s_rn[z] = 1.0 / S.normch[(size_t) v * NZMAX + z];
  const row_t *base[4];
  double gaw[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    base[q] = xdata + (size_t) S.xga_off[(size_t) v * 4 + q] * ni * st + (live ? e : 0) + (size_t) (part * NIT) * st;
    gaw[q] = S.xga_w[(size_t) v * 4 + q];
  }
  __syncthreads();
  double H[NS][NIT];
  int cur[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) {
    cur[s] = -1;
#pragma unroll
    for (int m = 0; m < NIT; m++) H[s][m] = 0.0;
  }
  double *out = S.xillz + (size_t) v * nz_stride * st + e;
  for (int z = 0; z < nz; z++) {
    // refresh the slots whose rest corner changed (uniform over the CTA): contract the table over (Gamma, A_Fe)
#pragma unroll
    for (int s = 0; s < NS; s++) {
      const int off = s_off[z * NS + s];
      if (off != cur[s]) {
        cur[s] = off;
        const size_t ro = (size_t) off * ni * st;
        const row_t *p0 = base[0] + ro, *p1 = base[1] + ro, *p2 = base[2] + ro, *p3 = base[3] + ro;
#pragma unroll
        for (int m = 0; m < NIT; m++) {
          if (ST || part * NIT + m < ni) {
            const int o = m * st;
            H[s][m] = gaw[0] * (double) __ldg(p0 + o) + gaw[1] * (double) __ldg(p1 + o)
                      + gaw[2] * (double) __ldg(p2 + o) + gaw[3] * (double) __ldg(p3 + o);
          }
        }
      }
    }
    double w[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) w[s] = s_w[z * NS + s];
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < NIT; m++) {
      double g = w[0] * H[0][m];
#pragma unroll
      for (int s = 1; s < NS; s++) g += w[s] * H[s][m];
      acc += s_dist[z * XL_NI + part * NIT + m] * g;
    }
    if (SPLIT == 2) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (live && part == 0) out[(size_t) z * st] = acc * s_rn[z];
  }
}

int xill_kernel_init() { return 0; }

static bool g_xill_generic = false;   // test hook: run the any-table instantiation (run-time strides) on standard tables too
void xill_force_generic(int on) { g_xill_generic = on != 0; }

constexpr int XL_CST = 2528;   // xc_stride of the 2999-bin xillver grid (2521 convolution bins overlap it)

template <bool CG>
static void launch_xill_t(const VPar *vps, const DevTables &T, const Scratch &S, long n, int which, cudaStream_t st) {
  const XillDev &X = T.xill[which];
  // every table of the 2999-bin xillver grid has these row lengths
  const bool std_rows = !g_xill_generic && (X.n_incl == XL_NI) && (CG ? X.xc_stride == XL_CST : X.stride == 3008);
  constexpr int ST = CG ? XL_CST : 3008;
  const int nb = CG ? X.xc_n : X.n_ener;   // bins a vector's CTAs cover
  if (X.npar == 6) {
    dim3 grid((unsigned) n, (nb + XL_NT / 2 - 1) / (XL_NT / 2));
    if (std_rows) k_xill<8, 2, ST, CG><<<grid, XL_NT, 0, st>>>(vps, T, S, which, S.nz_cap);
    else k_xill<8, 2, 0, CG><<<grid, XL_NT, 0, st>>>(vps, T, S, which, S.nz_cap);
  } else {
    dim3 grid((unsigned) n, (nb + XL_NT - 1) / XL_NT);
    if (std_rows) k_xill<4, 1, ST, CG><<<grid, XL_NT, 0, st>>>(vps, T, S, which, S.nz_cap);
    else k_xill<4, 1, 0, CG><<<grid, XL_NT, 0, st>>>(vps, T, S, which, S.nz_cap);
  }
}

// conv_grid: blend the rebinned rows (XillDev::datac) and file the zone spectra on the convolution grid (rows of
// xc_stride) instead of the table grid (rows of stride); the caller checked that the table has them
void launch_xill(const VPar *vps, const DevTables &T, const Scratch &S, long n, int which, int conv_grid, cudaStream_t st) {
  // tables with more than XL_NI inclinations are rejected at load (tables.cu)
  if (conv_grid) launch_xill_t<true>(vps, T, S, n, which, st);
  else launch_xill_t<false>(vps, T, S, n, which, st);
}

}  // namespace rx
