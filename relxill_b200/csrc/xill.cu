// xill.cu — xillver table gather/blend kernel (the HBM-side stage of the pipeline).
// Compiled with FMA contraction on (build.py): pure weighted sums, no discrete decisions.
//
// Replaces interp_5d_tab_incl / interp_6d_tab_incl for every inclination (src/xilltable.c:812-876,
// 999-1019) fused with calc_xillver_angdep (src/Xillspec.cpp:557-573) and the division by the
// normalisation change (src/Relxill.cpp:380-385):
//   xill[z][e] = sum_m dist[z][m] * sum_c w_c(z) * tab[node_c(z)][m][e] / normch[z].
//
// Two things make this cheap on the GPU:
//  * corner-stationary: the zones of a vector share most interpolation corners (constant-xi lamp-post
//    zones differ only in the Ecut node), so the (corner, zone, weight) triples are sorted by corner
//    (k_zone) and every distinct table row is read ONCE per (vector, energy tile) and scattered to all
//    zones that use it — HBM/L2 traffic is the distinct rows, not zones x 16 rows;
//  * factorised weights: the multilinear weight is a product over the table axes and Gamma / A_Fe do not
//    change between zones, so each distinct "rest" corner (logXi, Ecut|kTe[, Dens]) is first contracted
//    over the 4 (Gamma, A_Fe) nodes in registers, and zones then blend 4 (5-D) or 8 (6-D) such rows
//    instead of 16 / 32: ~4x fewer FMAs.
// One CTA per (vector, tile of 256 energy bins); per-zone accumulators live in shared memory (one column
// per thread: conflict-free); table rows stream through as coalesced float loads.
#include <cuda_runtime.h>

#include "common.h"
#include "devutil.cuh"
#include "kernels.h"

namespace rx {

constexpr int XL_NT = 128;
constexpr int XL_CH = 64;    // entries staged per chunk

template <int NI>   // NI = number of inclinations known at compile time (0: run-time)
__global__ void __launch_bounds__(XL_NT) k_xill(const VPar *__restrict__ vps, DevTables T, Scratch S, int which,
                                                int nz_stride, int nz_max) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double *acc = reinterpret_cast<double *>(smraw);                 // [nz_max][XL_NT]
  double *s_dist = acc + (size_t) nz_max * XL_NT;                 // [nz_max][MAX_INCL]
  double *s_cw = s_dist + (size_t) nz_max * MAX_INCL;             // [XL_CH][MAX_INCL]  w * dist[z][m] of the staged entries
  int *s_key = reinterpret_cast<int *>(s_cw + XL_CH * MAX_INCL);  // [XL_CH]
  const int v = blockIdx.y, t = threadIdx.x;
  if (S.status[v] != ST_OK) return;
  const VPar &vp = vps[v];
  const int nz = vp.nz;
  const XillDev &X = T.xill[which];
  const int ni = NI ? NI : X.n_incl;
  const int st = X.stride, ne = X.n_ener;
  const int e = blockIdx.x * XL_NT + t;
  const bool live = e < ne;
  for (int z = 0; z < nz; z++) acc[z * XL_NT + t] = 0.0;
  for (int q = t; q < nz * MAX_INCL; q += XL_NT) s_dist[q] = S.dist[(size_t) v * NZMAX * MAX_INCL + q];
  const int n_ent = S.xn[v];
  const int *gk = S.xkey + (size_t) v * NZMAX * 32;
  const double *gw = S.xwsort + (size_t) v * NZMAX * 32;
  const float *base[4];
  double gaw[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    base[q] = X.data + (size_t) S.xga_off[(size_t) v * 4 + q] * ni * st + (live ? e : 0);
    gaw[q] = S.xga_w[(size_t) v * 4 + q];
  }
  int cur = -1;
  double H[NI ? NI : MAX_INCL];
  for (int c0 = 0; c0 < n_ent; c0 += XL_CH) {
    __syncthreads();
    const int nchunk = min(XL_CH, n_ent - c0);
    if (t < nchunk) s_key[t] = gk[c0 + t];
    for (int q = t; q < nchunk * ni; q += XL_NT) {   // combined coefficient of entry i and inclination m
      const int i = q / ni, m = q - i * ni;
      const int z = gk[c0 + i] & 63;
      s_cw[i * MAX_INCL + m] = gw[c0 + i] * s_dist[z * MAX_INCL + m];
    }
    __syncthreads();
    if (!live) continue;
    for (int i = 0; i < nchunk; i++) {
      const int key = s_key[i];
      const int off = key >> 6, z = key & 63;
      if (off != cur) {   // new rest corner: contract the table over the (Gamma, A_Fe) nodes
        cur = off;
        const size_t ro = (size_t) off * ni * st;
#pragma unroll
        for (int m = 0; m < (NI ? NI : MAX_INCL); m++) {
          if (NI || m < ni) {
            const size_t o = ro + (size_t) m * st;
            H[m] = gaw[0] * (double) __ldg(base[0] + o) + gaw[1] * (double) __ldg(base[1] + o)
                   + gaw[2] * (double) __ldg(base[2] + o) + gaw[3] * (double) __ldg(base[3] + o);
          }
        }
      }
      const double *cw = s_cw + i * MAX_INCL;
      double sacc = 0.0;
#pragma unroll
      for (int m = 0; m < (NI ? NI : MAX_INCL); m++)
        if (NI || m < ni) sacc += cw[m] * H[m];
      acc[z * XL_NT + t] += sacc;
    }
  }
  if (!live) return;
  for (int z = 0; z < nz; z++)
    S.xillz[((size_t) v * nz_stride + z) * st + e] = acc[z * XL_NT + t] / S.normch[(size_t) v * NZMAX + z];
}

static size_t xill_smem(int nz_max) {
  return ((size_t) nz_max * XL_NT + (size_t) nz_max * MAX_INCL + (size_t) XL_CH * MAX_INCL) * sizeof(double) + XL_CH * sizeof(int);
}

int xill_kernel_init() {
  cudaError_t e;
  e = cudaFuncSetAttribute(k_xill<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) xill_smem(NZMAX));
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(k_xill<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) xill_smem(NZMAX));
  if (e != cudaSuccess) return 1;
  cudaFuncSetAttribute(k_xill<10>, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_xill<0>, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared);
  return 0;
}

void launch_xill(const VPar *vps, const DevTables &T, const Scratch &S, long n, int which, int nz_max, int n_ener,
                 int n_incl, cudaStream_t st) {
  dim3 grid((n_ener + XL_NT - 1) / XL_NT, (unsigned) n);
  const size_t sm = xill_smem(nz_max);
  if (n_incl == 10) k_xill<10><<<grid, XL_NT, sm, st>>>(vps, T, S, which, S.nz_cap, nz_max);
  else k_xill<0><<<grid, XL_NT, sm, st>>>(vps, T, S, which, S.nz_cap, nz_max);
}

}  // namespace rx
