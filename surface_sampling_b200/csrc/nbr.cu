// Periodic directed neighbour list, receiver-sorted CSR, bit-exact membership test.
// Replaces NFF AtomsBatch.update_nbr_list (reference call site mcmc/dynamics.py:129; flags
// mcmc/utils/misc.py:34-42).  Semantics and the exact fp32 arithmetic are specified in
// include/vssr_b200.h and restated for the CPU in oracle/nbrlist.py.
//
// Mapping: one warp per receiver atom i; lane l handles sender j = j0 + l and walks that pair's
// own image range S (derived from fractional coordinates, so unwrapped positions are fine) in
// lexicographic order.  Rows come out sorted by (j, S0, S1, S2) with a warp prefix sum — no
// atomics, no sort, deterministic.  At the path's sizes (N <= ~130 atoms per structure, cell edge
// comparable to the cutoff) a spatial cell grid degenerates to one cell per axis, so the periodic
// images themselves are the "cells" that get enumerated.
#include "common.cuh"

namespace {

struct CellInfo {
  float c[9];     // rows a, b, c (fp32, as given)
  double inv[9];  // inverse (fp64): frac = x . inv
  double rch[3];  // cutoff / perpendicular height
  bool pbc[3];
};

__device__ __forceinline__ void load_cell(const float* __restrict__ cell, const uint8_t* __restrict__ pbc,
                                          int b, float rc, CellInfo& ci) {
  double m[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    ci.c[k] = __ldg(cell + 9 * b + k);
    m[k] = (double)ci.c[k];
  }
  // cofactors: cross products of rows
  double c0x = m[4] * m[8] - m[5] * m[7], c0y = m[5] * m[6] - m[3] * m[8], c0z = m[3] * m[7] - m[4] * m[6];  // b x c
  double c1x = m[7] * m[2] - m[8] * m[1], c1y = m[8] * m[0] - m[6] * m[2], c1z = m[6] * m[1] - m[7] * m[0];  // c x a
  double c2x = m[1] * m[5] - m[2] * m[4], c2y = m[2] * m[3] - m[0] * m[5], c2z = m[0] * m[4] - m[1] * m[3];  // a x b
  double det = m[0] * c0x + m[1] * c0y + m[2] * c0z;
  double idet = 1.0 / det;
  // inverse of row-vector matrix M: inv[:,k] = (cross_k)/det  -> frac_k = x . cross_k / det
  ci.inv[0] = c0x * idet; ci.inv[3] = c0y * idet; ci.inv[6] = c0z * idet;
  ci.inv[1] = c1x * idet; ci.inv[4] = c1y * idet; ci.inv[7] = c1z * idet;
  ci.inv[2] = c2x * idet; ci.inv[5] = c2y * idet; ci.inv[8] = c2z * idet;
  double vol = fabs(det);
  double n0 = sqrt(c0x * c0x + c0y * c0y + c0z * c0z);
  double n1 = sqrt(c1x * c1x + c1y * c1y + c1z * c1z);
  double n2 = sqrt(c2x * c2x + c2y * c2y + c2z * c2z);
  ci.rch[0] = (double)rc * n0 / vol;
  ci.rch[1] = (double)rc * n1 / vol;
  ci.rch[2] = (double)rc * n2 / vol;
#pragma unroll
  for (int k = 0; k < 3; ++k) ci.pbc[k] = __ldg(pbc + 3 * b + k) != 0;
}

// Visit every image of pair (i,j) that passes the fp32 test, in lexicographic S order.
template <typename F>
__device__ __forceinline__ int visit_pair(const CellInfo& ci, float xi, float yi, float zi, float xj, float yj,
                                          float zj, float rc2, F&& emit) {
  // fractional difference (fp64) bounds the image range; +-1e-3 slack keeps it a superset
  double dx = (double)xj - (double)xi, dy = (double)yj - (double)yi, dz = (double)zj - (double)zi;
  int lo[3], hi[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double df = dx * ci.inv[k] + dy * ci.inv[3 + k] + dz * ci.inv[6 + k];
    if (ci.pbc[k]) {
      lo[k] = (int)ceil(-df - ci.rch[k] - 1e-3);
      hi[k] = (int)floor(-df + ci.rch[k] + 1e-3);
    } else {
      lo[k] = 0; hi[k] = 0;
    }
  }
  const float fx = __fsub_rn(xj, xi), fy = __fsub_rn(yj, yi), fz = __fsub_rn(zj, zi);
  int n = 0;
  for (int s0 = lo[0]; s0 <= hi[0]; ++s0)
    for (int s1 = lo[1]; s1 <= hi[1]; ++s1)
      for (int s2 = lo[2]; s2 <= hi[2]; ++s2) {
        const float f0 = (float)s0, f1 = (float)s1, f2 = (float)s2;
        float ox = __fadd_rn(__fadd_rn(__fmul_rn(f0, ci.c[0]), __fmul_rn(f1, ci.c[3])), __fmul_rn(f2, ci.c[6]));
        float oy = __fadd_rn(__fadd_rn(__fmul_rn(f0, ci.c[1]), __fmul_rn(f1, ci.c[4])), __fmul_rn(f2, ci.c[7]));
        float oz = __fadd_rn(__fadd_rn(__fmul_rn(f0, ci.c[2]), __fmul_rn(f1, ci.c[5])), __fmul_rn(f2, ci.c[8]));
        float rx = __fadd_rn(fx, ox), ry = __fadd_rn(fy, oy), rz = __fadd_rn(fz, oz);
        float d2 = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
        if (d2 < rc2 && d2 != 0.0f) {
          emit(n, s0, s1, s2);
          ++n;
        }
      }
  return n;
}

template <bool FILL>
__global__ void __launch_bounds__(128) nbr_kernel(const float* __restrict__ pos, const int32_t* __restrict__ atom_ptr,
                                                  const float* __restrict__ cell, const uint8_t* __restrict__ pbc,
                                                  int n_struct, int n_atoms, float rc, int32_t* __restrict__ deg,
                                                  const int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                                  int8_t* __restrict__ shift, long long e_cap,
                                                  int32_t* __restrict__ status) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_atoms) return;
  const int i = warp;
  const int b = struct_of_atom(atom_ptr, n_struct, i);
  const int a0 = __ldg(atom_ptr + b), a1 = __ldg(atom_ptr + b + 1);
  CellInfo ci;
  load_cell(cell, pbc, b, rc, ci);
  const float rc2 = __fmul_rn(rc, rc);
  const float xi = __ldg(pos + 3 * i), yi = __ldg(pos + 3 * i + 1), zi = __ldg(pos + 3 * i + 2);
  long long base = FILL ? (long long)__ldg(rowptr + i) : 0;
  int total = 0;
  for (int j0 = a0; j0 < a1; j0 += 32) {
    const int j = j0 + lane;
    int cnt = 0;
    float xj = 0.f, yj = 0.f, zj = 0.f;
    if (j < a1) {
      xj = __ldg(pos + 3 * j); yj = __ldg(pos + 3 * j + 1); zj = __ldg(pos + 3 * j + 2);
      cnt = visit_pair(ci, xi, yi, zi, xj, yj, zj, rc2, [](int, int, int, int) {});
    }
    // inclusive warp scan of cnt
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int chunk_total = __shfl_sync(0xffffffffu, incl, 31);
    if (FILL && cnt > 0) {
      const long long start = base + (incl - cnt);
      visit_pair(ci, xi, yi, zi, xj, yj, zj, rc2, [&](int n, int s0, int s1, int s2) {
        const long long e = start + n;
        if (e < e_cap) {
          col[e] = j;
          reinterpret_cast<char4*>(shift)[e] = make_char4((signed char)s0, (signed char)s1, (signed char)s2, 0);
        }
      });
    }
    base += chunk_total;
    total += chunk_total;
  }
  if (!FILL && lane == 0) deg[i] = total;
  if (FILL && lane == 0 && base > e_cap) atomicOr(status, VSSR_STATUS_EDGE_OVERFLOW);
}

// exclusive scan of deg[0..n) into rowptr[0..n], rowptr[n] = total.  Single CTA, chunked.
__global__ void __launch_bounds__(1024) scan_kernel(const int32_t* __restrict__ deg, int n, int32_t* __restrict__ rowptr) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int idx = base + tid;
    const int v = idx < n ? deg[idx] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = warp_tot[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;  // exclusive warp offsets
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = carry + warp_tot[wid] + incl - v;
    if (idx < n) rowptr[idx] = excl;
    __syncthreads();
    if (tid == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (tid == 0) rowptr[n] = carry_s;
}

}  // namespace

extern "C" int vssr_nbr_build(const float* pos, const int32_t* atom_ptr, const float* cell, const uint8_t* pbc,
                              int32_t n_struct, int32_t n_atoms, float cutoff, int32_t* deg, int32_t* rowptr,
                              int32_t* col, int8_t* shift, int64_t e_cap, int32_t* status, void* stream) {
  if (!pos || !atom_ptr || !cell || !pbc || !deg || !rowptr || !col || !shift || !status) return VSSR_ERR_ARG;
  if (n_struct < 0 || n_atoms < 0 || e_cap < 0) return VSSR_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_atoms == 0) {
    VSSR_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int32_t), st));
    return VSSR_OK;
  }
  const int warps_per_block = 4;
  const int grid = ceil_div(n_atoms, warps_per_block);
  VSSR_PROF(VSSR_K_NBR, st, nbr_kernel<false><<<grid, 128, 0, st>>>(pos, atom_ptr, cell, pbc, n_struct, n_atoms, cutoff, deg, nullptr, nullptr,
                                          nullptr, 0, status));
  VSSR_PROF(VSSR_K_NBR, st, scan_kernel<<<1, 1024, 0, st>>>(deg, n_atoms, rowptr));
  VSSR_PROF(VSSR_K_NBR, st, nbr_kernel<true><<<grid, 128, 0, st>>>(pos, atom_ptr, cell, pbc, n_struct, n_atoms, cutoff, deg, rowptr, col, shift,
                                         (long long)e_cap, status));
  return VSSR_OK;
}
