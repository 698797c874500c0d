// Shared helpers for libvssr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vssr_b200.h"

extern long long g_vssr_launches;        // kernels launched (graph replays count their kernel nodes); defined in api.cu
extern long long g_vssr_graph_launches;  // cudaGraphLaunch calls

// Kernel classes for the optional per-class CUDA-event profile (bench.py roofline).
enum VssrKernelClass {
  VSSR_K_NBR = 0, VSSR_K_GEOM, VSSR_K_GEMM, VSSR_K_MSG_FWD, VSSR_K_MSG_BWD, VSSR_K_ELEMWISE, VSSR_K_READOUT,
  VSSR_K_ENSEMBLE, VSSR_K_FIRE, VSSR_K_CLASSICAL, VSSR_K_MSG_FWD_MEMO, VSSR_K_MSG_BWD_MEMO, VSSR_K_NCLASS
};
void vssr_prof_begin(int cls, cudaStream_t st);  // no-ops unless vssr_profile_enable(1)
void vssr_prof_end(int cls, cudaStream_t st);
bool vssr_prof_active();

#define VSSR_LAUNCH_CHECK()                                  \
  do {                                                       \
    ++g_vssr_launches;                                       \
    cudaError_t _e = cudaGetLastError();                     \
    if (_e != cudaSuccess) return (int)_e;                   \
  } while (0)

// launch wrapper: VSSR_PROF(cls, stream, kernel<<<...>>>(...));
#define VSSR_PROF(cls, st, ...)                              \
  do {                                                       \
    vssr_prof_begin((cls), (st));                            \
    __VA_ARGS__;                                             \
    vssr_prof_end((cls), (st));                              \
    VSSR_LAUNCH_CHECK();                                     \
  } while (0)

#define VSSR_CUDA(call)                                      \
  do {                                                       \
    cudaError_t _e = (call);                                 \
    if (_e != cudaSuccess) return (int)_e;                   \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// binary search: structure that owns atom a (atom_ptr ascending, atom_ptr[0] == 0)
__device__ __forceinline__ int struct_of_atom(const int32_t* __restrict__ atom_ptr, int n_struct, int a) {
  int lo = 0, hi = n_struct;  // invariant: atom_ptr[lo] <= a < atom_ptr[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(atom_ptr + mid) <= a) lo = mid; else hi = mid;
  }
  return lo;
}
