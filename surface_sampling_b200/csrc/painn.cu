// PaiNN ensemble: forward + hand-written backward (forces) for a batch of structures x models.
// Replaces NFF Painn.forward + autograd energy_grad (reference call site
// mcmc/calculators/calculators.py:484 via EnsembleNFF.calculate); algorithm restated in
// oracle/painn.py, dataflow prototyped and checked against autograd in oracle/painn_manual.py.
//
// Dataflow (per conv layer, all models x all atoms at once; activations stay in HBM/L2):
//   F1 h1  = s_in W1^T + b1                      GEMM  [A,128]x[128,128]
//   F2 phi = swish(h1) W2^T + b2                 GEMM  [A,128]x[128,384]   (swish on the A load)
//   F3 message: thread f of receiver i walks i's CSR row serially (deterministic), builds the
//      filter w(d) = Wd.(rbf*env) + bd*env in registers, gathers phi_j / v_j      -> s_mid, v_mid
//   F4 [Uv|Vv] = v_mid [U^T|V^T]                 GEMM  [3A,128]x[128,256]
//   F5 nrm = ||Vv||  -> cat = [s_mid | nrm]
//   F6 h3 = cat W3^T + b3 ; F7 a = swish(h3) W4^T + b4
//   F8 s_out = s_mid + <Uv,Vv> a_sv + a_ss ; v_out = v_mid + Uv a_vv
// Backward mirrors it; the message backward is a GATHER over the receiver's own row in which
// edge i<-j is handled together with its reverse j<-i (same d, unit negated), so sender-side
// gradients and the position gradient of atom i are produced by atom i's CTA: no atomics.
//
// Vector features are stored [A,3,128] (Cartesian-major) so U/V act as plain row GEMMs.
#include <cuda.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <initializer_list>
#include <type_traits>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

#include "common.cuh"
#include "painn_layout.h"

namespace {
using namespace painn;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float swishf_(float x) { return x * sigmoidf_(x); }
__device__ __forceinline__ float dswishf_(float x) {
  float s = sigmoidf_(x);
  return s * (1.0f + x * (1.0f - s));
}

// ------------------------------------------------------------------------------------------
// SGEMM (fp32 FMA):  C[M,N] (op)= T(A)[M,K] . B[K,N],  batched over models on blockIdx.z.
//   AMODE 0: A as is; 1: swish(A); 2: dswish(A) * avec[k]
//   EPI   0: C = acc; 1: C = acc + bias[n]; 2: C = acc * dswish(aux[m,n]); 3: C += acc;
//         4: C = acc + bias[n] and C2 = swish(C)  (activation for the next GEMM, computed once)
// Tile 128 x BN x 16, 256 threads, 8 x (BN/16) outputs per thread, register-prefetch pipeline.
// ------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A; int lda; long long sA;
  const float* B; int ldb; long long sB;
  const float* bias; long long sBias;
  const float* aux; int ldaux; long long sAux;
  const float* avec; long long sAvec;
  float* C; int ldc; long long sC;
  int M, N, K;
  float* C2;   // EPI 4 only: second output swish(C) with C's leading dimension and model stride
};

#include "gemm_tc.cuh"

// GEMM dispatcher: every node MLP runs on the tcgen05 kernels (csrc/gemm_tc.cuh), 3xTF32.  Bnk = the K-major [N][K]
// copy of the weight inside the exact-fp32 block of the packed weights; its TF32 hi / lo parts live one and two W_TOTAL
// further (painn_layout.h).
template <int BN, int AMODE, int EPI>
int run_gemm(GemmArgs g, const float* Bnk, int n_models, cudaStream_t st) {
  // wave quantisation: a 128-column GEMM on ~8000 rows is 195 tiles = 1.3 waves of the 148 persistent CTAs, i.e.
  // two rounds; with 64-column tiles it is 390 half-size tiles = three half rounds
  if (BN == 128 && AMODE == 0 && g.N == 128) {
    const long long tiles = (long long)ceil_div(g.M, 128) * n_models;
    if (tiles < 2 * 148)
      return launch_gemm_tc<64, AMODE, EPI>(g, Bnk + W_TOTAL, Bnk + 2 * W_TOTAL, W_STRIDE, n_models, st);
  }
  return launch_gemm_tc<BN, AMODE, EPI>(g, Bnk + W_TOTAL, Bnk + 2 * W_TOTAL, W_STRIDE, n_models, st);
}

// ------------------------------------------------------------------------------------------
// Edge geometry (shared by all models and layers) + excluded volume.
// One warp per receiver atom; lanes over the row in chunks of 32.  Edges inside the model cutoff
// are COMPACTED to the front of the row (order preserved: ballot + popc), nvalid[i] of them:
// one 384-byte record per edge, erec[e][96 floats]:
//   [0..3]   (ux,uy,uz,d)   [4] sender (global atom index)  [5] -1  [6] (j_local,S) key  [7] 1/d
//   [8..51]  (rbf_n*env) duplicated as pairs (v,v) for n=1..20, then (env,env),(denv,denv)
//   [52..91] d(rbf_n*env)/dd duplicated as pairs                              [92..95] pad
// (a record is what one warp prefetches with a single cp.async per lane).  Edges whose filter is memoised
// (see FilterCacheView) go to a second compact list of 32-byte records mrec[e][8], nmemo[i] per row.
// The pair duplication feeds the packed FFMA2 path (sm_100 fma.rn.f32x2) without register moves.
// grad0[a] = excluded-volume gradient (same for every model); evex[a] its energy.
// ------------------------------------------------------------------------------------------
constexpr int REC = 96, REC_EJ = 4, REC_RE = 8, REC_DRE = 52;
constexpr int MREC = 8;   // memoised-edge record: (ux,uy,uz,d), sender, slot, 1/d, pad
// compact direct-edge record for the k-block kernels (256 B): [0..3] (u,d) [4] sender [6] key [7] 1/d
// [8..27] rbf_n*env [28] env [29] denv [32..51] d(rbf_n*env)/dd   (scalars: FFMA2 broadcasts an .F32 operand)

// Radial-filter memo ("frozen-pair cache").  The filter w(d) = Wd.(rbf(d)*env(d)) + bd*env(d) and its
// derivative q(d) depend on the edge only through the scalar d.  In VSSR-MC every chain shares the same
// frozen bulk framework, so ~3/4 of all edges have, in every chain and at every FIRE step, bit-identical
// d.  Their w and q (per model and layer, 384 floats each) are computed once (vssr_painn_filter_cache_build)
// and looked up by exact distance match: an edge (i_local, j_local, S) of a chain is resolved to the
// framework edge with the same key by binary search in the framework's CSR row and uses the memo only if
// its fp32 d is bitwise equal to the framework's.  No promise from the caller is needed.
struct FilterCacheView {
  int n0;                    // framework atoms (0 = no cache)
  int nslots_cap;            // slot capacity (row stride of wc/qc per model-layer)
  const int32_t* rowptr;     // [n0+1] framework CSR (cutoff+skin list); row i holds nvalid[i] compacted edges
  const int32_t* nvalid;     // [n0]
  const int32_t* key;        // [E0] j_local*125 + (S0+2)*25 + (S1+2)*5 + (S2+2), ascending within a row
  const int32_t* slot;       // [E0] memo slot or -1
  const float* d;            // [E0] fp32 distance of the framework edge
  const float* wc;           // [M*3][nslots_cap][384]
  const float* qc;           // [M*3][nslots_cap][384]
  const uint8_t* frozen;     // [n0] framework atoms flagged frozen (FixAtoms)
  // the framework's own memoised-edge lists ("canonical" lists): a structure whose rows all carry exactly
  // these edges can be processed from them, two structures per CTA, with each filter row loaded once
  const int32_t* nmemo0;     // [n0]
  const float* mrec0;        // [E0][MREC], row i at rowptr[i]
  const int32_t* order0;     // [n0] rows by nmemo0, descending
};

__global__ void __launch_bounds__(128) edge_geometry_kernel(
    const float* __restrict__ pos, const int32_t* __restrict__ atom_ptr, const float* __restrict__ cell, int n_struct,
    int n_atoms, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const int8_t* __restrict__ shift, long long e_cap, float cutoff, FilterCacheView fc,
    int32_t* __restrict__ nvalid, float* __restrict__ erec, int32_t* __restrict__ nmemo, float* __restrict__ mrec,
    float* __restrict__ evex, float* __restrict__ grad0) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n_atoms) return;
  const int b = struct_of_atom(atom_ptr, n_struct, i);
  const int a0 = __ldg(atom_ptr + b);
  float c[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) c[k] = __ldg(cell + 9 * b + k);
  const float xi = __ldg(pos + 3 * i), yi = __ldg(pos + 3 * i + 1), zi = __ldg(pos + 3 * i + 2);
  long long e0 = __ldg(rowptr + i), e1 = __ldg(rowptr + i + 1);
  if (e1 > e_cap) e1 = e_cap;
  if (e0 > e_cap) e0 = e_cap;
  // framework row of this atom (if it is one of the first n0 atoms of its structure)
  int f0 = 0, f1 = 0;
  if (i - a0 < fc.n0) { f0 = __ldg(fc.rowptr + (i - a0)); f1 = f0 + __ldg(fc.nvalid + (i - a0)); }
  const float pi_over_rc = 3.14159265358979323846f / cutoff;
  float ev = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
  int nv = 0, nm = 0;   // direct / memoised edges of this row
  for (long long base = e0; base < e1; base += 32) {
    const long long e = base + lane;
    bool valid = false;
    int j = 0, key = -1;
    float rx = 0.f, ry = 0.f, rz = 0.f, dp = 0.f;
    if (e < e1) {
      j = __ldg(col + e);
      const char4 s = reinterpret_cast<const char4*>(shift)[e];
      if (s.x >= -2 && s.x <= 2 && s.y >= -2 && s.y <= 2 && s.z >= -2 && s.z <= 2)
        key = (j - a0) * 125 + (s.x + 2) * 25 + (s.y + 2) * 5 + (s.z + 2);
      const float f0 = (float)s.x, f1 = (float)s.y, f2 = (float)s.z;
      // same fp32 offset arithmetic as the neighbour list
      const float ox = __fadd_rn(__fadd_rn(__fmul_rn(f0, c[0]), __fmul_rn(f1, c[3])), __fmul_rn(f2, c[6]));
      const float oy = __fadd_rn(__fadd_rn(__fmul_rn(f0, c[1]), __fmul_rn(f1, c[4])), __fmul_rn(f2, c[7]));
      const float oz = __fadd_rn(__fadd_rn(__fmul_rn(f0, c[2]), __fmul_rn(f1, c[5])), __fmul_rn(f2, c[8]));
      rx = __fadd_rn(__fsub_rn(__ldg(pos + 3 * j), xi), ox);
      ry = __fadd_rn(__fsub_rn(__ldg(pos + 3 * j + 1), yi), oy);
      rz = __fadd_rn(__fsub_rn(__ldg(pos + 3 * j + 2), zi), oz);
      const float d2p = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
      dp = sqrtf(d2p);
      valid = dp <= cutoff;
    }
    // filter memo lookup: same (j_local, S) in the framework row AND bit-identical distance
    float d = 0.f;
    int slot = -1;
    if (valid) {
      d = sqrtf((rx * rx + 1e-10f) + (ry * ry + 1e-10f) + (rz * rz + 1e-10f));
      if (key >= 0 && j - a0 < fc.n0) {
        int lo = f0, hi = f1;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(fc.key + mid) < key) lo = mid + 1; else hi = mid;
        }
        if (lo < f1 && __ldg(fc.key + lo) == key && __float_as_int(__ldg(fc.d + lo)) == __float_as_int(d))
          slot = __ldg(fc.slot + lo);
      }
    }
    const bool is_memo = valid && slot >= 0, is_direct = valid && slot < 0;
    const unsigned mask_m = __ballot_sync(0xffffffffu, is_memo), mask_d = __ballot_sync(0xffffffffu, is_direct);
    if (valid) {
      const float inv_d = 1.0f / d;
      const float ux = rx * inv_d, uy = ry * inv_d, uz = rz * inv_d;
      if (is_memo) {
        // memoised edges: 32-byte record (unit, d, sender, slot), own compact list
        const long long w = e0 + nm + __popc(mask_m & ((1u << lane) - 1u));
        float4* mr = reinterpret_cast<float4*>(mrec + w * MREC);
        mr[0] = make_float4(ux, uy, uz, d);
        mr[1] = make_float4(__int_as_float(j), __int_as_float(slot), inv_d, 0.f);
      } else {
        const long long w = e0 + nv + __popc(mask_d & ((1u << lane) - 1u));
        float* rec = erec + w * REC;
        *reinterpret_cast<float4*>(rec) = make_float4(ux, uy, uz, d);
        *reinterpret_cast<float4*>(rec + 4) = make_float4(__int_as_float(j), __int_as_float(-1), __int_as_float(key), inv_d);
        float env = 0.f, denv = 0.f;
        const bool inside = d < cutoff;
        if (inside) {
          float sn, cs;
          sincosf(pi_over_rc * d, &sn, &cs);
          env = 0.5f * (cs + 1.0f);
          denv = -0.5f * pi_over_rc * sn;
        }
        float2* rrow = reinterpret_cast<float2*>(rec + REC_RE);
        float2* drow = reinterpret_cast<float2*>(rec + REC_DRE);
#pragma unroll
        for (int n = 0; n < NRBF; ++n) {
          float r = 0.f, dr = 0.f;
          if (inside) {
            const float coef = (float)(n + 1) * pi_over_rc;
            float sn, cs;
            sincosf(coef * d, &sn, &cs);
            r = sn * inv_d;
            dr = (coef * cs - r) * inv_d;
          }
          const float a = r * env, bq = dr * env + r * denv;
          rrow[n] = make_float2(a, a);
          drow[n] = make_float2(bq, bq);
        }
        rrow[20] = make_float2(env, env);
        rrow[21] = make_float2(denv, denv);
      }
      // excluded volume (sigma/d)^12 on the plain distance
      const float q = 1.5f / dp;
      const float q2 = q * q, q4 = q2 * q2;
      const float vex = q4 * q4 * q4;
      ev += vex;
      const float gg = 24.0f * vex / dp;      // -2 * dvex/dd : edge A and its reverse B
      gx += gg * ux; gy += gg * uy; gz += gg * uz;
    }
    nv += __popc(mask_d);
    nm += __popc(mask_m);
  }
  ev = warp_sum(ev); gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
  if (lane == 0) {
    nvalid[i] = nv;
    nmemo[i] = nm;
    evex[i] = ev;
    grad0[3 * i] = gx; grad0[3 * i + 1] = gy; grad0[3 * i + 2] = gz;
  }
}

// s0[m][a][:] = embed_m[z[a]][:]
__global__ void embed_kernel(const float* __restrict__ weights, const int32_t* __restrict__ z, int n_atoms,
                             float* __restrict__ s0) {
  const int m = blockIdx.y;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // float4 index
  if (idx >= (long long)n_atoms * (F / 4)) return;
  const int a = (int)(idx / (F / 4)), f4 = (int)(idx % (F / 4));
  int zz = __ldg(z + a);
  zz = zz < 0 ? 0 : (zz >= NEMB ? NEMB - 1 : zz);
  const float4* emb = reinterpret_cast<const float4*>(weights + (long long)m * W_STRIDE + W_EMBED);
  reinterpret_cast<float4*>(s0 + (long long)m * n_atoms * F)[idx] = emb[zz * (F / 4) + f4];
}

// F5: cat[a][128+f] = sqrt(sum_c (Vv[c][f]^2 + 1e-15))
__global__ void nrm_kernel(const float* __restrict__ UV, int n_atoms, float* __restrict__ cat) {
  const int m = blockIdx.y;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_atoms * F) return;
  const long long a = idx / F;
  const int f = (int)(idx % F);
  const float* uv = UV + ((long long)m * n_atoms + a) * 3 * 2 * F;
  const float x = uv[F + f], y = uv[2 * F + F + f], zz = uv[4 * F + F + f];
  cat[((long long)m * n_atoms + a) * 2 * F + F + f] = sqrtf((x * x + 1e-15f) + (y * y + 1e-15f) + (zz * zz + 1e-15f));
}

// F8: s_out = s_mid + <Uv,Vv> a_sv + a_ss ; v_out = v_mid + Uv a_vv
__global__ void update_fwd_kernel(const float* __restrict__ UV, const float* __restrict__ a, const float* __restrict__ cat,
                                  const float* __restrict__ v_mid, int n_atoms, float* __restrict__ s_out,
                                  float* __restrict__ v_out) {
  const int m = blockIdx.y;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_atoms * F) return;
  const long long am = (long long)m * n_atoms + idx / F;
  const int f = (int)(idx % F);
  const float* uv = UV + am * 6 * F;
  const float* aa = a + am * F3;
  const float ux = uv[f], uy = uv[2 * F + f], uz = uv[4 * F + f];
  const float vx = uv[F + f], vy = uv[3 * F + f], vz = uv[5 * F + f];
  const float inner = ux * vx + uy * vy + uz * vz;
  const float avv = aa[f], asv = aa[F + f], ass = aa[2 * F + f];
  s_out[am * F + f] = cat[am * 2 * F + f] + inner * asv + ass;
  const float* vm = v_mid + am * 3 * F;
  float* vo = v_out + am * 3 * F;
  vo[f] = vm[f] + ux * avv;
  vo[F + f] = vm[F + f] + uy * avv;
  vo[2 * F + f] = vm[2 * F + f] + uz * avv;
}

// readout energy: e_atom[m][a] = swish(h5).w6 + b6 + evex[a]; one warp per (m, a)
__global__ void __launch_bounds__(128) readout_energy_kernel(const float* __restrict__ weights, const float* __restrict__ h5,
                                                             const float* __restrict__ evex, int n_atoms,
                                                             float* __restrict__ e_atom) {
  const int m = blockIdx.y;
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (a >= n_atoms) return;
  const float* w = weights + (long long)m * W_STRIDE;
  const float* h = h5 + ((long long)m * n_atoms + a) * FH;
  float acc = swishf_(h[lane]) * __ldg(w + R_W6 + lane) + swishf_(h[lane + 32]) * __ldg(w + R_W6 + lane + 32);
  acc = warp_sum(acc);
  if (lane == 0) e_atom[(long long)m * n_atoms + a] = acc + __ldg(w + R_B6) + evex[a];
}

// per-structure energy sum in fp64, one warp per (m, structure), fixed order
__global__ void __launch_bounds__(128) energy_reduce_kernel(const float* __restrict__ e_atom, const int32_t* __restrict__ atom_ptr,
                                                            int n_struct, int n_atoms, double* __restrict__ energy) {
  const int m = blockIdx.y;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= n_struct) return;
  const int a0 = atom_ptr[b], a1 = atom_ptr[b + 1];
  double acc = 0.0;
  for (int a = a0 + lane; a < a1; a += 32) acc += (double)e_atom[(long long)m * n_atoms + a];
  acc = warp_sum(acc);
  if (lane == 0) energy[(long long)m * n_struct + b] = acc;
}

// grad[m][a][c] = grad0[a][c]  (excluded-volume part, identical for all models)
__global__ void grad_init_kernel(const float* __restrict__ grad0, int n3, float* __restrict__ grad) {
  const int m = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n3) grad[(long long)m * n3 + idx] = grad0[idx];
}

// B8: update elementwise backward.  in: ds, dv, UV, a.  out: da[A,384], dUV[A,3,256]
__global__ void update_bwd_kernel(const float* __restrict__ ds, const float* __restrict__ dv, const float* __restrict__ UV,
                                  const float* __restrict__ a, int n_atoms, float* __restrict__ da,
                                  float* __restrict__ dUV) {
  const int m = blockIdx.y;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_atoms * F) return;
  const long long am = (long long)m * n_atoms + idx / F;
  const int f = (int)(idx % F);
  const float* uv = UV + am * 6 * F;
  const float* aa = a + am * F3;
  const float ux = uv[f], uy = uv[2 * F + f], uz = uv[4 * F + f];
  const float vx = uv[F + f], vy = uv[3 * F + f], vz = uv[5 * F + f];
  const float gs = ds[am * F + f];
  const float* gv = dv + am * 3 * F;
  const float gvx = gv[f], gvy = gv[F + f], gvz = gv[2 * F + f];
  const float avv = aa[f], asv = aa[F + f];
  const float inner = ux * vx + uy * vy + uz * vz;
  float* d = da + am * F3;
  d[f] = gvx * ux + gvy * uy + gvz * uz;
  d[F + f] = gs * inner;
  d[2 * F + f] = gs;
  const float t = gs * asv;
  float* du = dUV + am * 6 * F;
  du[f] = gvx * avv + t * vx;       du[F + f] = t * ux;
  du[2 * F + f] = gvy * avv + t * vy; du[3 * F + f] = t * uy;
  du[4 * F + f] = gvz * avv + t * vz; du[5 * F + f] = t * uz;
}

// B5: ds += dcat[:128] ; dVv[c][f] += dcat[128+f]/nrm[f] * Vv[c][f]
__global__ void nrm_bwd_kernel(const float* __restrict__ dcat, const float* __restrict__ cat, const float* __restrict__ UV,
                               int n_atoms, float* __restrict__ ds, float* __restrict__ dUV) {
  const int m = blockIdx.y;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_atoms * F) return;
  const long long am = (long long)m * n_atoms + idx / F;
  const int f = (int)(idx % F);
  ds[am * F + f] += dcat[am * 2 * F + f];
  const float t = dcat[am * 2 * F + F + f] / cat[am * 2 * F + F + f];
  const float* uv = UV + am * 6 * F;
  float* du = dUV + am * 6 * F;
  du[F + f] += t * uv[F + f];
  du[3 * F + f] += t * uv[3 * F + f];
  du[5 * F + f] += t * uv[5 * F + f];
}

#include "painn_message.cuh"
#include "painn_message_tc.cuh"

// ---- filter memo construction (one-time per framework + weights) ----
struct CacheBlob {
  int32_t* rowptr; int32_t* nvalid; int32_t* key; int32_t* slot; float* d; float* wc; float* qc; int32_t* counter;
  uint8_t* frozen;   // [n0] copy of fixed0
  int32_t* nmemo0; int32_t* order0; float* mrec0;
  size_t bytes;
};
CacheBlob carve_cache(void* base, int M, int n0, long long e_cap0) {
  CacheBlob c;
  size_t off = 0;
  auto take = [&](size_t nbytes) -> void* {
    void* p = base ? reinterpret_cast<char*>(base) + off : nullptr;
    off += ((nbytes + 255) / 256) * 256;
    return p;
  };
  c.rowptr = (int32_t*)take((size_t)(n0 + 1) * 4);
  c.nvalid = (int32_t*)take((size_t)n0 * 4);
  c.key = (int32_t*)take((size_t)e_cap0 * 4);
  c.slot = (int32_t*)take((size_t)e_cap0 * 4);
  c.d = (float*)take((size_t)e_cap0 * 4);
  c.counter = (int32_t*)take(4);
  c.frozen = (uint8_t*)take((size_t)n0);
  c.nmemo0 = (int32_t*)take((size_t)n0 * 4);
  c.order0 = (int32_t*)take((size_t)n0 * 4);
  c.mrec0 = (float*)take((size_t)e_cap0 * MREC * 4);
  c.wc = (float*)take((size_t)M * NCONV * e_cap0 * F3 * 4);
  c.qc = (float*)take((size_t)M * NCONV * e_cap0 * F3 * 4);
  c.bytes = off;
  return c;
}
FilterCacheView cache_view(const void* blob, int M, int n0, long long e_cap0) {
  FilterCacheView v{};
  if (!blob || n0 <= 0) return v;
  CacheBlob c = carve_cache(const_cast<void*>(blob), M, n0, e_cap0);
  v.n0 = n0; v.nslots_cap = (int)e_cap0; v.rowptr = c.rowptr; v.nvalid = c.nvalid; v.key = c.key; v.slot = c.slot;
  v.d = c.d; v.wc = c.wc; v.qc = c.qc; v.frozen = c.frozen;
  v.nmemo0 = c.nmemo0; v.mrec0 = c.mrec0; v.order0 = c.order0;
  return v;
}

// one thread: number the framework edges whose two atoms are both frozen (their d never changes)
__global__ void cache_slot_kernel(const float* __restrict__ erec, const int32_t* __restrict__ rowptr,
                                  const int32_t* __restrict__ nvalid, const uint8_t* __restrict__ fixed0, int n0,
                                  int32_t* __restrict__ key, int32_t* __restrict__ slot, float* __restrict__ d,
                                  int32_t* __restrict__ counter) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int n = 0;
  for (int i = 0; i < n0; ++i) {
    const int e0 = rowptr[i];
    for (int w = e0; w < e0 + nvalid[i]; ++w) {
      const float* rec = erec + (long long)w * REC;
      const int j = __float_as_int(rec[REC_EJ]);
      key[w] = __float_as_int(rec[6]);
      d[w] = rec[3];
      slot[w] = (fixed0[i] && fixed0[j] && __float_as_int(rec[6]) >= 0) ? n++ : -1;
    }
  }
  *counter = n;
}

// grid (e_cap0, M*NCONV), 128 threads (feature f): w_k, q_k of memoised edges, same FMA order as the
// message kernels' direct evaluation
__global__ void __launch_bounds__(128) cache_fill_kernel(const float* __restrict__ weights, const float* __restrict__ erec,
                                                         const int32_t* __restrict__ slot, int e_cap0,
                                                         float* __restrict__ wc, float* __restrict__ qc) {
  const int w = blockIdx.x, ml = blockIdx.y, f = threadIdx.x;
  const int sl = slot[w];
  if (sl < 0) return;
  const int m = ml / NCONV, layer = ml % NCONV;
  const float* __restrict__ wl = weights + (long long)m * W_STRIDE + W_LAYER0 + (long long)layer * L_SIZE;
  const float* rec = erec + (long long)w * REC;
  const float env = rec[REC_RE + 40], denv = rec[REC_RE + 42];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float bd = wl[L_BD + k * F + f];
    float wv = bd * env, qv = bd * denv;
    for (int n = 0; n < NRBF; ++n) {
      const float wd = wl[L_WDT + n * F3 + k * F + f];
      wv = fmaf(wd, rec[REC_RE + 2 * n], wv);
      qv = fmaf(wd, rec[REC_DRE + 2 * n], qv);
    }
    wc[((long long)ml * e_cap0 + sl) * F3 + k * F + f] = wv;
    qc[((long long)ml * e_cap0 + sl) * F3 + k * F + f] = qv;
  }
}


struct Workspace {
  // edge records (compacted per row by edge_geometry_kernel)
  int32_t* nvalid; float* erec; int32_t* nmemo; float* mrec; float* evex; float* grad0; float* gradp;
  int32_t* order_d; int32_t* order_m;   // per-structure row order, most direct / memoised edges first
  int32_t* canonical;                   // [A] (first n_struct used): structure carries exactly the framework's memo lists
  // activations
  float* s[NCONV + 1];      // [M,A,128]
  float* v[NCONV + 1];      // [M,A,3,128]  (v[0] unused: zeros)
  float* h1[NCONV]; float* phi[NCONV]; float* cat[NCONV]; float* vmid[NCONV]; float* UV[NCONV];
  float* h3[NCONV]; float* a[NCONV];
  float* h5; float* e_atom; float* act;   // act: swish(h1)/swish(h3) scratch [M,A,128]
  // gradients
  float* ds; float* dvA; float* dvB; float* dphi; float* dh1; float* da; float* dh3; float* dcat; float* dUV;
  size_t bytes;
};

Workspace carve(void* base, int M, int A, long long e_cap) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t nfloats) -> float* {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += ((nfloats * sizeof(float) + 255) / 256) * 256;
    return p;
  };
  const size_t MA = (size_t)M * (size_t)A;
  w.nvalid = reinterpret_cast<int32_t*>(take(A));
  w.erec = take((size_t)e_cap * REC);
  w.nmemo = reinterpret_cast<int32_t*>(take(A));
  w.order_d = reinterpret_cast<int32_t*>(take(A));
  w.order_m = reinterpret_cast<int32_t*>(take(A));
  w.canonical = reinterpret_cast<int32_t*>(take(A));
  w.mrec = take((size_t)e_cap * MREC);
  w.evex = take(A);
  w.grad0 = take((size_t)A * 3);
  w.gradp = take(MA * 2 * 3);
  for (int l = 0; l <= NCONV; ++l) w.s[l] = take(MA * F);
  w.v[0] = nullptr;
  for (int l = 1; l <= NCONV; ++l) w.v[l] = take(MA * 3 * F);
  for (int l = 0; l < NCONV; ++l) {
    w.h1[l] = take(MA * F); w.phi[l] = take(MA * F3); w.cat[l] = take(MA * 2 * F); w.vmid[l] = take(MA * 3 * F);
    w.UV[l] = take(MA * 6 * F); w.h3[l] = take(MA * F); w.a[l] = take(MA * F3);
  }
  w.h5 = take(MA * FH); w.e_atom = take(MA); w.act = take(MA * F);
  w.ds = take(MA * F); w.dvA = take(MA * 3 * F); w.dvB = take(MA * 3 * F); w.dphi = take(MA * F3);
  w.dh1 = take(MA * F); w.da = take(MA * F3); w.dh3 = take(MA * F); w.dcat = take(MA * 2 * F);
  w.dUV = take(MA * 6 * F);
  w.bytes = off;
  return w;
}

}  // namespace

extern "C" int64_t vssr_painn_weight_floats(void) { return (int64_t)painn::W_STRIDE; }

extern "C" size_t vssr_painn_workspace_bytes(int32_t n_models, int32_t n_atoms, int64_t e_cap) {
  return carve(nullptr, n_models, n_atoms, e_cap).bytes;
}

extern "C" int vssr_painn_energy_grad(const float* weights, int32_t n_models, const float* pos, const int32_t* z,
                                      const int32_t* atom_ptr, const float* cell, int32_t n_struct, int32_t n_atoms,
                                      int32_t max_atoms_per_struct, const int32_t* rowptr, const int32_t* col,
                                      const int8_t* shift, int64_t e_cap, float cutoff, const void* filter_cache,
                                      int32_t fc_n0, int64_t fc_e_cap0, int32_t fc_flags, void* workspace,
                                      size_t workspace_bytes, double* energy, float* grad, float* embedding,
                                      void* stream) {
  if (!weights || !pos || !z || !atom_ptr || !cell || !rowptr || !col || !shift || !workspace || !energy || !grad)
    return VSSR_ERR_ARG;
  if (n_models <= 0 || n_struct <= 0 || n_atoms <= 0) return VSSR_ERR_ARG;
  Workspace w = carve(workspace, n_models, n_atoms, e_cap);
  if (w.bytes > workspace_bytes) return VSSR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int M = n_models, A = n_atoms;
  const long long MA_F = (long long)A * F;  // per-model stride of [A,128]
  const dim3 ew_grid(ceil_div((long long)A * F, 256), M);

  // message kernels: shared-memory staged FFMA2 kernels.  A CTA of the direct pass stages at most cap_* atoms; larger
  // structures are covered by several launches over "sender windows" (painn_message.cuh: sender_window), so there is
  // no atom-count cliff (round 1's global-gather fallback kernels are gone).
  const int nmax = max_atoms_per_struct;
  if (nmax <= 0) return VSSR_ERR_ARG;   // the caller knows its largest structure (sizes the staging areas)
  const FilterCacheView fc = cache_view(filter_cache, n_models, fc_n0, fc_e_cap0);
  constexpr bool staged = true;
  const size_t kSmemCap = 223 * 1024;   // 227 KB per SM minus the kernels' static shared memory (row table: 3 KB)
  // direct pass: staged rows of one window + the per-warp record rings.  The first-layer forward runs two CTAs per SM.
  constexpr int B_FWD0 = MsgFwdLayout<true>::PER * 4, B_FWD = MsgFwdLayout<false>::PER * 4;
  constexpr int B_BWD0 = MsgBwdLayout<true>::PER * 4, B_BWD = MsgBwdLayout<false>::PER * 4;
  const int cap_fwd0 = (int)((kSmemCap / 2 - 1024 - MSG_PIPE_BYTES_FWD) / B_FWD0), cap_fwd = (int)((kSmemCap - MSG_PIPE_BYTES_FWD) / B_FWD);
  const int cap_bwd0 = (int)((kSmemCap - MSG_PIPE_BYTES) / B_BWD0), cap_bwd = (int)((kSmemCap - MSG_PIPE_BYTES) / B_BWD);
  auto rows_of = [&](int cap) { return nmax < cap ? nmax : cap; };
  auto windows_of = [&](int cap) { return (nmax + cap - 1) / cap; };
  const size_t smem_fwd0 = (size_t)rows_of(cap_fwd0) * B_FWD0 + MSG_PIPE_BYTES_FWD, smem_fwd = (size_t)rows_of(cap_fwd) * B_FWD + MSG_PIPE_BYTES_FWD;
  const size_t smem_bwd0 = (size_t)rows_of(cap_bwd0) * B_BWD0 + MSG_PIPE_BYTES, smem_bwd = (size_t)rows_of(cap_bwd) * B_BWD + MSG_PIPE_BYTES;
  // memo pass: only framework rows (the first n0 atoms) are ever gathered, whatever the structure carries on top
  const int nrows_memo = fc.n0 > 0 ? (nmax < fc.n0 ? nmax : fc.n0) : 0;
  const size_t memo6_ring = (size_t)(MEMO_THREADS_BWD / 32) * MEMO6_STAGES * MEMO6_STAGE_FLOATS * 4;
  const size_t sm_bwdm0 = (size_t)nrows_memo * B_BWD0 + memo6_ring, sm_bwdm = (size_t)nrows_memo * B_BWD + memo6_ring;
  const size_t memo_ring = (size_t)(MEMO_THREADS_FWD / 32) * MEMO_RING_BYTES_PER_WARP;
  const size_t sm_fwd0 = (size_t)nrows_memo * B_FWD0 + memo_ring, sm_fwd = (size_t)nrows_memo * B_FWD + memo_ring;
  const size_t sm_state = (size_t)nrows_memo * MEMO_STATE_PER * 4 + memo_ring;
  const bool want_constrained = (fc_flags & VSSR_FC_CONSTRAINED_GRAD) != 0;
  // two passes: memoised edges (light kernels), then direct edges; a framework too large for the staging area simply
  // runs without the memo
  const bool memo = staged && fc.n0 > 0 && sm_bwdm <= kSmemCap && sm_fwd <= kSmemCap;
  const bool constrained = memo && want_constrained;   // no dE/dx wanted on frozen atoms
  // VSSR_MSG_TC=1: direct FORWARD pass with the radial filter on the tensor core (painn_message_tc.cuh).  Experimental
  // and OFF by default: parity-green but slower than message_fwd_v2 (profiles/round2_notes.md section 4)
  static int tc_env = -1;
  if (tc_env < 0) { const char* e = getenv("VSSR_MSG_TC"); tc_env = (e && atoi(e) > 0) ? 1 : 0; }
  const size_t tc_rows_budget = 227 * 1024 - 6144 /* static shared memory of the kernel */ - 1024 - mtc::FIXED_BYTES;
  const int cap_tc0 = (int)(tc_rows_budget / B_FWD0), cap_tc = (int)(tc_rows_budget / B_FWD);
  const bool tc_fwd = tc_env != 0 && nmax <= 2 * mtc::MAXROWS;
  const size_t smem_tc0 = mtc::fwd_smem_bytes(rows_of(cap_tc0), true), smem_tc = mtc::fwd_smem_bytes(rows_of(cap_tc), false);
  constexpr int n_chunks = 2;   // CTAs per (structure, feature half, model) in the direct pass
  const dim3 v2_grid((F / MSG_FC) * M, n_struct * n_chunks);   // (half, model) fastest: see message_fwd_v2
  const dim3 memo_grid(n_struct, F / MSG_FC, M);
  // group kernels: G canonical structures per CTA (as many as fit in shared memory), no ring
  constexpr int G_FWD0 = 4, T_FWD0 = 512, G_FWD = 2, T_FWD = 832, G_STATE = 2, T_STATE = 512;
  const bool group_on = memo && !(fc_flags & VSSR_FC_NO_PAIR);
  // a group is taken when its structures are canonical AND the FRAMEWORK rows of G structures (n0 atoms each -- the only
  // rows a memoised edge can touch) fit the 227 KB of an SM: independent of how many adsorbates the chains carry
  const size_t sp_fwd0 = (size_t)G_FWD0 * fc.n0 * B_FWD0, sp_fwd = (size_t)G_FWD * fc.n0 * B_FWD,
               sp_state = (size_t)G_STATE * fc.n0 * MEMO_STATE_PER * 4;
  const int ga_fwd0 = sp_fwd0 <= kSmemCap ? G_FWD0 * fc.n0 : 0, ga_fwd = sp_fwd <= kSmemCap ? G_FWD * fc.n0 : 0,
            ga_state = sp_state <= kSmemCap ? G_STATE * fc.n0 : 0;
  const bool pair_fwd0 = group_on && n_struct >= G_FWD0 && ga_fwd0 > 0;
  const bool pair_fwd = group_on && n_struct >= G_FWD && ga_fwd > 0;
  const bool pair_state = group_on && constrained && n_struct >= G_STATE && ga_state > 0;
  if (staged) {
    // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: cache what was configured per device
    static size_t cfg_dev[64][16] = {};
    int dev = 0;
    VSSR_CUDA(cudaGetDevice(&dev));
    size_t* cfg = cfg_dev[dev & 63];
    auto want = [&](int k, const void* fn, size_t bytes) -> int {
      if (bytes > cfg[k]) {
        VSSR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cfg[k] = bytes;
      }
      return VSSR_OK;
    };
    int rc0;
    if ((rc0 = want(0, (const void*)message_fwd_v2<true>, smem_fwd0))) return rc0;
    if ((rc0 = want(1, (const void*)message_fwd_v2<false>, smem_fwd))) return rc0;
    if ((rc0 = want(2, (const void*)message_bwd_v2<true>, smem_bwd0))) return rc0;
    if ((rc0 = want(3, (const void*)message_bwd_v2<false>, smem_bwd))) return rc0;
    if (tc_fwd) {
      if ((rc0 = want(12, (const void*)mtc::message_fwd_tc<true>, smem_tc0))) return rc0;
      if ((rc0 = want(13, (const void*)mtc::message_fwd_tc<false>, smem_tc))) return rc0;
    }
    if (memo) {
      if ((rc0 = want(4, (const void*)message_fwd_memo<true>, sm_fwd0))) return rc0;
      if ((rc0 = want(5, (const void*)message_fwd_memo<false>, sm_fwd))) return rc0;
      if ((rc0 = want(6, (const void*)message_bwd_memo<true>, sm_bwdm0))) return rc0;
      if ((rc0 = want(7, (const void*)message_bwd_memo<false>, sm_bwdm))) return rc0;
      if ((rc0 = want(8, (const void*)message_bwd_memo_state, sm_state))) return rc0;
      if (pair_fwd0 && (rc0 = want(9, (const void*)message_fwd_memo_group<true, G_FWD0, T_FWD0>, sp_fwd0))) return rc0;
      if (pair_fwd && (rc0 = want(10, (const void*)message_fwd_memo_group<false, G_FWD, T_FWD>, sp_fwd))) return rc0;
      if (pair_state && (rc0 = want(11, (const void*)message_bwd_memo_state_group<G_STATE, T_STATE>, sp_state))) return rc0;
    }
  }

  VSSR_PROF(VSSR_K_GEOM, st, edge_geometry_kernel<<<ceil_div(A, 4), 128, 0, st>>>(
      pos, atom_ptr, cell, n_struct, A, rowptr, col, shift, (long long)e_cap, cutoff, memo ? fc : FilterCacheView{},
      w.nvalid, w.erec, w.nmemo, w.mrec, w.evex, w.grad0));
  if (staged)
    VSSR_PROF(VSSR_K_GEOM, st, row_order_kernel<<<n_struct, 128, (size_t)2 * nmax * sizeof(int32_t), st>>>(
        atom_ptr, w.nvalid, w.nmemo, w.order_d, w.order_m, memo ? fc.n0 : 0, fc.nmemo0, w.canonical));
  VSSR_PROF(VSSR_K_ELEMWISE, st, embed_kernel<<<dim3(ceil_div((long long)A * (F / 4), 256), M), 256, 0, st>>>(weights, z, A, w.s[0]));

  int rc;
  for (int l = 0; l < NCONV; ++l) {
    const float* wl = weights + W_LAYER0 + (long long)l * L_SIZE;
    GemmArgs g{};
    // F1
    g = GemmArgs{w.s[l], F, MA_F, wl + L_W1T, F, W_STRIDE, wl + L_B1, W_STRIDE, nullptr, 0, 0, nullptr, 0,
                 w.h1[l], F, MA_F, A, F, F, w.act};
    if ((rc = run_gemm<128, 0, 4>(g, wl + L_W1, M, st))) return rc;
    // F2
    g = GemmArgs{w.act, F, MA_F, wl + L_W2T, F3, W_STRIDE, wl + L_B2, W_STRIDE, nullptr, 0, 0, nullptr, 0,
                 w.phi[l], F3, (long long)A * F3, A, F3, F};
    if ((rc = run_gemm<128, 0, 1>(g, wl + L_W2, M, st))) return rc;
    // F3: memo pass (group kernels for canonical structures, one-structure kernels for the rest), then the direct pass
    // over 1..W sender windows (accumulating in window order)
    if (staged) {
      if (l == 0) {
        if (pair_fwd0)
          VSSR_PROF(VSSR_K_MSG_FWD_MEMO, st, message_fwd_memo_group<true, G_FWD0, T_FWD0><<<dim3(n_struct / G_FWD0, F / MSG_FC, M), T_FWD0, sp_fwd0, st>>>(
              l, A, atom_ptr, w.canonical, fc, w.phi[l], w.s[l], nullptr, w.cat[l], w.vmid[l], n_struct, ga_fwd0));
        if (memo)
          VSSR_PROF(VSSR_K_MSG_FWD_MEMO, st, message_fwd_memo<true><<<memo_grid, MEMO_THREADS_FWD, sm_fwd0, st>>>(
              l, A, atom_ptr, rowptr, w.order_m, w.nmemo, w.mrec, fc, w.phi[l], w.s[l], nullptr, w.cat[l], w.vmid[l],
              pair_fwd0 ? w.canonical : nullptr, n_struct, G_FWD0, ga_fwd0));
        if (tc_fwd)
          for (int win = 0; win < windows_of(cap_tc0); ++win)
            VSSR_PROF(VSSR_K_MSG_FWD, st, mtc::message_fwd_tc<true><<<v2_grid, mtc::THREADS, smem_tc0, st>>>(
                weights, l, A, atom_ptr, n_chunks, rowptr, w.nvalid, w.erec, w.phi[l], w.s[l], nullptr,
                w.cat[l], w.vmid[l], memo ? 1 : 0, cap_tc0, win));
        else
        for (int win = 0; win < windows_of(cap_fwd0); ++win)
          VSSR_PROF(VSSR_K_MSG_FWD, st, message_fwd_v2<true><<<v2_grid, MSG_THREADS, smem_fwd0, st>>>(
              weights, l, A, atom_ptr, n_chunks, rowptr, w.order_d, w.nvalid, w.erec, w.phi[l], w.s[l], nullptr,
              w.cat[l], w.vmid[l], memo ? 1 : 0, cap_fwd0, win));
      } else {
        if (pair_fwd)
          VSSR_PROF(VSSR_K_MSG_FWD_MEMO, st, message_fwd_memo_group<false, G_FWD, T_FWD><<<dim3(n_struct / G_FWD, F / MSG_FC, M), T_FWD, sp_fwd, st>>>(
              l, A, atom_ptr, w.canonical, fc, w.phi[l], w.s[l], w.v[l], w.cat[l], w.vmid[l], n_struct, ga_fwd));
        if (memo)
          VSSR_PROF(VSSR_K_MSG_FWD_MEMO, st, message_fwd_memo<false><<<memo_grid, MEMO_THREADS_FWD, sm_fwd, st>>>(
              l, A, atom_ptr, rowptr, w.order_m, w.nmemo, w.mrec, fc, w.phi[l], w.s[l], w.v[l], w.cat[l], w.vmid[l],
              pair_fwd ? w.canonical : nullptr, n_struct, G_FWD, ga_fwd));
        if (tc_fwd)
          for (int win = 0; win < windows_of(cap_tc); ++win)
            VSSR_PROF(VSSR_K_MSG_FWD, st, mtc::message_fwd_tc<false><<<v2_grid, mtc::THREADS, smem_tc, st>>>(
                weights, l, A, atom_ptr, n_chunks, rowptr, w.nvalid, w.erec, w.phi[l], w.s[l], w.v[l],
                w.cat[l], w.vmid[l], memo ? 1 : 0, cap_tc, win));
        else
        for (int win = 0; win < windows_of(cap_fwd); ++win)
          VSSR_PROF(VSSR_K_MSG_FWD, st, message_fwd_v2<false><<<v2_grid, MSG_THREADS, smem_fwd, st>>>(
              weights, l, A, atom_ptr, n_chunks, rowptr, w.order_d, w.nvalid, w.erec, w.phi[l], w.s[l], w.v[l],
              w.cat[l], w.vmid[l], memo ? 1 : 0, cap_fwd, win));
      }
    }
    // F4
    g = GemmArgs{w.vmid[l], F, (long long)A * 3 * F, wl + L_UVT, 2 * F, W_STRIDE, nullptr, 0, nullptr, 0, 0, nullptr, 0,
                 w.UV[l], 2 * F, (long long)A * 6 * F, 3 * A, 2 * F, F};
    if ((rc = run_gemm<128, 0, 0>(g, wl + L_UV, M, st))) return rc;
    // F5
    VSSR_PROF(VSSR_K_ELEMWISE, st, nrm_kernel<<<ew_grid, 256, 0, st>>>(w.UV[l], A, w.cat[l]));
    // F6
    g = GemmArgs{w.cat[l], 2 * F, (long long)A * 2 * F, wl + L_W3T, F, W_STRIDE, wl + L_B3, W_STRIDE, nullptr, 0, 0,
                 nullptr, 0, w.h3[l], F, MA_F, A, F, 2 * F, w.act};
    if ((rc = run_gemm<128, 0, 4>(g, wl + L_W3, M, st))) return rc;
    // F7
    g = GemmArgs{w.act, F, MA_F, wl + L_W4T, F3, W_STRIDE, wl + L_B4, W_STRIDE, nullptr, 0, 0, nullptr, 0,
                 w.a[l], F3, (long long)A * F3, A, F3, F};
    if ((rc = run_gemm<128, 0, 1>(g, wl + L_W4, M, st))) return rc;
    // F8
    VSSR_PROF(VSSR_K_ELEMWISE, st, update_fwd_kernel<<<ew_grid, 256, 0, st>>>(w.UV[l], w.a[l], w.cat[l], w.vmid[l], A,
                                                                           w.s[l + 1], w.v[l + 1]));
  }
  // readout
  {
    GemmArgs g{w.s[NCONV], F, MA_F, weights + R_W5T, FH, W_STRIDE, weights + R_B5, W_STRIDE, nullptr, 0, 0, nullptr, 0,
               w.h5, FH, (long long)A * FH, A, FH, F};
    if ((rc = run_gemm<64, 0, 1>(g, weights + R_W5, M, st))) return rc;
    VSSR_PROF(VSSR_K_READOUT, st, readout_energy_kernel<<<dim3(ceil_div(A, 4), M), 128, 0, st>>>(weights, w.h5, w.evex, A, w.e_atom));
    VSSR_PROF(VSSR_K_READOUT, st, energy_reduce_kernel<<<dim3(ceil_div(n_struct, 4), M), 128, 0, st>>>(w.e_atom, atom_ptr, n_struct, A, energy));
  }
  if (embedding) {
    VSSR_CUDA(cudaMemcpyAsync(embedding, w.s[NCONV], (size_t)M * A * F * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }

  // ---------------- backward ----------------
  VSSR_PROF(VSSR_K_ELEMWISE, st, grad_init_kernel<<<dim3(ceil_div(3 * A, 256), M), 256, 0, st>>>(w.grad0, 3 * A, grad));
  {
    // ds = (dswish(h5) * w6) . W5      [A,64]x[64,128]
    GemmArgs g{w.h5, FH, (long long)A * FH, weights + R_W5, F, W_STRIDE, nullptr, 0, nullptr, 0, 0, weights + R_W6,
               W_STRIDE, w.ds, F, MA_F, A, F, FH};
    if ((rc = run_gemm<128, 2, 0>(g, weights + R_W5T, M, st))) return rc;
  }
  VSSR_CUDA(cudaMemsetAsync(w.dvA, 0, (size_t)M * A * 3 * F * sizeof(float), st));
  float* dv_cur = w.dvA;
  float* dv_nxt = w.dvB;
  for (int l = NCONV - 1; l >= 0; --l) {
    const float* wl = weights + W_LAYER0 + (long long)l * L_SIZE;
    GemmArgs g{};
    // B8
    VSSR_PROF(VSSR_K_ELEMWISE, st, update_bwd_kernel<<<ew_grid, 256, 0, st>>>(w.ds, dv_cur, w.UV[l], w.a[l], A, w.da, w.dUV));
    // B7: dh3 = (da . W4) * dswish(h3)
    g = GemmArgs{w.da, F3, (long long)A * F3, wl + L_W4, F, W_STRIDE, nullptr, 0, w.h3[l], F, MA_F, nullptr, 0,
                 w.dh3, F, MA_F, A, F, F3};
    if ((rc = run_gemm<128, 0, 2>(g, wl + L_W4T, M, st))) return rc;
    // B6: dcat = dh3 . W3
    g = GemmArgs{w.dh3, F, MA_F, wl + L_W3, 2 * F, W_STRIDE, nullptr, 0, nullptr, 0, 0, nullptr, 0,
                 w.dcat, 2 * F, (long long)A * 2 * F, A, 2 * F, F};
    if ((rc = run_gemm<128, 0, 0>(g, wl + L_W3T, M, st))) return rc;
    // B5
    VSSR_PROF(VSSR_K_ELEMWISE, st, nrm_bwd_kernel<<<ew_grid, 256, 0, st>>>(w.dcat, w.cat[l], w.UV[l], A, w.ds, w.dUV));
    // B4: dv += dUV . [U;V]      [3A,256]x[256,128]
    g = GemmArgs{w.dUV, 2 * F, (long long)A * 6 * F, wl + L_UV, F, W_STRIDE, nullptr, 0, nullptr, 0, 0, nullptr, 0,
                 dv_cur, F, (long long)A * 3 * F, 3 * A, F, 2 * F};
    if ((rc = run_gemm<128, 0, 3>(g, wl + L_UVT, M, st))) return rc;
    // B3
    if (staged) {
      // accum bit 0: dphi/dv_in were started by a memo pass; bit 1: so was gradp (later sender windows set both)
      if (l == 0) {
        if (memo && !constrained)
          VSSR_PROF(VSSR_K_MSG_BWD_MEMO, st, message_bwd_memo<true><<<memo_grid, MEMO_THREADS_BWD, sm_bwdm0, st>>>(
              l, A, atom_ptr, rowptr, w.order_m, w.nmemo, w.mrec, fc, w.phi[l], nullptr, w.ds, dv_cur, nullptr, nullptr, w.gradp));
        for (int win = 0; win < windows_of(cap_bwd0); ++win)
          VSSR_PROF(VSSR_K_MSG_BWD, st, message_bwd_v2<true><<<v2_grid, MSG_THREADS, smem_bwd0, st>>>(
              weights, l, A, atom_ptr, n_chunks, rowptr, w.order_d, w.nvalid, w.erec, w.phi[l], nullptr, w.ds,
              dv_cur, nullptr, nullptr, w.gradp, (memo && !constrained) ? 3 : 0, constrained ? fc.frozen : nullptr, fc.n0,
              cap_bwd0, win));
      } else {
        if (pair_state)
          VSSR_PROF(VSSR_K_MSG_BWD_MEMO, st, message_bwd_memo_state_group<G_STATE, T_STATE><<<dim3(n_struct / G_STATE, F / MSG_FC, M), T_STATE, sp_state, st>>>(
              l, A, atom_ptr, w.canonical, fc, w.phi[l], w.v[l], w.ds, dv_cur, w.dphi, dv_nxt, n_struct, ga_state));
        if (constrained)
          VSSR_PROF(VSSR_K_MSG_BWD_MEMO, st, message_bwd_memo_state<<<memo_grid, MEMO_THREADS_FWD, sm_state, st>>>(
              l, A, atom_ptr, rowptr, w.order_m, w.nmemo, w.mrec, fc, w.phi[l], w.v[l], w.ds, dv_cur, w.dphi, dv_nxt,
              pair_state ? w.canonical : nullptr, n_struct, G_STATE, ga_state));
        else if (memo)
          VSSR_PROF(VSSR_K_MSG_BWD_MEMO, st, message_bwd_memo<false><<<memo_grid, MEMO_THREADS_BWD, sm_bwdm, st>>>(
              l, A, atom_ptr, rowptr, w.order_m, w.nmemo, w.mrec, fc, w.phi[l], w.v[l], w.ds, dv_cur, w.dphi, dv_nxt, w.gradp));
        for (int win = 0; win < windows_of(cap_bwd); ++win)
          VSSR_PROF(VSSR_K_MSG_BWD, st, message_bwd_v2<false><<<v2_grid, MSG_THREADS, smem_bwd, st>>>(
              weights, l, A, atom_ptr, n_chunks, rowptr, w.order_d, w.nvalid, w.erec, w.phi[l], w.v[l], w.ds,
              dv_cur, w.dphi, dv_nxt, w.gradp, constrained ? 1 : (memo ? 3 : 0), constrained ? fc.frozen : nullptr, fc.n0,
              cap_bwd, win));
      }
      VSSR_PROF(VSSR_K_ELEMWISE, st, grad_accum_kernel<<<dim3(ceil_div(3 * A, 256), M), 256, 0, st>>>(w.gradp, 3 * A, grad));
    }
    if (l > 0) {
      // B2: dh1 = (dphi . W2) * dswish(h1)
      g = GemmArgs{w.dphi, F3, (long long)A * F3, wl + L_W2, F, W_STRIDE, nullptr, 0, w.h1[l], F, MA_F, nullptr, 0,
                   w.dh1, F, MA_F, A, F, F3};
    if ((rc = run_gemm<128, 0, 2>(g, wl + L_W2T, M, st))) return rc;
      // B1: ds += dh1 . W1
      g = GemmArgs{w.dh1, F, MA_F, wl + L_W1, F, W_STRIDE, nullptr, 0, nullptr, 0, 0, nullptr, 0,
                   w.ds, F, MA_F, A, F, F};
    if ((rc = run_gemm<128, 0, 3>(g, wl + L_W1T, M, st))) return rc;
      float* t = dv_cur; dv_cur = dv_nxt; dv_nxt = t;
    }
  }
  if (constrained)
    VSSR_PROF(VSSR_K_ELEMWISE, st, zero_frozen_grad_kernel<<<dim3(ceil_div(A, 256), M), 256, 0, st>>>(
        atom_ptr, n_struct, A, fc.n0, fc.frozen, grad));
  return VSSR_OK;
}

namespace {
// out[0] direct edges, [1] memoised edges, [2] direct edges whose receiver is a frozen framework atom,
// [3] canonical structures, [4] edges of the 6 A list -- of the LAST evaluation that used this workspace
__global__ void edge_stats_kernel(const int32_t* __restrict__ atom_ptr, int n_struct, int n_atoms,
                                  const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nvalid,
                                  const int32_t* __restrict__ nmemo, const int32_t* __restrict__ canonical, int n0,
                                  const uint8_t* __restrict__ frozen, unsigned long long* __restrict__ out) {
  unsigned long long d = 0, m = 0, df = 0, c = 0;
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n_atoms; a += gridDim.x * blockDim.x) {
    const int il = a - __ldg(atom_ptr + struct_of_atom(atom_ptr, n_struct, a));
    const int nv = __ldg(nvalid + a);
    d += nv;
    m += nmemo ? __ldg(nmemo + a) : 0;
    if (frozen && il < n0 && frozen[il]) df += nv;
    if (canonical && a < n_struct) c += __ldg(canonical + a) != 0;
  }
  atomicAdd(out, d); atomicAdd(out + 1, m); atomicAdd(out + 2, df); atomicAdd(out + 3, c);
  if (blockIdx.x == 0 && threadIdx.x == 0) out[4] = (unsigned long long)__ldg(rowptr + n_atoms);
}
}  // namespace

extern "C" int vssr_painn_edge_stats(const void* workspace, int32_t n_models, int32_t n_atoms, int64_t e_cap,
                                     const int32_t* atom_ptr, int32_t n_struct, const int32_t* rowptr,
                                     const void* filter_cache, int32_t fc_n0, int64_t fc_e_cap0, int64_t* out5,
                                     void* stream) {
  if (!workspace || !atom_ptr || !rowptr || !out5 || n_atoms <= 0 || n_struct <= 0) return VSSR_ERR_ARG;
  const Workspace w = carve(const_cast<void*>(workspace), n_models, n_atoms, e_cap);
  const FilterCacheView fc = cache_view(filter_cache, n_models, fc_n0, fc_e_cap0);
  cudaStream_t st = (cudaStream_t)stream;
  VSSR_CUDA(cudaMemsetAsync(out5, 0, 5 * sizeof(int64_t), st));
  edge_stats_kernel<<<ceil_div(n_atoms, 256), 256, 0, st>>>(atom_ptr, n_struct, n_atoms, rowptr, w.nvalid,
                                                            fc.n0 > 0 ? w.nmemo : nullptr, fc.n0 > 0 ? w.canonical : nullptr,
                                                            fc.n0, fc.frozen, reinterpret_cast<unsigned long long*>(out5));
  VSSR_LAUNCH_CHECK();
  return VSSR_OK;
}

extern "C" size_t vssr_painn_filter_cache_bytes(int32_t n_models, int32_t n0, int64_t e_cap0) {
  return carve_cache(nullptr, n_models, n0, e_cap0).bytes;
}

namespace {
// scratch of the memo build: one-structure neighbour list + the geometry kernel's outputs
struct CacheScratch {
  int32_t *atom_ptr, *deg, *col, *status, *nmemo;
  int8_t* shift;
  float *erec, *mrec, *evex, *grad0;
  size_t bytes;
};
CacheScratch carve_cache_scratch(void* base, int n0, long long e_cap0) {
  size_t off = 0;
  auto take = [&](size_t nbytes) -> void* { void* p = reinterpret_cast<char*>(base) + off; off += ((nbytes + 255) / 256) * 256; return p; };
  CacheScratch s;
  s.atom_ptr = (int32_t*)take(8);
  s.deg = (int32_t*)take((size_t)n0 * 4);
  s.col = (int32_t*)take((size_t)e_cap0 * 4);
  s.shift = (int8_t*)take((size_t)e_cap0 * 4);
  s.status = (int32_t*)take(4);
  s.erec = (float*)take((size_t)e_cap0 * REC * 4);
  s.nmemo = (int32_t*)take((size_t)n0 * 4);
  s.mrec = (float*)take((size_t)e_cap0 * MREC * 4);
  s.evex = (float*)take((size_t)n0 * 4);
  s.grad0 = (float*)take((size_t)n0 * 12);
  s.bytes = off;
  return s;
}
}  // namespace

extern "C" size_t vssr_painn_filter_cache_workspace_bytes(int32_t n0, int64_t e_cap0) {
  return carve_cache_scratch(nullptr, n0, e_cap0).bytes;
}

extern "C" int vssr_painn_filter_cache_build(const float* weights, int32_t n_models, const float* pos0, const float* cell,
                                             const uint8_t* pbc, const uint8_t* fixed0, int32_t n0, float cutoff,
                                             float skin, int64_t e_cap0, void* cache, size_t cache_bytes, void* workspace,
                                             size_t workspace_bytes, int32_t* nslots_out, void* stream) {
  if (!weights || !pos0 || !cell || !pbc || !fixed0 || !cache || !workspace || n0 <= 0 || n_models <= 0) return VSSR_ERR_ARG;
  CacheBlob c = carve_cache(cache, n_models, n0, e_cap0);
  if (c.bytes > cache_bytes) return VSSR_ERR_WORKSPACE;
  const CacheScratch sc = carve_cache_scratch(workspace, n0, e_cap0);
  if (sc.bytes > workspace_bytes) return VSSR_ERR_WORKSPACE;
  int32_t *atom_ptr = sc.atom_ptr, *deg = sc.deg, *col = sc.col, *status = sc.status, *nmemo = sc.nmemo;
  int8_t* shift = sc.shift;
  float *erec = sc.erec, *mrec = sc.mrec, *evex = sc.evex, *grad0 = sc.grad0;
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t ptr_h[2] = {0, n0};
  VSSR_CUDA(cudaMemcpyAsync(atom_ptr, ptr_h, 8, cudaMemcpyHostToDevice, st));
  VSSR_CUDA(cudaMemsetAsync(status, 0, 4, st));
  VSSR_CUDA(cudaMemsetAsync(c.slot, 0xFF, (size_t)e_cap0 * 4, st));
  VSSR_CUDA(cudaMemcpyAsync(c.frozen, fixed0, (size_t)n0, cudaMemcpyDeviceToDevice, st));
  int rc = vssr_nbr_build(pos0, atom_ptr, cell, pbc, 1, n0, cutoff + skin, deg, c.rowptr, col, shift, e_cap0, status, stream);
  if (rc) return rc;
  VSSR_PROF(VSSR_K_GEOM, st, edge_geometry_kernel<<<ceil_div(n0, 4), 128, 0, st>>>(
      pos0, atom_ptr, cell, 1, n0, c.rowptr, col, shift, (long long)e_cap0, cutoff, FilterCacheView{}, c.nvalid, erec, nmemo, mrec,
      evex, grad0));
  VSSR_PROF(VSSR_K_GEOM, st, cache_slot_kernel<<<1, 32, 0, st>>>(erec, c.rowptr, c.nvalid, fixed0, n0, c.key, c.slot, c.d, c.counter));
  VSSR_PROF(VSSR_K_GEOM, st, cache_fill_kernel<<<dim3((unsigned)e_cap0, n_models * NCONV), 128, 0, st>>>(
      weights, erec, c.slot, (int)e_cap0, c.wc, c.qc));
  // the framework's own memoised lists: the same geometry kernel, now WITH the memo, run on the framework itself
  {
    const FilterCacheView self = cache_view(cache, n_models, n0, e_cap0);
    VSSR_PROF(VSSR_K_GEOM, st, edge_geometry_kernel<<<ceil_div(n0, 4), 128, 0, st>>>(
        pos0, atom_ptr, cell, 1, n0, c.rowptr, col, shift, (long long)e_cap0, cutoff, self, deg, erec, c.nmemo0, c.mrec0,
        evex, grad0));
    VSSR_PROF(VSSR_K_GEOM, st, row_order_kernel<<<1, 128, (size_t)2 * n0 * sizeof(int32_t), st>>>(
        atom_ptr, c.nmemo0, c.nmemo0, c.order0, nmemo, 0, nullptr, nullptr));
  }
  int32_t host[2] = {0, 0};
  VSSR_CUDA(cudaMemcpyAsync(&host[0], c.counter, 4, cudaMemcpyDeviceToHost, st));
  VSSR_CUDA(cudaMemcpyAsync(&host[1], status, 4, cudaMemcpyDeviceToHost, st));
  VSSR_CUDA(cudaStreamSynchronize(st));
  if (host[1] & VSSR_STATUS_EDGE_OVERFLOW) return VSSR_ERR_WORKSPACE;
  if (nslots_out) *nslots_out = host[0];
  return VSSR_OK;
}
