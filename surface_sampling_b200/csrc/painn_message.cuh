// PaiNN message passing, forward and backward — shared-memory staged, FFMA2-packed (sm_100).
// Included by painn.cu inside its anonymous namespace.
//
// Work decomposition: CTA = (structure, atom-chunk) x feature-half (64 of the 128 features) x model.
// The CTA first stages, for its 64 features, the per-atom rows it will gather from EVERY atom of
// the structure into shared memory (phi, v [, ds, dv]); after one barrier each of the 8 warps owns
// one receiver atom at a time and lane l owns the adjacent feature pair (2l, 2l+1) of the chunk.
// All arithmetic on the pair is issued as packed fp32x2 instructions (fma.rn.f32x2 -> SASS FFMA2),
// which on B200 sustain 66 TFLOP/s vs 42 for scalar FFMA (profiles/microbench/ffma2.cu).  The
// radial filter w(d) = Wd.(rbf*env) + bd*env lives in 60 register pairs per lane.
// Per-edge records (384 B: unit vector, sender, rbf rows pre-duplicated as (v,v) pairs) stream
// through a private 3-stage cp.async ring per warp, so the L2 latency of the next edges hides
// behind the current edge's ~60-120 FFMA2.
// Determinism: a lane walks its receiver's CSR row serially; there are no atomics.
#pragma once

constexpr int MSG_FC = 64;        // features per CTA
constexpr int MSG_THREADS = 256;  // 8 warps
constexpr int MSG_WARPS = MSG_THREADS / 32;
constexpr int MSG_STAGES = 3;     // cp.async ring depth per warp
constexpr int MSG_RS_FWD = REC + 3 * MSG_FC;   // ring stage floats: edge record + memoised w rows of the chunk
constexpr int MSG_RS_BWD = REC + 6 * MSG_FC;   // ... + memoised q rows
constexpr int MSG_PIPE_BYTES_FWD = MSG_WARPS * MSG_STAGES * MSG_RS_FWD * 4;
constexpr int MSG_PIPE_BYTES_BWD = MSG_WARPS * MSG_STAGES * MSG_RS_BWD * 4;

__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }

// smem floats per staged atom
template <bool FIRST> struct MsgFwdLayout { static constexpr int PER = FIRST ? 3 * MSG_FC : 6 * MSG_FC; };
template <bool FIRST> struct MsgBwdLayout { static constexpr int PER = FIRST ? 7 * MSG_FC : 10 * MSG_FC; };

// copy `rows` rows of MSG_FC floats: src row r at src + r*src_stride, dst at dst + r*MSG_FC (per atom a)
__device__ __forceinline__ void stage_rows(float* __restrict__ dst_atom0, int per, int dst_off,
                                           const float* __restrict__ src, long long src_atom_stride, int src_row_stride,
                                           int rows, int n, int tid) {
  constexpr int Q = MSG_FC / 4;  // float4 per row
  const int per_atom = rows * Q;
  for (int idx = tid; idx < n * per_atom; idx += MSG_THREADS) {
    const int a = idx / per_atom, r = (idx % per_atom) / Q, c4 = idx % Q;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + a * src_atom_stride + r * src_row_stride) + c4);
    *reinterpret_cast<float4*>(dst_atom0 + a * per + dst_off + r * MSG_FC + c4 * 4) = v;
  }
}

// one 16-byte cp.async per lane for the first n16*16 bytes of a record; if the edge's filter is
// memoised (slot >= 0) each lane also pulls its feature pair of the NK cached rows (8 bytes each)
// behind the record; then commit (always commits, so group accounting stays uniform past the row end)
template <int NK>
__device__ __forceinline__ void prefetch_record(float* dst, const float* src, int lane, int n16, bool live, int slot,
                                                const float* __restrict__ wrow, const float* __restrict__ qrow) {
  if (live) {
    if (lane < n16) {
      const unsigned d = (unsigned)__cvta_generic_to_shared(dst + lane * 4);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src + lane * 4));
    }
    if (slot >= 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + REC + (k * 32 + lane) * 2);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(wrow + (long long)slot * F3 + k * F + 2 * lane));
      }
      if (NK == 6) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const unsigned d = (unsigned)__cvta_generic_to_shared(dst + REC + 3 * MSG_FC + (k * 32 + lane) * 2);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(qrow + (long long)slot * F3 + k * F + 2 * lane));
        }
      }
    }
  }
  asm volatile("cp.async.commit_group;\n" ::);
}
// memo slot of the e-th valid edge of the current row: lanes hold edges [0,64) in two registers
__device__ __forceinline__ int row_slot(int s0, int s1, const int32_t* __restrict__ eslot_row, int e, int ne) {
  if (e >= ne) return -1;
  if (e < 32) return __shfl_sync(0xffffffffu, s0, e);
  if (e < 64) return __shfl_sync(0xffffffffu, s1, e - 32);
  return __ldg(eslot_row + e);
}
__device__ __forceinline__ void wait_record() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(MSG_STAGES - 2));
  __syncwarp();
}

template <bool FIRST>
__global__ void __launch_bounds__(MSG_THREADS, 1) message_fwd_v2(
    const float* __restrict__ weights, int layer, int n_atoms, const int32_t* __restrict__ atom_ptr, int n_chunks,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nvalid, const float* __restrict__ erec,
    const int32_t* __restrict__ eslot, FilterCacheView fc, const float* __restrict__ phi,
    const float* __restrict__ s_in, const float* __restrict__ v_in, float* __restrict__ cat,
    float* __restrict__ v_mid) {
  extern __shared__ __align__(16) float smem_all[];
  constexpr int PER = MsgFwdLayout<FIRST>::PER;
  constexpr int N16 = (REC_RE + 44) / 4;  // 13 x 16 B: geometry + rbf rows
  constexpr int RS = MSG_RS_FWD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* ring = smem_all + warp * MSG_STAGES * RS;
  float* smem = smem_all + MSG_WARPS * MSG_STAGES * RS;
  const int b = blockIdx.x / n_chunks, ch = blockIdx.x % n_chunks;
  const int h = blockIdx.y, m = blockIdx.z;
  const int a0 = __ldg(atom_ptr + b), n = __ldg(atom_ptr + b + 1) - a0;
  const long long mA = (long long)m * n_atoms;
  phi += (mA + a0) * F3 + h * MSG_FC;
  s_in += (mA + a0) * F;
  cat += (mA + a0) * 2 * F;
  v_mid += (mA + a0) * 3 * F;
  if (!FIRST) v_in += (mA + a0) * 3 * F + h * MSG_FC;

  stage_rows(smem, PER, 0, phi, F3, F, 3, n, tid);
  if (!FIRST) stage_rows(smem, PER, 3 * MSG_FC, v_in, 3 * F, F, 3, n, tid);

  const int f0 = h * MSG_FC + 2 * lane;  // first feature of this lane's pair
  const float* __restrict__ wl = weights + (long long)m * W_STRIDE + W_LAYER0 + (long long)layer * L_SIZE;
  // memoised filter rows of this (model, layer), pre-offset to this CTA's feature half
  const float* __restrict__ wrow = fc.wc ? fc.wc + (long long)(m * NCONV + layer) * fc.nslots_cap * F3 + h * MSG_FC : nullptr;
  float2 wd0[NRBF], wd1[NRBF], wd2[NRBF];
#pragma unroll
  for (int q = 0; q < NRBF; ++q) {
    wd0[q] = ld2(wl + L_WDT + q * F3 + f0);
    wd1[q] = ld2(wl + L_WDT + q * F3 + F + f0);
    wd2[q] = ld2(wl + L_WDT + q * F3 + 2 * F + f0);
  }
  const float2 bd0 = ld2(wl + L_BD + f0), bd1 = ld2(wl + L_BD + F + f0), bd2 = ld2(wl + L_BD + 2 * F + f0);
  __syncthreads();

  for (int il = ch + n_chunks * warp; il < n; il += n_chunks * MSG_WARPS) {
    const int i = a0 + il;
    const long long e0 = __ldg(rowptr + i);
    const float* rec0 = erec + e0 * REC;
    const int ne = __ldg(nvalid + i);
    const int32_t* srow = eslot + e0;
    const int s0r = lane < ne ? __ldg(srow + lane) : -1, s1r = lane + 32 < ne ? __ldg(srow + lane + 32) : -1;
    float2 ds = dup2(0.f), dvx = dup2(0.f), dvy = dup2(0.f), dvz = dup2(0.f);
    __syncwarp();
#pragma unroll
    for (int s = 0; s < MSG_STAGES - 1; ++s)
      prefetch_record<3>(ring + s * RS, rec0 + (long long)s * REC, lane, N16, s < ne, row_slot(s0r, s1r, srow, s, ne), wrow, nullptr);
    for (int e = 0; e < ne; ++e) {
      wait_record();
      const int nx = e + MSG_STAGES - 1;
      prefetch_record<3>(ring + (nx % MSG_STAGES) * RS, rec0 + (long long)nx * REC, lane, N16, nx < ne,
                         row_slot(s0r, s1r, srow, nx, ne), wrow, nullptr);
      const float* rec = ring + (e % MSG_STAGES) * RS;
      const float4 g = *reinterpret_cast<const float4*>(rec);
      const float* sj = smem + (__float_as_int(rec[REC_EJ]) - a0) * PER + 2 * lane;
      const float2 p0 = ld2(sj), p1 = ld2(sj + MSG_FC), p2 = ld2(sj + 2 * MSG_FC);
      float2 w0, w1, w2;
      if (__float_as_int(rec[REC_SLOT]) >= 0) {   // memoised filter (warp-uniform branch)
        w0 = ld2(rec + REC + 2 * lane); w1 = ld2(rec + REC + MSG_FC + 2 * lane); w2 = ld2(rec + REC + 2 * MSG_FC + 2 * lane);
      } else {
        const float4* r4 = reinterpret_cast<const float4*>(rec + REC_RE);
        const float4 ev = r4[10];  // (env,env,denv,denv)
        const float2 env2 = make_float2(ev.x, ev.y);
        w0 = __fmul2_rn(bd0, env2); w1 = __fmul2_rn(bd1, env2); w2 = __fmul2_rn(bd2, env2);
#pragma unroll
        for (int q = 0; q < NRBF / 2; ++q) {
          const float4 t = r4[q];
          const float2 ra = make_float2(t.x, t.y), rb = make_float2(t.z, t.w);
          w0 = __ffma2_rn(wd0[2 * q], ra, w0); w1 = __ffma2_rn(wd1[2 * q], ra, w1); w2 = __ffma2_rn(wd2[2 * q], ra, w2);
          w0 = __ffma2_rn(wd0[2 * q + 1], rb, w0); w1 = __ffma2_rn(wd1[2 * q + 1], rb, w1);
          w2 = __ffma2_rn(wd2[2 * q + 1], rb, w2);
        }
      }
      const float2 x0 = __fmul2_rn(p0, w0), x1 = __fmul2_rn(p1, w1), x2 = __fmul2_rn(p2, w2);
      ds = __fadd2_rn(ds, x1);
      dvx = __ffma2_rn(x2, dup2(g.x), dvx);
      dvy = __ffma2_rn(x2, dup2(g.y), dvy);
      dvz = __ffma2_rn(x2, dup2(g.z), dvz);
      if (!FIRST) {
        const float2 vx = ld2(sj + 3 * MSG_FC), vy = ld2(sj + 4 * MSG_FC), vz = ld2(sj + 5 * MSG_FC);
        dvx = __ffma2_rn(x0, vx, dvx); dvy = __ffma2_rn(x0, vy, dvy); dvz = __ffma2_rn(x0, vz, dvz);
      }
    }
    const float2 s0 = ld2(s_in + (long long)il * F + f0);
    *reinterpret_cast<float2*>(cat + (long long)il * 2 * F + f0) = __fadd2_rn(s0, ds);
    if (!FIRST) {
      const float* si = smem + il * PER + 2 * lane;
      dvx = __fadd2_rn(dvx, ld2(si + 3 * MSG_FC));
      dvy = __fadd2_rn(dvy, ld2(si + 4 * MSG_FC));
      dvz = __fadd2_rn(dvz, ld2(si + 5 * MSG_FC));
    }
    float* vo = v_mid + (long long)il * 3 * F + f0;
    *reinterpret_cast<float2*>(vo) = dvx;
    *reinterpret_cast<float2*>(vo + F) = dvy;
    *reinterpret_cast<float2*>(vo + 2 * F) = dvz;
  }
}

// Backward: gather over the receiver's own row; edge A = i<-j together with its reverse B = j<-i.
// Outputs dphi[i], dv_in[i] (both skipped for the first layer) and the per-feature-half partial of
// dE/dx_i in gradp[m][h][i][3] (summed over halves in fixed order by grad_accum_kernel).
template <bool FIRST>
__global__ void __launch_bounds__(MSG_THREADS, 1) message_bwd_v2(
    const float* __restrict__ weights, int layer, int n_atoms, const int32_t* __restrict__ atom_ptr, int n_chunks,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nvalid, const float* __restrict__ erec,
    const int32_t* __restrict__ eslot, FilterCacheView fc, const float* __restrict__ phi,
    const float* __restrict__ v_in, const float* __restrict__ ds, const float* __restrict__ dv,
    float* __restrict__ dphi, float* __restrict__ dv_in, float* __restrict__ gradp) {
  extern __shared__ __align__(16) float smem_all[];
  constexpr int PER = MsgBwdLayout<FIRST>::PER;
  constexpr int O_V = 3 * MSG_FC;                          // only when !FIRST
  constexpr int O_DS = FIRST ? 3 * MSG_FC : 6 * MSG_FC;
  constexpr int O_DV = O_DS + MSG_FC;
  constexpr int N16 = (REC_DRE + 40) / 4;                  // 23 x 16 B: whole record
  constexpr int RS = MSG_RS_BWD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* ring = smem_all + warp * MSG_STAGES * RS;
  float* smem = smem_all + MSG_WARPS * MSG_STAGES * RS;
  const int b = blockIdx.x / n_chunks, ch = blockIdx.x % n_chunks;
  const int h = blockIdx.y, m = blockIdx.z;
  const int a0 = __ldg(atom_ptr + b), n = __ldg(atom_ptr + b + 1) - a0;
  const long long mA = (long long)m * n_atoms;
  phi += (mA + a0) * F3 + h * MSG_FC;
  ds += (mA + a0) * F + h * MSG_FC;
  dv += (mA + a0) * 3 * F + h * MSG_FC;
  if (!FIRST) {
    v_in += (mA + a0) * 3 * F + h * MSG_FC;
    dphi += (mA + a0) * F3;
    dv_in += (mA + a0) * 3 * F;
  }
  stage_rows(smem, PER, 0, phi, F3, F, 3, n, tid);
  if (!FIRST) stage_rows(smem, PER, O_V, v_in, 3 * F, F, 3, n, tid);
  stage_rows(smem, PER, O_DS, ds, F, F, 1, n, tid);
  stage_rows(smem, PER, O_DV, dv, 3 * F, F, 3, n, tid);

  const int f0 = h * MSG_FC + 2 * lane;
  const float* __restrict__ wl = weights + (long long)m * W_STRIDE + W_LAYER0 + (long long)layer * L_SIZE;
  const float* __restrict__ wrow = fc.wc ? fc.wc + (long long)(m * NCONV + layer) * fc.nslots_cap * F3 + h * MSG_FC : nullptr;
  const float* __restrict__ qrow = fc.qc ? fc.qc + (long long)(m * NCONV + layer) * fc.nslots_cap * F3 + h * MSG_FC : nullptr;
  float2 wd0[NRBF], wd1[NRBF], wd2[NRBF];
#pragma unroll
  for (int q = 0; q < NRBF; ++q) {
    wd0[q] = ld2(wl + L_WDT + q * F3 + f0);
    wd1[q] = ld2(wl + L_WDT + q * F3 + F + f0);
    wd2[q] = ld2(wl + L_WDT + q * F3 + 2 * F + f0);
  }
  const float2 bd0 = ld2(wl + L_BD + f0), bd1 = ld2(wl + L_BD + F + f0), bd2 = ld2(wl + L_BD + 2 * F + f0);
  __syncthreads();

  for (int il = ch + n_chunks * warp; il < n; il += n_chunks * MSG_WARPS) {
    const int i = a0 + il;
    const long long e0 = __ldg(rowptr + i);
    const float* rec0 = erec + e0 * REC;
    const int ne = __ldg(nvalid + i);
    const int32_t* srow = eslot + e0;
    const int s0r = lane < ne ? __ldg(srow + lane) : -1, s1r = lane + 32 < ne ? __ldg(srow + lane + 32) : -1;
    const float* si = smem + il * PER + 2 * lane;
    const float2 pi0 = ld2(si), pi1 = ld2(si + MSG_FC), pi2 = ld2(si + 2 * MSG_FC);
    const float2 gsi = ld2(si + O_DS);
    const float2 gvix = ld2(si + O_DV), gviy = ld2(si + O_DV + MSG_FC), gviz = ld2(si + O_DV + 2 * MSG_FC);
    float2 vix = dup2(0.f), viy = dup2(0.f), viz = dup2(0.f);
    if (!FIRST) { vix = ld2(si + O_V); viy = ld2(si + O_V + MSG_FC); viz = ld2(si + O_V + 2 * MSG_FC); }
    float2 dp0 = dup2(0.f), dp1 = dup2(0.f), dp2n = dup2(0.f);   // dphi_i (dp2n holds -dphi_i[2])
    float2 dvx = dup2(0.f), dvy = dup2(0.f), dvz = dup2(0.f);    // sender-side dv_in
    float2 gnx = dup2(0.f), gny = dup2(0.f), gnz = dup2(0.f);    // -(per-feature dE/dx_i)
    __syncwarp();
#pragma unroll
    for (int s = 0; s < MSG_STAGES - 1; ++s)
      prefetch_record<6>(ring + s * RS, rec0 + (long long)s * REC, lane, N16, s < ne, row_slot(s0r, s1r, srow, s, ne), wrow, qrow);
    for (int e = 0; e < ne; ++e) {
      wait_record();
      const int nx = e + MSG_STAGES - 1;
      prefetch_record<6>(ring + (nx % MSG_STAGES) * RS, rec0 + (long long)nx * REC, lane, N16, nx < ne,
                         row_slot(s0r, s1r, srow, nx, ne), wrow, qrow);
      const float* rec = ring + (e % MSG_STAGES) * RS;
      const float4 g = *reinterpret_cast<const float4*>(rec);
      const float* sj = smem + (__float_as_int(rec[REC_EJ]) - a0) * PER + 2 * lane;
      const float2 pj0 = ld2(sj), pj1 = ld2(sj + MSG_FC), pj2 = ld2(sj + 2 * MSG_FC);
      const float2 gsj = ld2(sj + O_DS);
      const float2 gvjx = ld2(sj + O_DV), gvjy = ld2(sj + O_DV + MSG_FC), gvjz = ld2(sj + O_DV + 2 * MSG_FC);
      float2 w0, w1, w2, q0, q1, q2;
      if (__float_as_int(rec[REC_SLOT]) >= 0) {   // memoised filter and derivative (warp-uniform branch)
        const float* mw = rec + REC + 2 * lane;
        w0 = ld2(mw); w1 = ld2(mw + MSG_FC); w2 = ld2(mw + 2 * MSG_FC);
        q0 = ld2(mw + 3 * MSG_FC); q1 = ld2(mw + 4 * MSG_FC); q2 = ld2(mw + 5 * MSG_FC);
      } else {
        const float4* r4 = reinterpret_cast<const float4*>(rec + REC_RE);
        const float4* d4 = reinterpret_cast<const float4*>(rec + REC_DRE);
        const float4 ev = r4[10];
        const float2 env2 = make_float2(ev.x, ev.y), denv2 = make_float2(ev.z, ev.w);
        w0 = __fmul2_rn(bd0, env2); w1 = __fmul2_rn(bd1, env2); w2 = __fmul2_rn(bd2, env2);
        q0 = __fmul2_rn(bd0, denv2); q1 = __fmul2_rn(bd1, denv2); q2 = __fmul2_rn(bd2, denv2);
#pragma unroll
        for (int q = 0; q < NRBF / 2; ++q) {
          const float4 t = r4[q];
          const float4 u = d4[q];
          const float2 ra = make_float2(t.x, t.y), rb = make_float2(t.z, t.w);
          const float2 da = make_float2(u.x, u.y), db = make_float2(u.z, u.w);
          w0 = __ffma2_rn(wd0[2 * q], ra, w0); w1 = __ffma2_rn(wd1[2 * q], ra, w1); w2 = __ffma2_rn(wd2[2 * q], ra, w2);
          q0 = __ffma2_rn(wd0[2 * q], da, q0); q1 = __ffma2_rn(wd1[2 * q], da, q1); q2 = __ffma2_rn(wd2[2 * q], da, q2);
          w0 = __ffma2_rn(wd0[2 * q + 1], rb, w0); w1 = __ffma2_rn(wd1[2 * q + 1], rb, w1);
          w2 = __ffma2_rn(wd2[2 * q + 1], rb, w2);
          q0 = __ffma2_rn(wd0[2 * q + 1], db, q0); q1 = __ffma2_rn(wd1[2 * q + 1], db, q1);
          q2 = __ffma2_rn(wd2[2 * q + 1], db, q2);
        }
      }
      const float2 ux = dup2(g.x), uy = dup2(g.y), uz = dup2(g.z);
      // edge A (i receives from j): dxA1 = gsi, dxA2 = gvi.u, dxA0 = gvi.vj
      const float2 dxA2 = __ffma2_rn(gviz, uz, __ffma2_rn(gviy, uy, __fmul2_rn(gvix, ux)));
      // edge B (j receives from i, unit negated): dxB1 = gsj, dxB2 = -(gvj.u) =: -nB2, dxB0 = gvj.vi
      const float2 nB2 = __ffma2_rn(gvjz, uz, __ffma2_rn(gvjy, uy, __fmul2_rn(gvjx, ux)));
      // dd = sum_k (dwA_k + dwB_k) q_k
      float2 t1 = __ffma2_rn(gsi, pj1, __fmul2_rn(gsj, pi1));                 // dxA1*pj1 + dxB1*pi1
      float2 t2 = __ffma2_rn(dxA2, pj2, neg2(__fmul2_rn(nB2, pi2)));          // dxA2*pj2 + dxB2*pi2
      float2 dd = __ffma2_rn(t2, q2, __fmul2_rn(t1, q1));
      if (!FIRST) {
        const float2 vjx = ld2(sj + O_V), vjy = ld2(sj + O_V + MSG_FC), vjz = ld2(sj + O_V + 2 * MSG_FC);
        const float2 dxA0 = __ffma2_rn(gviz, vjz, __ffma2_rn(gviy, vjy, __fmul2_rn(gvix, vjx)));
        const float2 dxB0 = __ffma2_rn(gvjz, viz, __ffma2_rn(gvjy, viy, __fmul2_rn(gvjx, vix)));
        const float2 t0 = __ffma2_rn(dxA0, pj0, __fmul2_rn(dxB0, pi0));
        dd = __ffma2_rn(t0, q0, dd);
        dp0 = __ffma2_rn(dxB0, w0, dp0);
        dp1 = __ffma2_rn(gsj, w1, dp1);
        dp2n = __ffma2_rn(nB2, w2, dp2n);
        const float2 tv = __fmul2_rn(pi0, w0);
        dvx = __ffma2_rn(tv, gvjx, dvx); dvy = __ffma2_rn(tv, gvjy, dvy); dvz = __ffma2_rn(tv, gvjz, dvz);
      }
      // unit-vector chain: delta = gvi*(pj2 w2) - gvj*(pi2 w2); project out the radial part
      const float2 ta = __fmul2_rn(pj2, w2), tbn = neg2(__fmul2_rn(pi2, w2));
      const float2 ex = __ffma2_rn(gvix, ta, __fmul2_rn(gvjx, tbn));
      const float2 ey = __ffma2_rn(gviy, ta, __fmul2_rn(gvjy, tbn));
      const float2 ez = __ffma2_rn(gviz, ta, __fmul2_rn(gvjz, tbn));
      const float2 proj = __ffma2_rn(ez, uz, __ffma2_rn(ey, uy, __fmul2_rn(ex, ux)));
      const float2 invd = dup2(1.0f / g.w);
      const float2 c = __ffma2_rn(neg2(proj), invd, dd);   // dd - proj/d  (multiplies u)
      gnx = __ffma2_rn(c, ux, __ffma2_rn(ex, invd, gnx));
      gny = __ffma2_rn(c, uy, __ffma2_rn(ey, invd, gny));
      gnz = __ffma2_rn(c, uz, __ffma2_rn(ez, invd, gnz));
    }
    if (!FIRST) {
      float* dpo = dphi + (long long)il * F3 + f0;
      *reinterpret_cast<float2*>(dpo) = dp0;
      *reinterpret_cast<float2*>(dpo + F) = dp1;
      *reinterpret_cast<float2*>(dpo + 2 * F) = neg2(dp2n);
      float* dvo = dv_in + (long long)il * 3 * F + f0;
      *reinterpret_cast<float2*>(dvo) = __fadd2_rn(gvix, dvx);
      *reinterpret_cast<float2*>(dvo + F) = __fadd2_rn(gviy, dvy);
      *reinterpret_cast<float2*>(dvo + 2 * F) = __fadd2_rn(gviz, dvz);
    }
    float gx = warp_sum(gnx.x + gnx.y), gy = warp_sum(gny.x + gny.y), gz = warp_sum(gnz.x + gnz.y);
    if (lane == 0) {
      float* gp = gradp + (((long long)m * 2 + h) * n_atoms + i) * 3;
      gp[0] = -gx; gp[1] = -gy; gp[2] = -gz;
    }
  }
}

// grad[m][a][c] += gradp[m][0][a][c] + gradp[m][1][a][c]   (fixed order)
__global__ void grad_accum_kernel(const float* __restrict__ gradp, int n3, float* __restrict__ grad) {
  const int m = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n3) return;
  const float* p = gradp + (long long)m * 2 * n3;
  grad[(long long)m * n3 + idx] += p[idx] + p[n3 + idx];
}
