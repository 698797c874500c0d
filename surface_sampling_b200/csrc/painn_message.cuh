// PaiNN message passing, forward and backward — shared-memory staged, FFMA2-packed (sm_100).
// Included by painn.cu inside its anonymous namespace.
//
// Work decomposition: CTA = (structure [, atom-chunk]) x feature-half (64 of the 128 features) x model.
// The CTA first stages, for its 64 features, the per-atom rows it will gather from EVERY atom of
// the structure into shared memory (phi, v [, ds, dv]); after one barrier each warp owns one receiver
// atom at a time and lane l owns the adjacent feature pair (2l, 2l+1) of the chunk.  All arithmetic on
// the pair is issued as packed fp32x2 instructions (fma.rn.f32x2 -> SASS FFMA2), which on B200 sustain
// 66 TFLOP/s vs 42 for scalar FFMA (profiles/microbench/ffma2.cu).
//
// Two passes per layer, because the two kinds of edges want opposite register budgets:
//   *_memo*  : edges whose radial filter is memoised (frozen pairs, see FilterCacheView).  The filter
//              rows are simply loaded, so a thread needs 60-120 registers and the CTA runs 16-26 warps.
//              *_group kernels: G "canonical" structures per CTA walk the framework's own edge lists and
//              share every filter row (the normal case); *_memo / *_memo_state: one structure per CTA with
//              the rows in a per-warp cp.async ring (fallback); message_bwd_memo: full-gradient variant.
//   *_v2     : all remaining ("direct") edges.  The filter w(d) = Wd.(rbf*env) + bd*env is evaluated in
//              60 register pairs per lane (120 FFMA2 per edge in the backward), 8 warps per CTA; per-edge
//              records stream through a private cp.async ring per warp.  Frozen receivers in constrained
//              mode only need the state terms (no dw/dd).
// The second pass accumulates onto the first (fixed order: memoised edges, then direct edges).
// Rows are handed to warps most-expensive-first (row_order_kernel + a shared-memory counter).
// Determinism: a lane walks its receiver's edge lists serially; there are no atomics on data.
#pragma once

constexpr int MSG_FC = 64;        // features per CTA
constexpr int MSG_THREADS = 256;  // direct pass: 8 warps
constexpr int MSG_WARPS = MSG_THREADS / 32;
constexpr int MSG_STAGES = 3;     // cp.async ring depth per warp
constexpr int MSG_PIPE_BYTES = MSG_WARPS * MSG_STAGES * REC * 4;
constexpr int MSG_STAGES_FWD = 4;    // direct forward pass: two edges per iteration, two more in flight
constexpr int MSG_PIPE_BYTES_FWD = MSG_WARPS * MSG_STAGES_FWD * REC * 4;
constexpr int MEMO_THREADS_FWD = 768;   // memo pass: 24 warps share one staged structure
constexpr int MEMO_THREADS_BWD = 384;   // full-gradient memo backward: 12 warps, (w,q) rows through a 3-stage ring

__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
// *p += x without the load latency: one fire-and-forget reduction per address and launch -- a single IEEE addition onto
// the stored value, i.e. the bits of load + add + store (launches are stream-ordered, so the result stays deterministic)
__device__ __forceinline__ void red_add2(float* p, float2 x) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x.x), "f"(x.y) : "memory");
}

// smem floats per staged atom
template <bool FIRST> struct MsgFwdLayout { static constexpr int PER = FIRST ? 3 * MSG_FC : 6 * MSG_FC; };
template <bool FIRST> struct MsgBwdLayout { static constexpr int PER = FIRST ? 7 * MSG_FC : 10 * MSG_FC; };

// copy `rows` rows of MSG_FC floats: src row r at src + r*src_stride, dst at dst + r*MSG_FC (per atom a).
// cp.async (LDGSTS): no register round trip, so every 16-byte piece of the ~100-170 KB staged per CTA is in
// flight at once instead of one load per thread per memory latency.  Completed by stage_wait().
__device__ __forceinline__ void stage_rows(float* __restrict__ dst_atom0, int per, int dst_off,
                                           const float* __restrict__ src, long long src_atom_stride, int src_row_stride,
                                           int rows, int n, int tid, int nthreads) {
  constexpr int Q = MSG_FC / 4;  // float4 per row
  const int per_atom = rows * Q;
  for (int idx = tid; idx < n * per_atom; idx += nthreads) {
    const int a = idx / per_atom, r = (idx % per_atom) / Q, c4 = idx % Q;
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_atom0 + a * per + dst_off + r * MSG_FC + c4 * 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src + a * src_atom_stride + r * src_row_stride + c4 * 4));
  }
}
// all staged rows of this thread have landed (call before the __syncthreads that publishes them)
__device__ __forceinline__ void stage_wait() {
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
}

// one 16-byte cp.async per lane for the first n16*16 bytes of a record, then commit (always commits,
// so group accounting stays uniform even past the end of the row)
__device__ __forceinline__ void prefetch_record(float* dst, const float* src, int lane, int n16, bool live) {
  if (live && lane < n16) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + lane * 4);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src + lane * 4));
  }
  asm volatile("cp.async.commit_group;\n" ::);
}
__device__ __forceinline__ void wait_record() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(MSG_STAGES - 2));
  __syncwarp();
}

// Sender windows.  A CTA of the direct pass stages the rows of at most `cap` atoms; a structure with n > cap atoms is
// covered by W = ceil(n / cap) launches ("windows" of ceil(n / W) consecutive atoms): launch `win` handles, for every
// receiver, exactly the edges whose SENDER lies in its window and accumulates onto the earlier launches.  Direct
// records of a row are sorted by sender (CSR order survives the compaction), so a window is a contiguous slice of the
// row, found by binary search.  W depends on the structure's own atom count only, never on the rest of the batch, and
// windows are applied in ascending order: results stay bitwise batch-invariant.  n <= cap is the old single launch.
// Row table of a direct-pass CTA: (local atom, first record, record count) of the structure's rows in LPT order, read
// once by all threads while the rows are being staged -- a warp that grabs a row then starts its record ring at once
// instead of walking three dependent global loads (order -> rowptr -> nvalid) with 8 warps per SM to hide them.
constexpr int MSG_MAXROWS = 256;   // larger structures read the tail of the table from global memory
struct RowTable { int il[MSG_MAXROWS], e0[MSG_MAXROWS], ne[MSG_MAXROWS]; };
__device__ __forceinline__ void row_table_fill(RowTable& rt, const int32_t* __restrict__ order, const int32_t* __restrict__ rowptr,
                                               const int32_t* __restrict__ nvalid, int a0, int n, int tid, int nthreads) {
  for (int t = tid; t < n && t < MSG_MAXROWS; t += nthreads) {
    const int il = __ldg(order + a0 + t);
    rt.il[t] = il; rt.e0[t] = __ldg(rowptr + a0 + il); rt.ne[t] = __ldg(nvalid + a0 + il);
  }
}
__device__ __forceinline__ void row_table_get(const RowTable& rt, const int32_t* __restrict__ order, const int32_t* __restrict__ rowptr,
                                              const int32_t* __restrict__ nvalid, int a0, int t, int& il, int& e0, int& ne) {
  if (t < MSG_MAXROWS) { il = rt.il[t]; e0 = rt.e0[t]; ne = rt.ne[t]; }
  else { il = __ldg(order + a0 + t); e0 = __ldg(rowptr + a0 + il); ne = __ldg(nvalid + a0 + il); }
}

struct SenderWindow { int lo, hi; bool live; };
__device__ __forceinline__ SenderWindow sender_window(int n, int cap, int win) {
  const int W = (n + cap - 1) / cap;
  const int wsz = W > 0 ? (n + W - 1) / W : 0;
  SenderWindow w;
  w.live = win < W;
  w.lo = win * wsz;
  w.hi = min(n, w.lo + wsz);
  return w;
}
// [e_lo, e_hi) = the direct records of a row whose local sender index lies in [lo, hi).  Records are sorted by sender,
// so the bounds are counts: every lane reads the sender of one record (32 independent loads in flight, one memory
// latency per 32 edges instead of a chain of dependent probes) and the warp counts by ballot.  Warp-uniform result.
__device__ __forceinline__ void window_slice(const float* __restrict__ rec0, int ne, int a0, int lo, int hi, int lane,
                                             int& e_lo, int& e_hi) {
  e_lo = 0; e_hi = 0;
  for (int base = 0; base < ne; base += 32) {
    const int e = base + lane;
    const int jl = e < ne ? __float_as_int(__ldg(rec0 + (long long)e * REC + REC_EJ)) - a0 : 0x7fffffff;
    e_lo += __popc(__ballot_sync(0xffffffffu, jl < lo));
    e_hi += __popc(__ballot_sync(0xffffffffu, jl < hi));
  }
}

// The same for every row a CTA owns (table slots t = ch, ch + n_chunks, ...), done once in the prologue while the sender
// rows are being staged: four rows per warp and step, their sender loads all in flight before the first ballot, so the
// CTA pays two or three memory latencies here instead of one per row inside the row loop.  The table then holds the
// window's slice of each row; slots beyond the table (n > MSG_MAXROWS) are sliced in the row loop as before.
__device__ __forceinline__ void row_table_slice(RowTable& rt, const float* __restrict__ erec, int a0, int n, int ch, int n_chunks,
                                                int lo, int hi, int warp, int lane, int nwarps) {
  const int nt = min(n, MSG_MAXROWS);
  for (int k = 4 * warp; ch + n_chunks * k < nt; k += 4 * nwarps) {
    int e0[4], ne[4], jl[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = ch + n_chunks * (k + u);
      e0[u] = t < nt ? rt.e0[t] : 0;
      ne[u] = t < nt ? rt.ne[t] : 0;
#pragma unroll
      for (int sg = 0; sg < 2; ++sg) {
        const int e = 32 * sg + lane;
        jl[u][sg] = e < ne[u] ? __float_as_int(__ldg(erec + (long long)(e0[u] + e) * REC + REC_EJ)) - a0 : 0x7fffffff;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = ch + n_chunks * (k + u);
      int e_lo = 0, e_hi = 0;
#pragma unroll
      for (int sg = 0; sg < 2; ++sg) {
        e_lo += __popc(__ballot_sync(0xffffffffu, jl[u][sg] < lo));
        e_hi += __popc(__ballot_sync(0xffffffffu, jl[u][sg] < hi));
      }
      for (int base = 64; base < ne[u]; base += 32) {     // rows with more than 64 direct records (warp-uniform)
        const int e = base + lane;
        const int j = e < ne[u] ? __float_as_int(__ldg(erec + (long long)(e0[u] + e) * REC + REC_EJ)) - a0 : 0x7fffffff;
        e_lo += __popc(__ballot_sync(0xffffffffu, j < lo));
        e_hi += __popc(__ballot_sync(0xffffffffu, j < hi));
      }
      __syncwarp();     // every lane has read the slot (orders the reads above against the write below)
      if (lane == 0 && t < nt) { rt.e0[t] = e0[u] + e_lo; rt.ne[t] = e_hi - e_lo; }
    }
  }
}

// Rows (receivers) are handed to warps dynamically, most expensive first (row_order_kernel sorts each
// structure's rows by edge count): longest-processing-time-first scheduling.  A row is always computed
// whole by one warp, so the result does not depend on which warp takes it.
__device__ __forceinline__ int next_row(int* ctr, int lane) {
  int t = 0;
  if (lane == 0) t = atomicAdd(ctr, 1);
  return __shfl_sync(0xffffffffu, t, 0);
}

// is structure b handled by a group kernel (G consecutive canonical structures per CTA)?
// (A memoised edge joins two frozen FRAMEWORK atoms, so a group kernel only ever stages the first n0 rows of each
// structure: whether a group fits shared memory is decided on the host from n0 alone -- `max_atoms` >= G*n0 -- and does
// not change when the chains collect adsorbates.)
__device__ __forceinline__ bool in_canonical_group(const int32_t* __restrict__ canonical, const int32_t* __restrict__ atom_ptr,
                                                   int b, int n_struct, int G, int max_atoms) {
  (void)atom_ptr;
  if (!canonical || max_atoms <= 0) return false;
  const int q = (b / G) * G;
  if (q + G > n_struct) return false;
  for (int s = 0; s < G; ++s)
    if (!__ldg(canonical + q + s)) return false;
  return true;
}

// order[a0 + rank] = local index of the row with that rank (cost descending, index ascending on ties).
// With a framework (n0 > 0) also: canonical[b] = 1 iff every framework row of structure b carries exactly the
// framework's own memoised edges (same count => same set and order, both are in CSR order) -- then the pair
// kernels below may walk the framework's lists instead of the structure's.
__global__ void __launch_bounds__(128) row_order_kernel(const int32_t* __restrict__ atom_ptr, const int32_t* __restrict__ cost_a,
                                                        const int32_t* __restrict__ cost_b, int32_t* __restrict__ order_a,
                                                        int32_t* __restrict__ order_b, int n0,
                                                        const int32_t* __restrict__ nmemo0, int32_t* __restrict__ canonical) {
  extern __shared__ int32_t costs[];   // [2][n]
  __shared__ int mismatch;
  const int b = blockIdx.x;
  const int a0 = atom_ptr[b], n = atom_ptr[b + 1] - a0;
  if (threadIdx.x == 0) mismatch = 0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) { costs[k] = cost_a[a0 + k]; costs[n + k] = cost_b[a0 + k]; }
  __syncthreads();
  for (int k = threadIdx.x; k < 2 * n; k += blockDim.x) {
    const int which = k >= n, il = which ? k - n : k;
    const int32_t* c = costs + which * n;
    const int mine = c[il];
    int rank = 0;
    for (int o = 0; o < n; ++o) rank += (c[o] > mine) || (c[o] == mine && o < il);
    (which ? order_b : order_a)[a0 + rank] = il;
    if (which && n0 > 0 && il < n0 && mine != nmemo0[il]) mismatch = 1;   // benign race: all writers store 1
  }
  if (canonical) {
    __syncthreads();
    if (threadIdx.x == 0) canonical[b] = (n0 > 0 && n >= n0 && !mismatch) ? 1 : 0;
  }
}

// Walk the memoised edges of one receiver.  Lane l loads record base+l (32 records per sweep, coalesced) and the
// per-edge fields are broadcast by shuffle.  The filter rows (w0,w1,w2 of this feature half: 768 B per edge) travel
// through a private cp.async ring per warp, MEMO_STAGES deep.  (Register prefetching does not survive ptxas: every
// LDG of the kernel is tracked by one scoreboard, so waiting for the oldest row also waits for the newest;
// profiles/r2_notes.md section 1.)
// body(g, j, inv_d, rows) with rows[k * MSG_FC] = w_k pair of this lane.
constexpr int MEMO_STAGES = 4;
constexpr int MEMO_STAGE_FLOATS = 3 * MSG_FC;
constexpr int MEMO_RING_BYTES_PER_WARP = MEMO_STAGES * MEMO_STAGE_FLOATS * 4;
template <bool ROW0, typename Body>
__device__ __forceinline__ void memo_walk_ring(const float4* __restrict__ mr, int ne, int lane,
                                               const float* __restrict__ wbase, float* __restrict__ ring, Body body) {
  const int grp = lane >> 4, c4 = (lane & 15) * 4;
  for (int base = 0; base < ne; base += 32) {
    const int cnt = min(32, ne - base);
    float4 gl = make_float4(0.f, 0.f, 0.f, 1.f), jl = make_float4(0.f, 0.f, 1.f, 0.f);
    if (lane < cnt) { gl = __ldg(mr + 2 * (base + lane)); jl = __ldg(mr + 2 * (base + lane) + 1); }
    auto issue = [&](int e) {   // e is warp-uniform
      if (e < cnt) {
        const int slot = __shfl_sync(0xffffffffu, __float_as_int(jl.y), e);
        const float* src = wbase + (long long)slot * F3 + c4;
        float* dst = ring + (e % MEMO_STAGES) * MEMO_STAGE_FLOATS + c4;
        {
          const unsigned d = (unsigned)__cvta_generic_to_shared(dst + (1 + grp) * MSG_FC);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src + (1 + grp) * F));
        }
        if (ROW0 && grp == 0) {
          const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
        }
      }
      asm volatile("cp.async.commit_group;\n" ::);
    };
    __syncwarp();   // every lane is done reading the ring (previous sweep / receiver)
#pragma unroll
    for (int p = 0; p < MEMO_STAGES - 1; ++p) issue(p);
    for (int e = 0; e < cnt; ++e) {
      asm volatile("cp.async.wait_group %0;\n" ::"n"(MEMO_STAGES - 2));
      __syncwarp();
      issue(e + MEMO_STAGES - 1);
      const float4 g = make_float4(__shfl_sync(0xffffffffu, gl.x, e), __shfl_sync(0xffffffffu, gl.y, e),
                                   __shfl_sync(0xffffffffu, gl.z, e), __shfl_sync(0xffffffffu, gl.w, e));
      const int j = __shfl_sync(0xffffffffu, __float_as_int(jl.x), e);
      const float inv_d = __shfl_sync(0xffffffffu, jl.z, e);
      body(g, j, inv_d, ring + (e % MEMO_STAGES) * MEMO_STAGE_FLOATS + 2 * lane);
    }
  }
}

// Six-row variant for the full-gradient backward: w0,w1,w2 and their d-derivatives q0,q1,q2 (1536 B per edge).
constexpr int MEMO6_STAGES = 3;
constexpr int MEMO6_STAGE_FLOATS = 6 * MSG_FC;
template <bool ROW0, typename Body>
__device__ __forceinline__ void memo_walk_ring6(const float4* __restrict__ mr, int ne, int lane, const float* __restrict__ wbase,
                                                const float* __restrict__ qbase, float* __restrict__ ring, Body body) {
  const int grp = lane >> 4, c4 = (lane & 15) * 4;
  auto cp16 = [](float* dst, const float* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
  };
  for (int base = 0; base < ne; base += 32) {
    const int cnt = min(32, ne - base);
    float4 gl = make_float4(0.f, 0.f, 0.f, 1.f), jl = make_float4(0.f, 0.f, 1.f, 0.f);
    if (lane < cnt) { gl = __ldg(mr + 2 * (base + lane)); jl = __ldg(mr + 2 * (base + lane) + 1); }
    auto issue = [&](int e) {   // e is warp-uniform
      if (e < cnt) {
        const long long so = (long long)__shfl_sync(0xffffffffu, __float_as_int(jl.y), e) * F3 + c4;
        float* dst = ring + (e % MEMO6_STAGES) * MEMO6_STAGE_FLOATS + c4;
        cp16(dst + (1 + grp) * MSG_FC, wbase + so + (1 + grp) * F);
        cp16(dst + (4 + grp) * MSG_FC, qbase + so + (1 + grp) * F);
        if (ROW0) cp16(dst + 3 * grp * MSG_FC, (grp ? qbase : wbase) + so);
      }
      asm volatile("cp.async.commit_group;\n" ::);
    };
    __syncwarp();
#pragma unroll
    for (int p = 0; p < MEMO6_STAGES - 1; ++p) issue(p);
    for (int e = 0; e < cnt; ++e) {
      asm volatile("cp.async.wait_group %0;\n" ::"n"(MEMO6_STAGES - 2));
      __syncwarp();
      issue(e + MEMO6_STAGES - 1);
      const float4 g = make_float4(__shfl_sync(0xffffffffu, gl.x, e), __shfl_sync(0xffffffffu, gl.y, e),
                                   __shfl_sync(0xffffffffu, gl.z, e), __shfl_sync(0xffffffffu, gl.w, e));
      const int j = __shfl_sync(0xffffffffu, __float_as_int(jl.x), e);
      const float inv_d = __shfl_sync(0xffffffffu, jl.z, e);
      body(g, j, inv_d, ring + (e % MEMO6_STAGES) * MEMO6_STAGE_FLOATS + 2 * lane);
    }
  }
}

// ============================================================================================
// forward
// ============================================================================================
// shared tail of both forward passes: x = phi_j * w ; ds += x1 ; dv += x2*u + x0*v_j
template <bool FIRST>
__device__ __forceinline__ void fwd_edge(const float4 g, const float* __restrict__ sj, float2 w0, float2 w1, float2 w2,
                                         float2& ds, float2& dvx, float2& dvy, float2& dvz) {
  const float2 p1 = ld2(sj + MSG_FC), p2 = ld2(sj + 2 * MSG_FC);
  const float2 x1 = __fmul2_rn(p1, w1), x2 = __fmul2_rn(p2, w2);
  ds = __fadd2_rn(ds, x1);
  dvx = __ffma2_rn(x2, dup2(g.x), dvx);
  dvy = __ffma2_rn(x2, dup2(g.y), dvy);
  dvz = __ffma2_rn(x2, dup2(g.z), dvz);
  if (!FIRST) {
    const float2 x0 = __fmul2_rn(ld2(sj), w0);
    const float2 vx = ld2(sj + 3 * MSG_FC), vy = ld2(sj + 4 * MSG_FC), vz = ld2(sj + 5 * MSG_FC);
    dvx = __ffma2_rn(x0, vx, dvx); dvy = __ffma2_rn(x0, vy, dvy); dvz = __ffma2_rn(x0, vz, dvz);
  }
}

// memo pass: cat[:, :128] = s_in + sum_memo x1 ; v_mid = v_in + sum_memo (...)   (always writes)
template <bool FIRST>
__global__ void __launch_bounds__(MEMO_THREADS_FWD, 1) message_fwd_memo(
    int layer, int n_atoms, const int32_t* __restrict__ atom_ptr, const int32_t* __restrict__ rowptr,
    const int32_t* __restrict__ order, const int32_t* __restrict__ nmemo, const float* __restrict__ mrec, FilterCacheView fc,
    const float* __restrict__ phi, const float* __restrict__ s_in, const float* __restrict__ v_in,
    float* __restrict__ cat, float* __restrict__ v_mid, const int32_t* __restrict__ canonical, int n_struct, int group,
    int group_atoms) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int row_ctr;
  if (in_canonical_group(canonical, atom_ptr, blockIdx.x, n_struct, group, group_atoms)) return;   // done by message_fwd_memo_group
  if (threadIdx.x == 0) row_ctr = 0;
  constexpr int PER = MsgFwdLayout<FIRST>::PER;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x, h = blockIdx.y, m = blockIdx.z;
  const int a0 = __ldg(atom_ptr + b), n = __ldg(atom_ptr + b + 1) - a0;
  const long long mA = (long long)m * n_atoms;
  phi += (mA + a0) * F3 + h * MSG_FC;
  s_in += (mA + a0) * F;
  cat += (mA + a0) * 2 * F;
  v_mid += (mA + a0) * 3 * F;
  if (!FIRST) v_in += (mA + a0) * 3 * F + h * MSG_FC;
  const int ns = min(n, fc.n0);   // memoised edges join framework atoms: only their rows are ever gathered
  float* ring = smem + (size_t)ns * PER + warp * (MEMO_STAGES * MEMO_STAGE_FLOATS);
  stage_rows(smem, PER, 0, phi, F3, F, 3, ns, tid, MEMO_THREADS_FWD);
  if (!FIRST) stage_rows(smem, PER, 3 * MSG_FC, v_in, 3 * F, F, 3, ns, tid, MEMO_THREADS_FWD);
  const int f0 = h * MSG_FC + 2 * lane;
  const float* __restrict__ wbase = fc.wc + (long long)(m * NCONV + layer) * fc.nslots_cap * F3 + h * MSG_FC;
  stage_wait();
  __syncthreads();
  for (int t = next_row(&row_ctr, lane); t < n; t = next_row(&row_ctr, lane)) {
    const int il = __ldg(order + a0 + t);
    const int i = a0 + il;
    const float4* mr = reinterpret_cast<const float4*>(mrec + (long long)__ldg(rowptr + i) * MREC);
    const int ne = __ldg(nmemo + i);
    float2 ds = dup2(0.f), dvx = dup2(0.f), dvy = dup2(0.f), dvz = dup2(0.f);
    memo_walk_ring<!FIRST>(mr, ne, lane, wbase, ring, [&](const float4 g, int j, float, const float* r) {
      fwd_edge<FIRST>(g, smem + (j - a0) * PER + 2 * lane, FIRST ? dup2(0.f) : ld2(r), ld2(r + MSG_FC),
                      ld2(r + 2 * MSG_FC), ds, dvx, dvy, dvz);
    });
    const float2 s0 = ld2(s_in + (long long)il * F + f0);
    *reinterpret_cast<float2*>(cat + (long long)il * 2 * F + f0) = __fadd2_rn(s0, ds);
    if (!FIRST) {
      if (il < ns) {
        const float* si = smem + il * PER + 2 * lane;
        dvx = __fadd2_rn(dvx, ld2(si + 3 * MSG_FC));
        dvy = __fadd2_rn(dvy, ld2(si + 4 * MSG_FC));
        dvz = __fadd2_rn(dvz, ld2(si + 5 * MSG_FC));
      } else {   // adsorbate row (not staged)
        const float* vg = v_in + (long long)il * 3 * F + 2 * lane;
        dvx = __fadd2_rn(dvx, ldg2(vg)); dvy = __fadd2_rn(dvy, ldg2(vg + F)); dvz = __fadd2_rn(dvz, ldg2(vg + 2 * F));
      }
    }
    float* vo = v_mid + (long long)il * 3 * F + f0;
    *reinterpret_cast<float2*>(vo) = dvx;
    *reinterpret_cast<float2*>(vo + F) = dvy;
    *reinterpret_cast<float2*>(vo + 2 * F) = dvz;
  }
}

// direct pass.  accum = 0: writes s_in + ds / v_in + dv ; accum = 1: adds onto the memo pass' output
// first layer: 40 weight pairs and 3 staged rows per atom -> two CTAs per SM fit (registers and shared memory)
template <bool FIRST>
__global__ void __launch_bounds__(MSG_THREADS, FIRST ? 2 : 1) message_fwd_v2(
    const float* __restrict__ weights, int layer, int n_atoms, const int32_t* __restrict__ atom_ptr, int n_chunks,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ order, const int32_t* __restrict__ nvalid,
    const float* __restrict__ erec,
    const float* __restrict__ phi, const float* __restrict__ s_in, const float* __restrict__ v_in,
    float* __restrict__ cat, float* __restrict__ v_mid, int accum, int cap_atoms, int win) {
  extern __shared__ __align__(16) float smem_all[];
  __shared__ int row_ctr;
  if (threadIdx.x == 0) row_ctr = 0;
  constexpr int PER = MsgFwdLayout<FIRST>::PER;
  constexpr int N16 = (REC_RE + 44) / 4;  // 13 x 16 B: geometry + rbf rows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* ring = smem_all + warp * MSG_STAGES_FWD * REC;
  float* smem = smem_all + MSG_WARPS * MSG_STAGES_FWD * REC;
  // grid = (halves x models, structures x chunks): the CTAs that stream the SAME edge records are neighbours in launch
  // order and run together, so the records are read from DRAM once and served to the other five from L2
  const int b = blockIdx.y / n_chunks, ch = blockIdx.y % n_chunks;
  const int h = blockIdx.x % (F / MSG_FC), m = blockIdx.x / (F / MSG_FC);
  const int a0 = __ldg(atom_ptr + b), n = __ldg(atom_ptr + b + 1) - a0;
  const long long mA = (long long)m * n_atoms;
  const SenderWindow wn = sender_window(n, cap_atoms, win);
  if (!wn.live) return;                          // this structure needs fewer windows than the largest of the batch
  const bool whole = wn.lo == 0 && wn.hi == n;   // single window: the row's whole list
  if (win > 0) accum = 1;
  phi += (mA + a0) * F3 + h * MSG_FC;
  s_in += (mA + a0) * F;
  cat += (mA + a0) * 2 * F;
  v_mid += (mA + a0) * 3 * F;
  if (!FIRST) v_in += (mA + a0) * 3 * F + h * MSG_FC;

  stage_rows(smem, PER, 0, phi + (long long)wn.lo * F3, F3, F, 3, wn.hi - wn.lo, tid, MSG_THREADS);
  if (!FIRST) stage_rows(smem, PER, 3 * MSG_FC, v_in + (long long)wn.lo * 3 * F, 3 * F, F, 3, wn.hi - wn.lo, tid, MSG_THREADS);
  const int sbase = a0 + wn.lo;                  // global index of the first staged atom
  __shared__ RowTable rt;
  row_table_fill(rt, order, rowptr, nvalid, a0, n, tid, MSG_THREADS);

  const int f0 = h * MSG_FC + 2 * lane;  // first feature of this lane's pair
  const float* __restrict__ wl = weights + (long long)m * W_STRIDE + W_LAYER0 + (long long)layer * L_SIZE;
  float2 wd0[NRBF], wd1[NRBF], wd2[NRBF];
#pragma unroll
  for (int q = 0; q < NRBF; ++q) {
    wd0[q] = ld2(wl + L_WDT + q * F3 + f0);
    wd1[q] = ld2(wl + L_WDT + q * F3 + F + f0);
    wd2[q] = ld2(wl + L_WDT + q * F3 + 2 * F + f0);
  }
  const float2 bd0 = ld2(wl + L_BD + f0), bd1 = ld2(wl + L_BD + F + f0), bd2 = ld2(wl + L_BD + 2 * F + f0);
  if (!whole) {
    __syncthreads();                  // row table complete
    row_table_slice(rt, erec, a0, n, ch, n_chunks, wn.lo, wn.hi, warp, lane, MSG_THREADS / 32);
  }
  stage_wait();
  __syncthreads();

  for (int t = ch + n_chunks * next_row(&row_ctr, lane); t < n; t = ch + n_chunks * next_row(&row_ctr, lane)) {
    int il, e0, ne;
    row_table_get(rt, order, rowptr, nvalid, a0, t, il, e0, ne);
    const float* rec0 = erec + (long long)e0 * REC;
    if (!whole && t >= MSG_MAXROWS) { // the slice of this row whose senders lie in the window (table rows: done above)
      int e_lo, e_hi;
      window_slice(rec0, ne, a0, wn.lo, wn.hi, lane, e_lo, e_hi);
      rec0 += (long long)e_lo * REC;
      ne = e_hi - e_lo;
    }
    if (accum && ne == 0) continue;   // nothing to add to the memo pass' / earlier windows' result
    float2 ds = dup2(0.f), dvx = dup2(0.f), dvy = dup2(0.f), dvz = dup2(0.f);
    // two edges per iteration: their filter evaluations (2 x 60 independent FFMA2 on the same weight registers)
    // and gathers interleave, which hides the shared-memory latencies that 8 warps per SM cannot
    __syncwarp();
#pragma unroll
    for (int s = 0; s < MSG_STAGES_FWD; ++s) prefetch_record(ring + s * REC, rec0 + (long long)s * REC, lane, N16, s < ne);
    for (int e = 0; e < ne; e += 2) {
      asm volatile("cp.async.wait_group %0;\n" ::"n"(MSG_STAGES_FWD - 2));   // records e and e+1 have landed
      __syncwarp();
      const float* reca = ring + (e % MSG_STAGES_FWD) * REC;
      const float* recb = ring + ((e + 1) % MSG_STAGES_FWD) * REC;
      const float4 ga = *reinterpret_cast<const float4*>(reca), gb = *reinterpret_cast<const float4*>(recb);
      const float4* ra4 = reinterpret_cast<const float4*>(reca + REC_RE);
      const float4* rb4 = reinterpret_cast<const float4*>(recb + REC_RE);
      const float4 eva = ra4[10], evb = rb4[10];  // (env,env,denv,denv)
      const float2 enva = make_float2(eva.x, eva.y), envb = make_float2(evb.x, evb.y);
      float2 w0a = __fmul2_rn(bd0, enva), w1a = __fmul2_rn(bd1, enva), w2a = __fmul2_rn(bd2, enva);
      float2 w0b = __fmul2_rn(bd0, envb), w1b = __fmul2_rn(bd1, envb), w2b = __fmul2_rn(bd2, envb);
#pragma unroll
      for (int q = 0; q < NRBF / 2; ++q) {
        const float4 ta = ra4[q], tb = rb4[q];
        const float2 a0r = make_float2(ta.x, ta.y), a1r = make_float2(ta.z, ta.w);
        const float2 b0r = make_float2(tb.x, tb.y), b1r = make_float2(tb.z, tb.w);
        w0a = __ffma2_rn(wd0[2 * q], a0r, w0a); w1a = __ffma2_rn(wd1[2 * q], a0r, w1a); w2a = __ffma2_rn(wd2[2 * q], a0r, w2a);
        w0b = __ffma2_rn(wd0[2 * q], b0r, w0b); w1b = __ffma2_rn(wd1[2 * q], b0r, w1b); w2b = __ffma2_rn(wd2[2 * q], b0r, w2b);
        w0a = __ffma2_rn(wd0[2 * q + 1], a1r, w0a); w1a = __ffma2_rn(wd1[2 * q + 1], a1r, w1a);
        w2a = __ffma2_rn(wd2[2 * q + 1], a1r, w2a);
        w0b = __ffma2_rn(wd0[2 * q + 1], b1r, w0b); w1b = __ffma2_rn(wd1[2 * q + 1], b1r, w1b);
        w2b = __ffma2_rn(wd2[2 * q + 1], b1r, w2b);
      }
      const int ja = __float_as_int(reca[REC_EJ]), jb = __float_as_int(recb[REC_EJ]);
      fwd_edge<FIRST>(ga, smem + (ja - sbase) * PER + 2 * lane, w0a, w1a, w2a, ds, dvx, dvy, dvz);
      if (e + 1 < ne) fwd_edge<FIRST>(gb, smem + (jb - sbase) * PER + 2 * lane, w0b, w1b, w2b, ds, dvx, dvy, dvz);
      __syncwarp();   // both stages are free again
      prefetch_record(ring + (e % MSG_STAGES_FWD) * REC, rec0 + (long long)(e + MSG_STAGES_FWD) * REC, lane, N16,
                      e + MSG_STAGES_FWD < ne);
      prefetch_record(ring + ((e + 1) % MSG_STAGES_FWD) * REC, rec0 + (long long)(e + 1 + MSG_STAGES_FWD) * REC, lane, N16,
                      e + 1 + MSG_STAGES_FWD < ne);
    }
    float* so = cat + (long long)il * 2 * F + f0;
    float* vo = v_mid + (long long)il * 3 * F + f0;
    if (accum) {
      red_add2(so, ds);
      red_add2(vo, dvx);
      red_add2(vo + F, dvy);
      red_add2(vo + 2 * F, dvz);
    } else {
      *reinterpret_cast<float2*>(so) = __fadd2_rn(ld2(s_in + (long long)il * F + f0), ds);
      if (!FIRST) {
        if (il >= wn.lo && il < wn.hi) {
          const float* si = smem + (il - wn.lo) * PER + 2 * lane;
          dvx = __fadd2_rn(dvx, ld2(si + 3 * MSG_FC));
          dvy = __fadd2_rn(dvy, ld2(si + 4 * MSG_FC));
          dvz = __fadd2_rn(dvz, ld2(si + 5 * MSG_FC));
        } else {   // own row outside the staged window
          const float* vg = v_in + (long long)il * 3 * F + 2 * lane;
          dvx = __fadd2_rn(dvx, ldg2(vg)); dvy = __fadd2_rn(dvy, ldg2(vg + F)); dvz = __fadd2_rn(dvz, ldg2(vg + 2 * F));
        }
      }
      *reinterpret_cast<float2*>(vo) = dvx;
      *reinterpret_cast<float2*>(vo + F) = dvy;
      *reinterpret_cast<float2*>(vo + 2 * F) = dvz;
    }
  }
}

// ============================================================================================
// backward: gather over the receiver's own row; edge A = i<-j together with its reverse B = j<-i.
// Outputs dphi[i], dv_in[i] (both skipped for the first layer) and the per-feature-half partial of
// dE/dx_i in gradp[m][h][i][3] (summed over halves in fixed order by grad_accum_kernel).
// ============================================================================================
template <bool FIRST> struct BwdOffsets {
  static constexpr int O_V = 3 * MSG_FC;                          // only when !FIRST
  static constexpr int O_DS = FIRST ? 3 * MSG_FC : 6 * MSG_FC;
  static constexpr int O_DV = O_DS + MSG_FC;
};

struct BwdOwn {   // receiver-side values of atom i for this lane's feature pair
  float2 pi0, pi1, pi2, gsi, gvix, gviy, gviz, vix, viy, viz;
};
struct BwdAcc {
  float2 dp0, dp1, dp2n, dvx, dvy, dvz, gnx, gny, gnz;
};

template <bool FIRST>
__device__ __forceinline__ void bwd_load_own(const float* __restrict__ si, BwdOwn& o) {
  using O = BwdOffsets<FIRST>;
  o.pi0 = ld2(si); o.pi1 = ld2(si + MSG_FC); o.pi2 = ld2(si + 2 * MSG_FC);
  o.gsi = ld2(si + O::O_DS);
  o.gvix = ld2(si + O::O_DV); o.gviy = ld2(si + O::O_DV + MSG_FC); o.gviz = ld2(si + O::O_DV + 2 * MSG_FC);
  o.vix = o.viy = o.viz = dup2(0.f);
  if (!FIRST) { o.vix = ld2(si + O::O_V); o.viy = ld2(si + O::O_V + MSG_FC); o.viz = ld2(si + O::O_V + 2 * MSG_FC); }
}

// the same values straight from global memory (own row outside the staged rows); pointers are the kernel's
// structure- and half-offset bases (phi: stride F3, ds: F, dv / v_in: 3F per atom)
template <bool FIRST>
__device__ __forceinline__ void bwd_load_own_global(const float* __restrict__ phi, const float* __restrict__ v_in,
                                                    const float* __restrict__ ds, const float* __restrict__ dv, int il, int lane,
                                                    BwdOwn& o) {
  const float* p = phi + (long long)il * F3 + 2 * lane;
  o.pi0 = ldg2(p); o.pi1 = ldg2(p + F); o.pi2 = ldg2(p + 2 * F);
  o.gsi = ldg2(ds + (long long)il * F + 2 * lane);
  const float* g = dv + (long long)il * 3 * F + 2 * lane;
  o.gvix = ldg2(g); o.gviy = ldg2(g + F); o.gviz = ldg2(g + 2 * F);
  o.vix = o.viy = o.viz = dup2(0.f);
  if (!FIRST) {
    const float* v = v_in + (long long)il * 3 * F + 2 * lane;
    o.vix = ldg2(v); o.viy = ldg2(v + F); o.viz = ldg2(v + 2 * F);
  }
}

// shared per-edge backward math given the filter rows w_k and their d-derivative q_k
template <bool FIRST>
__device__ __forceinline__ void bwd_edge(const float4 g, float inv_d, const float* __restrict__ sj, const BwdOwn& o, float2 w0,
                                         float2 w1, float2 w2, float2 q0, float2 q1, float2 q2, BwdAcc& a) {
  using O = BwdOffsets<FIRST>;
  const float2 pj1 = ld2(sj + MSG_FC), pj2 = ld2(sj + 2 * MSG_FC);
  const float2 gsj = ld2(sj + O::O_DS);
  const float2 gvjx = ld2(sj + O::O_DV), gvjy = ld2(sj + O::O_DV + MSG_FC), gvjz = ld2(sj + O::O_DV + 2 * MSG_FC);
  const float2 ux = dup2(g.x), uy = dup2(g.y), uz = dup2(g.z);
  // edge A (i receives from j): dxA1 = gsi, dxA2 = gvi.u, dxA0 = gvi.vj
  const float2 dxA2 = __ffma2_rn(o.gviz, uz, __ffma2_rn(o.gviy, uy, __fmul2_rn(o.gvix, ux)));
  // edge B (j receives from i, unit negated): dxB1 = gsj, dxB2 = -(gvj.u) =: -nB2, dxB0 = gvj.vi
  const float2 nB2 = __ffma2_rn(gvjz, uz, __ffma2_rn(gvjy, uy, __fmul2_rn(gvjx, ux)));
  // dd = sum_k (dwA_k + dwB_k) q_k
  const float2 t1 = __ffma2_rn(o.gsi, pj1, __fmul2_rn(gsj, o.pi1));              // dxA1*pj1 + dxB1*pi1
  const float2 t2 = __ffma2_rn(dxA2, pj2, neg2(__fmul2_rn(nB2, o.pi2)));         // dxA2*pj2 + dxB2*pi2
  float2 dd = __ffma2_rn(t2, q2, __fmul2_rn(t1, q1));
  if (!FIRST) {
    const float2 pj0 = ld2(sj);
    const float2 vjx = ld2(sj + O::O_V), vjy = ld2(sj + O::O_V + MSG_FC), vjz = ld2(sj + O::O_V + 2 * MSG_FC);
    const float2 dxA0 = __ffma2_rn(o.gviz, vjz, __ffma2_rn(o.gviy, vjy, __fmul2_rn(o.gvix, vjx)));
    const float2 dxB0 = __ffma2_rn(gvjz, o.viz, __ffma2_rn(gvjy, o.viy, __fmul2_rn(gvjx, o.vix)));
    const float2 t0 = __ffma2_rn(dxA0, pj0, __fmul2_rn(dxB0, o.pi0));
    dd = __ffma2_rn(t0, q0, dd);
    a.dp0 = __ffma2_rn(dxB0, w0, a.dp0);
    a.dp1 = __ffma2_rn(gsj, w1, a.dp1);
    a.dp2n = __ffma2_rn(nB2, w2, a.dp2n);
    const float2 tv = __fmul2_rn(o.pi0, w0);
    a.dvx = __ffma2_rn(tv, gvjx, a.dvx); a.dvy = __ffma2_rn(tv, gvjy, a.dvy); a.dvz = __ffma2_rn(tv, gvjz, a.dvz);
  }
  // unit-vector chain: delta = gvi*(pj2 w2) - gvj*(pi2 w2); project out the radial part
  const float2 ta = __fmul2_rn(pj2, w2), tbn = neg2(__fmul2_rn(o.pi2, w2));
  const float2 ex = __ffma2_rn(o.gvix, ta, __fmul2_rn(gvjx, tbn));
  const float2 ey = __ffma2_rn(o.gviy, ta, __fmul2_rn(gvjy, tbn));
  const float2 ez = __ffma2_rn(o.gviz, ta, __fmul2_rn(gvjz, tbn));
  const float2 proj = __ffma2_rn(ez, uz, __ffma2_rn(ey, uy, __fmul2_rn(ex, ux)));
  const float2 invd = dup2(inv_d);
  const float2 c = __ffma2_rn(neg2(proj), invd, dd);   // dd - proj/d  (multiplies u)
  a.gnx = __ffma2_rn(c, ux, __ffma2_rn(ex, invd, a.gnx));
  a.gny = __ffma2_rn(c, uy, __ffma2_rn(ey, invd, a.gny));
  a.gnz = __ffma2_rn(c, uz, __ffma2_rn(ez, invd, a.gnz));
}

template <bool FIRST>
__device__ __forceinline__ void bwd_store(const BwdOwn& o, const BwdAcc& a, int il, int i, int f0, int lane, int m, int h,
                                          int n_atoms, float* __restrict__ dphi, float* __restrict__ dv_in,
                                          float* __restrict__ gradp, int accum) {
  if (!FIRST) {
    float* dpo = dphi + (long long)il * F3 + f0;
    float* dvo = dv_in + (long long)il * 3 * F + f0;
    if (accum & 1) {
      red_add2(dpo, a.dp0);
      red_add2(dpo + F, a.dp1);
      red_add2(dpo + 2 * F, neg2(a.dp2n));
      red_add2(dvo, a.dvx);
      red_add2(dvo + F, a.dvy);
      red_add2(dvo + 2 * F, a.dvz);
    } else {
      *reinterpret_cast<float2*>(dpo) = a.dp0;
      *reinterpret_cast<float2*>(dpo + F) = a.dp1;
      *reinterpret_cast<float2*>(dpo + 2 * F) = neg2(a.dp2n);
      *reinterpret_cast<float2*>(dvo) = __fadd2_rn(o.gvix, a.dvx);
      *reinterpret_cast<float2*>(dvo + F) = __fadd2_rn(o.gviy, a.dvy);
      *reinterpret_cast<float2*>(dvo + 2 * F) = __fadd2_rn(o.gviz, a.dvz);
    }
  }
  const float gx = warp_sum(a.gnx.x + a.gnx.y), gy = warp_sum(a.gny.x + a.gny.y), gz = warp_sum(a.gnz.x + a.gnz.y);
  if (lane == 0) {
    float* gp = gradp + (((long long)m * 2 + h) * n_atoms + i) * 3;
    if (accum & 2) { atomicAdd(gp, -gx); atomicAdd(gp + 1, -gy); atomicAdd(gp + 2, -gz); }
    else { gp[0] = -gx; gp[1] = -gy; gp[2] = -gz; }
  }
}

template <bool FIRST>
__device__ __forceinline__ void bwd_stage(float* smem, const float* phi, const float* v_in, const float* ds, const float* dv,
                                          int n, int tid, int nthreads) {
  using O = BwdOffsets<FIRST>;
  constexpr int PER = MsgBwdLayout<FIRST>::PER;
  stage_rows(smem, PER, 0, phi, F3, F, 3, n, tid, nthreads);
  if (!FIRST) stage_rows(smem, PER, O::O_V, v_in, 3 * F, F, 3, n, tid, nthreads);
  stage_rows(smem, PER, O::O_DS, ds, F, F, 1, n, tid, nthreads);
  stage_rows(smem, PER, O::O_DV, dv, 3 * F, F, 3, n, tid, nthreads);
}

// memo pass (always writes its outputs; every atom of the structure is visited)
template <bool FIRST>
__global__ void __launch_bounds__(MEMO_THREADS_BWD, 1) message_bwd_memo(
    int layer, int n_atoms, const int32_t* __restrict__ atom_ptr, const int32_t* __restrict__ rowptr,
    const int32_t* __restrict__ order, const int32_t* __restrict__ nmemo, const float* __restrict__ mrec, FilterCacheView fc,
    const float* __restrict__ phi, const float* __restrict__ v_in, const float* __restrict__ ds,
    const float* __restrict__ dv, float* __restrict__ dphi, float* __restrict__ dv_in, float* __restrict__ gradp) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int row_ctr;
  if (threadIdx.x == 0) row_ctr = 0;
  constexpr int PER = MsgBwdLayout<FIRST>::PER;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x, h = blockIdx.y, m = blockIdx.z;
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
  const int ns = min(n, fc.n0);   // memoised edges join framework atoms: only their rows are ever gathered
  bwd_stage<FIRST>(smem, phi, v_in, ds, dv, ns, tid, MEMO_THREADS_BWD);
  float* ring = smem + (size_t)ns * PER + warp * (MEMO6_STAGES * MEMO6_STAGE_FLOATS);
  const int f0 = h * MSG_FC + 2 * lane;
  const long long ml = (long long)(m * NCONV + layer) * fc.nslots_cap * F3 + h * MSG_FC;
  const float* __restrict__ wbase = fc.wc + ml;
  const float* __restrict__ qbase = fc.qc + ml;
  stage_wait();
  __syncthreads();
  for (int t = next_row(&row_ctr, lane); t < n; t = next_row(&row_ctr, lane)) {
    const int il = __ldg(order + a0 + t);
    const int i = a0 + il;
    const float4* mr = reinterpret_cast<const float4*>(mrec + (long long)__ldg(rowptr + i) * MREC);
    const int ne = __ldg(nmemo + i);
    BwdOwn o;
    if (il < ns) bwd_load_own<FIRST>(smem + il * PER + 2 * lane, o);
    else bwd_load_own_global<FIRST>(phi, v_in, ds, dv, il, lane, o);
    BwdAcc a;
    a.dp0 = a.dp1 = a.dp2n = a.dvx = a.dvy = a.dvz = a.gnx = a.gny = a.gnz = dup2(0.f);
    memo_walk_ring6<!FIRST>(mr, ne, lane, wbase, qbase, ring, [&](const float4 g, int j, float inv_d, const float* r) {
      const float2 z2 = dup2(0.f);
      bwd_edge<FIRST>(g, inv_d, smem + (j - a0) * PER + 2 * lane, o, FIRST ? z2 : ld2(r), ld2(r + MSG_FC), ld2(r + 2 * MSG_FC),
                      FIRST ? z2 : ld2(r + 3 * MSG_FC), ld2(r + 4 * MSG_FC), ld2(r + 5 * MSG_FC), a);
    });
    bwd_store<FIRST>(o, a, il, i, f0, lane, m, h, n_atoms, dphi, dv_in, gradp, 0);
  }
}

// memo pass when the caller does not want gradients on frozen atoms (VSSR_FC_CONSTRAINED_GRAD): a memoised
// edge joins two frozen atoms, so its dE/dx terms (the q rows, the unit-vector chain, all of edge A) are
// never used; what remains is the state back-propagation through edge B,
//   dphi_i = (gv_j.v_i) w0 , gs_j w1 , -(gv_j.u) w2        dv_in_i = gv_i + (phi0_i w0) gv_j
// about a fifth of the arithmetic, 4 staged rows per atom instead of 10 and no q rows.  Layers > 0 only.
constexpr int MEMO_STATE_PER = 4 * MSG_FC;
__global__ void __launch_bounds__(MEMO_THREADS_FWD, 1) message_bwd_memo_state(
    int layer, int n_atoms, const int32_t* __restrict__ atom_ptr, const int32_t* __restrict__ rowptr,
    const int32_t* __restrict__ order, const int32_t* __restrict__ nmemo, const float* __restrict__ mrec, FilterCacheView fc,
    const float* __restrict__ phi, const float* __restrict__ v_in, const float* __restrict__ ds,
    const float* __restrict__ dv, float* __restrict__ dphi, float* __restrict__ dv_in,
    const int32_t* __restrict__ canonical, int n_struct, int group, int group_atoms) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int row_ctr;
  if (in_canonical_group(canonical, atom_ptr, blockIdx.x, n_struct, group, group_atoms)) return;   // done by message_bwd_memo_state_group
  if (threadIdx.x == 0) row_ctr = 0;
  constexpr int PER = MEMO_STATE_PER;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x, h = blockIdx.y, m = blockIdx.z;
  const int a0 = __ldg(atom_ptr + b), n = __ldg(atom_ptr + b + 1) - a0;
  const long long mA = (long long)m * n_atoms;
  const int f0 = h * MSG_FC + 2 * lane;
  phi += (mA + a0) * F3 + f0;
  v_in += (mA + a0) * 3 * F + f0;
  ds += (mA + a0) * F + h * MSG_FC;
  dv += (mA + a0) * 3 * F + h * MSG_FC;
  dphi += (mA + a0) * F3 + f0;
  dv_in += (mA + a0) * 3 * F + f0;
  const int ns = min(n, fc.n0);
  float* ring = smem + (size_t)ns * PER + warp * (MEMO_STAGES * MEMO_STAGE_FLOATS);
  stage_rows(smem, PER, 0, ds, F, F, 1, ns, tid, MEMO_THREADS_FWD);
  stage_rows(smem, PER, MSG_FC, dv, 3 * F, F, 3, ns, tid, MEMO_THREADS_FWD);
  const float* __restrict__ wbase = fc.wc + (long long)(m * NCONV + layer) * fc.nslots_cap * F3 + h * MSG_FC;
  stage_wait();
  __syncthreads();
  for (int t = next_row(&row_ctr, lane); t < n; t = next_row(&row_ctr, lane)) {
    const int il = __ldg(order + a0 + t);
    const int i = a0 + il;
    const float4* mr = reinterpret_cast<const float4*>(mrec + (long long)__ldg(rowptr + i) * MREC);
    const int ne = __ldg(nmemo + i);
    const float2 pi0 = ldg2(phi + (long long)il * F3);
    const float2 vix = ldg2(v_in + (long long)il * 3 * F), viy = ldg2(v_in + (long long)il * 3 * F + F),
                 viz = ldg2(v_in + (long long)il * 3 * F + 2 * F);
    float2 dp0 = dup2(0.f), dp1 = dup2(0.f), dp2n = dup2(0.f), dvx = dup2(0.f), dvy = dup2(0.f), dvz = dup2(0.f);
    memo_walk_ring<true>(mr, ne, lane, wbase, ring, [&](const float4 g, int j, float, const float* r) {
      const float* sj = smem + (j - a0) * PER + 2 * lane;
      const float2 gsj = ld2(sj), gvjx = ld2(sj + MSG_FC), gvjy = ld2(sj + 2 * MSG_FC), gvjz = ld2(sj + 3 * MSG_FC);
      const float2 w0 = ld2(r), w1 = ld2(r + MSG_FC), w2 = ld2(r + 2 * MSG_FC);
      // same operation order as bwd_edge, so both gradient modes give identical state
      const float2 nB2 = __ffma2_rn(gvjz, dup2(g.z), __ffma2_rn(gvjy, dup2(g.y), __fmul2_rn(gvjx, dup2(g.x))));
      const float2 dxB0 = __ffma2_rn(gvjz, viz, __ffma2_rn(gvjy, viy, __fmul2_rn(gvjx, vix)));
      dp0 = __ffma2_rn(dxB0, w0, dp0);
      dp1 = __ffma2_rn(gsj, w1, dp1);
      dp2n = __ffma2_rn(nB2, w2, dp2n);
      const float2 tv = __fmul2_rn(pi0, w0);
      dvx = __ffma2_rn(tv, gvjx, dvx); dvy = __ffma2_rn(tv, gvjy, dvy); dvz = __ffma2_rn(tv, gvjz, dvz);
    });
    float2 ox, oy, oz;      // own dv row
    if (il < ns) {
      const float* si = smem + il * PER + 2 * lane;
      ox = ld2(si + MSG_FC); oy = ld2(si + 2 * MSG_FC); oz = ld2(si + 3 * MSG_FC);
    } else {
      const float* dg = dv + (long long)il * 3 * F + 2 * lane;
      ox = ldg2(dg); oy = ldg2(dg + F); oz = ldg2(dg + 2 * F);
    }
    float* dpo = dphi + (long long)il * F3;
    float* dvo = dv_in + (long long)il * 3 * F;
    *reinterpret_cast<float2*>(dpo) = dp0;
    *reinterpret_cast<float2*>(dpo + F) = dp1;
    *reinterpret_cast<float2*>(dpo + 2 * F) = neg2(dp2n);
    *reinterpret_cast<float2*>(dvo) = __fadd2_rn(ox, dvx);
    *reinterpret_cast<float2*>(dvo + F) = __fadd2_rn(oy, dvy);
    *reinterpret_cast<float2*>(dvo + 2 * F) = __fadd2_rn(oz, dvz);
  }
}

// direct pass
template <bool FIRST>
__global__ void __launch_bounds__(MSG_THREADS, 1) message_bwd_v2(
    const float* __restrict__ weights, int layer, int n_atoms, const int32_t* __restrict__ atom_ptr, int n_chunks,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ order, const int32_t* __restrict__ nvalid,
    const float* __restrict__ erec,
    const float* __restrict__ phi, const float* __restrict__ v_in, const float* __restrict__ ds,
    const float* __restrict__ dv, float* __restrict__ dphi, float* __restrict__ dv_in, float* __restrict__ gradp,
    int accum, const uint8_t* __restrict__ frozen, int n0, int cap_atoms, int win) {
  extern __shared__ __align__(16) float smem_all[];
  __shared__ int row_ctr;
  if (threadIdx.x == 0) row_ctr = 0;
  constexpr int PER = MsgBwdLayout<FIRST>::PER;
  constexpr int N16 = (REC_DRE + 40) / 4;                  // 23 x 16 B: whole record
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* ring = smem_all + warp * MSG_STAGES * REC;
  float* smem = smem_all + MSG_WARPS * MSG_STAGES * REC;
  // grid = (halves x models, structures x chunks): the CTAs that stream the SAME edge records are neighbours in launch
  // order and run together, so the records are read from DRAM once and served to the other five from L2
  const int b = blockIdx.y / n_chunks, ch = blockIdx.y % n_chunks;
  const int h = blockIdx.x % (F / MSG_FC), m = blockIdx.x / (F / MSG_FC);
  const int a0 = __ldg(atom_ptr + b), n = __ldg(atom_ptr + b + 1) - a0;
  const long long mA = (long long)m * n_atoms;
  const SenderWindow wn = sender_window(n, cap_atoms, win);
  if (!wn.live) return;
  const bool whole = wn.lo == 0 && wn.hi == n;
  if (win > 0) accum = 3;                        // earlier windows started both the state outputs and gradp
  phi += (mA + a0) * F3 + h * MSG_FC;
  ds += (mA + a0) * F + h * MSG_FC;
  dv += (mA + a0) * 3 * F + h * MSG_FC;
  if (!FIRST) {
    v_in += (mA + a0) * 3 * F + h * MSG_FC;
    dphi += (mA + a0) * F3;
    dv_in += (mA + a0) * 3 * F;
  }
  bwd_stage<FIRST>(smem, phi + (long long)wn.lo * F3, FIRST ? v_in : v_in + (long long)wn.lo * 3 * F, ds + (long long)wn.lo * F,
                   dv + (long long)wn.lo * 3 * F, wn.hi - wn.lo, tid, MSG_THREADS);
  const int sbase = a0 + wn.lo;
  __shared__ RowTable rt;
  row_table_fill(rt, order, rowptr, nvalid, a0, n, tid, MSG_THREADS);

  const int f0 = h * MSG_FC + 2 * lane;
  const float* __restrict__ wl = weights + (long long)m * W_STRIDE + W_LAYER0 + (long long)layer * L_SIZE;
  float2 wd0[NRBF], wd1[NRBF], wd2[NRBF];
#pragma unroll
  for (int q = 0; q < NRBF; ++q) {
    wd0[q] = ld2(wl + L_WDT + q * F3 + f0);
    wd1[q] = ld2(wl + L_WDT + q * F3 + F + f0);
    wd2[q] = ld2(wl + L_WDT + q * F3 + 2 * F + f0);
  }
  const float2 bd0 = ld2(wl + L_BD + f0), bd1 = ld2(wl + L_BD + F + f0), bd2 = ld2(wl + L_BD + 2 * F + f0);
  if (!whole) {
    __syncthreads();                  // row table complete
    row_table_slice(rt, erec, a0, n, ch, n_chunks, wn.lo, wn.hi, warp, lane, MSG_THREADS / 32);
  }
  stage_wait();
  __syncthreads();

  for (int t = ch + n_chunks * next_row(&row_ctr, lane); t < n; t = ch + n_chunks * next_row(&row_ctr, lane)) {
    int il, e0, ne;
    row_table_get(rt, order, rowptr, nvalid, a0, t, il, e0, ne);
    const int i = a0 + il;
    const float* rec0 = erec + (long long)e0 * REC;
    if (!whole && t >= MSG_MAXROWS) {
      int e_lo, e_hi;
      window_slice(rec0, ne, a0, wn.lo, wn.hi, lane, e_lo, e_hi);
      rec0 += (long long)e_lo * REC;
      ne = e_hi - e_lo;
    }
    const bool own_staged = il >= wn.lo && il < wn.hi;
    if ((accum & 1) && ne == 0) {   // state already written by the memo pass / earlier windows; gradp may still need its zero
      if (!(accum & 2) && lane == 0) {
        float* gp = gradp + (((long long)m * 2 + h) * n_atoms + i) * 3;
        gp[0] = 0.f; gp[1] = 0.f; gp[2] = 0.f;
      }
      continue;
    }
    if (frozen && il < n0 && frozen[il]) {
      // constrained mode, frozen receiver: dE/dx_i is not wanted, so only the state terms of edge B remain
      // (filter w without its derivative: 60 + 13 FFMA2 instead of 180); nothing at all at the first layer
      if (lane == 0) {
        float* gp = gradp + (((long long)m * 2 + h) * n_atoms + i) * 3;
        gp[0] = 0.f; gp[1] = 0.f; gp[2] = 0.f;
      }
      if (FIRST) continue;
      using O = BwdOffsets<FIRST>;
      constexpr int N16L = (REC_RE + 44) / 4;   // geometry + rbf rows only
      BwdOwn ow;
      if (own_staged) bwd_load_own<FIRST>(smem + (il - wn.lo) * PER + 2 * lane, ow);
      else bwd_load_own_global<FIRST>(phi, v_in, ds, dv, il, lane, ow);
      const float2 pi0 = ow.pi0, vix = ow.vix, viy = ow.viy, viz = ow.viz;
      float2 dp0 = dup2(0.f), dp1 = dup2(0.f), dp2n = dup2(0.f), dvx = dup2(0.f), dvy = dup2(0.f), dvz = dup2(0.f);
      __syncwarp();
#pragma unroll
      for (int s = 0; s < MSG_STAGES - 1; ++s) prefetch_record(ring + s * REC, rec0 + (long long)s * REC, lane, N16L, s < ne);
      for (int e = 0; e < ne; ++e) {
        wait_record();
        const int nx = e + MSG_STAGES - 1;
        prefetch_record(ring + (nx % MSG_STAGES) * REC, rec0 + (long long)nx * REC, lane, N16L, nx < ne);
        const float* rec = ring + (e % MSG_STAGES) * REC;
        const float4 g = *reinterpret_cast<const float4*>(rec);
        const float4* r4 = reinterpret_cast<const float4*>(rec + REC_RE);
        const float4 ev = r4[10];
        const float2 env2 = make_float2(ev.x, ev.y);
        float2 w0 = __fmul2_rn(bd0, env2), w1 = __fmul2_rn(bd1, env2), w2 = __fmul2_rn(bd2, env2);
#pragma unroll
        for (int q = 0; q < NRBF / 2; ++q) {
          const float4 t4 = r4[q];
          const float2 ra = make_float2(t4.x, t4.y), rb = make_float2(t4.z, t4.w);
          w0 = __ffma2_rn(wd0[2 * q], ra, w0); w1 = __ffma2_rn(wd1[2 * q], ra, w1); w2 = __ffma2_rn(wd2[2 * q], ra, w2);
          w0 = __ffma2_rn(wd0[2 * q + 1], rb, w0); w1 = __ffma2_rn(wd1[2 * q + 1], rb, w1);
          w2 = __ffma2_rn(wd2[2 * q + 1], rb, w2);
        }
        const float* sj = smem + (__float_as_int(rec[REC_EJ]) - sbase) * PER + 2 * lane;
        const float2 gsj = ld2(sj + O::O_DS);
        const float2 gvjx = ld2(sj + O::O_DV), gvjy = ld2(sj + O::O_DV + MSG_FC), gvjz = ld2(sj + O::O_DV + 2 * MSG_FC);
        // same operation order as bwd_edge
        const float2 nB2 = __ffma2_rn(gvjz, dup2(g.z), __ffma2_rn(gvjy, dup2(g.y), __fmul2_rn(gvjx, dup2(g.x))));
        const float2 dxB0 = __ffma2_rn(gvjz, viz, __ffma2_rn(gvjy, viy, __fmul2_rn(gvjx, vix)));
        dp0 = __ffma2_rn(dxB0, w0, dp0);
        dp1 = __ffma2_rn(gsj, w1, dp1);
        dp2n = __ffma2_rn(nB2, w2, dp2n);
        const float2 tv = __fmul2_rn(pi0, w0);
        dvx = __ffma2_rn(tv, gvjx, dvx); dvy = __ffma2_rn(tv, gvjy, dvy); dvz = __ffma2_rn(tv, gvjz, dvz);
      }
      float* dpo = dphi + (long long)il * F3 + f0;
      float* dvo = dv_in + (long long)il * 3 * F + f0;
      if (accum & 1) {
        red_add2(dpo, dp0);
        red_add2(dpo + F, dp1);
        red_add2(dpo + 2 * F, neg2(dp2n));
        red_add2(dvo, dvx);
        red_add2(dvo + F, dvy);
        red_add2(dvo + 2 * F, dvz);
      } else {
        *reinterpret_cast<float2*>(dpo) = dp0;
        *reinterpret_cast<float2*>(dpo + F) = dp1;
        *reinterpret_cast<float2*>(dpo + 2 * F) = neg2(dp2n);
        *reinterpret_cast<float2*>(dvo) = __fadd2_rn(ow.gvix, dvx);
        *reinterpret_cast<float2*>(dvo + F) = __fadd2_rn(ow.gviy, dvy);
        *reinterpret_cast<float2*>(dvo + 2 * F) = __fadd2_rn(ow.gviz, dvz);
      }
      continue;
    }
    BwdOwn o;
    if (own_staged) bwd_load_own<FIRST>(smem + (il - wn.lo) * PER + 2 * lane, o);
    else bwd_load_own_global<FIRST>(phi, v_in, ds, dv, il, lane, o);
    BwdAcc a;
    a.dp0 = a.dp1 = a.dp2n = a.dvx = a.dvy = a.dvz = a.gnx = a.gny = a.gnz = dup2(0.f);
    __syncwarp();
#pragma unroll
    for (int s = 0; s < MSG_STAGES - 1; ++s) prefetch_record(ring + s * REC, rec0 + (long long)s * REC, lane, N16, s < ne);
    for (int e = 0; e < ne; ++e) {
      wait_record();
      const int nx = e + MSG_STAGES - 1;
      prefetch_record(ring + (nx % MSG_STAGES) * REC, rec0 + (long long)nx * REC, lane, N16, nx < ne);
      const float* rec = ring + (e % MSG_STAGES) * REC;
      const float4 g = *reinterpret_cast<const float4*>(rec);
      const float4* r4 = reinterpret_cast<const float4*>(rec + REC_RE);
      const float4* d4 = reinterpret_cast<const float4*>(rec + REC_DRE);
      const float4 ev = r4[10];
      const float2 env2 = make_float2(ev.x, ev.y), denv2 = make_float2(ev.z, ev.w);
      float2 w0 = __fmul2_rn(bd0, env2), w1 = __fmul2_rn(bd1, env2), w2 = __fmul2_rn(bd2, env2);
      float2 q0 = __fmul2_rn(bd0, denv2), q1 = __fmul2_rn(bd1, denv2), q2 = __fmul2_rn(bd2, denv2);
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
      bwd_edge<FIRST>(g, rec[7], smem + (__float_as_int(rec[REC_EJ]) - sbase) * PER + 2 * lane, o, w0, w1, w2, q0, q1, q2, a);
    }
    bwd_store<FIRST>(o, a, il, i, f0, lane, m, h, n_atoms, dphi, dv_in, gradp, accum);
  }
}

// ============================================================================================
// Pair kernels: two CANONICAL structures per CTA (row_order_kernel), walking the framework's own memoised
// lists.  Each filter row is loaded once (LDG, straight to registers) and used for both structures, so
// the per-edge LSU work -- the bound of the single-structure memo kernels -- nearly halves, and no
// per-chain records are read at all.  Same per-edge arithmetic in the same order => same bits.
// ============================================================================================
template <bool ROW0, int U, typename Body>
__device__ __forceinline__ void canon_walk(const float4* __restrict__ mr, int ne, int lane,
                                           const float* __restrict__ wlane, Body body) {
  for (int base = 0; base < ne; base += 32) {
    const int cnt = min(32, ne - base);
    float4 gl = make_float4(0.f, 0.f, 0.f, 1.f), jl = make_float4(0.f, 0.f, 1.f, 0.f);
    if (lane < cnt) { gl = __ldg(mr + 2 * (base + lane)); jl = __ldg(mr + 2 * (base + lane) + 1); }
#pragma unroll U
    for (int e = 0; e < cnt; ++e) {
      const int slot = __shfl_sync(0xffffffffu, __float_as_int(jl.y), e);
      const float* wr = wlane + (long long)slot * F3;
      const float2 w1 = ldg2(wr + F), w2 = ldg2(wr + 2 * F), w0 = ROW0 ? ldg2(wr) : dup2(0.f);
      const float4 g = make_float4(__shfl_sync(0xffffffffu, gl.x, e), __shfl_sync(0xffffffffu, gl.y, e),
                                   __shfl_sync(0xffffffffu, gl.z, e), __shfl_sync(0xffffffffu, gl.w, e));
      const int j = __shfl_sync(0xffffffffu, __float_as_int(jl.x), e);
      body(g, j, w0, w1, w2);
    }
  }
}

template <bool FIRST, int G, int T>
__global__ void __launch_bounds__(T, 1) message_fwd_memo_group(
    int layer, int n_atoms, const int32_t* __restrict__ atom_ptr, const int32_t* __restrict__ canonical,
    FilterCacheView fc, const float* __restrict__ phi, const float* __restrict__ s_in,
    const float* __restrict__ v_in, float* __restrict__ cat, float* __restrict__ v_mid, int n_struct, int group_atoms) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int row_ctr;
  constexpr int PER = MsgFwdLayout<FIRST>::PER;
  const int tid = threadIdx.x, lane = tid & 31;
  const int b0 = G * blockIdx.x, h = blockIdx.y, m = blockIdx.z;
  if (!in_canonical_group(canonical, atom_ptr, b0, n_struct, G, group_atoms)) return;   // left to the one-structure kernel
  if (tid == 0) row_ctr = 0;
  const long long mA = (long long)m * n_atoms;
  int a0[G], n[G];
  float* sm[G];
#pragma unroll
  for (int s = 0; s < G; ++s) { a0[s] = __ldg(atom_ptr + b0 + s); n[s] = __ldg(atom_ptr + b0 + s + 1) - a0[s]; }
  const int nst = fc.n0;      // staged rows per structure: the framework atoms (senders of every memoised edge)
#pragma unroll
  for (int s = 0; s < G; ++s) sm[s] = smem + (size_t)s * nst * PER;
#pragma unroll
  for (int s = 0; s < G; ++s) {
    stage_rows(sm[s], PER, 0, phi + (mA + a0[s]) * F3 + h * MSG_FC, F3, F, 3, nst, tid, T);
    if (!FIRST) stage_rows(sm[s], PER, 3 * MSG_FC, v_in + (mA + a0[s]) * 3 * F + h * MSG_FC, 3 * F, F, 3, nst, tid, T);
  }
  const int f0 = h * MSG_FC + 2 * lane;
  const float* __restrict__ wlane = fc.wc + (long long)(m * NCONV + layer) * fc.nslots_cap * F3 + f0;
  int nrows = n[0];
#pragma unroll
  for (int s = 1; s < G; ++s) nrows = max(nrows, n[s]);
  stage_wait();
  __syncthreads();
  for (int t = next_row(&row_ctr, lane); t < nrows; t = next_row(&row_ctr, lane)) {
    const int il = t < fc.n0 ? __ldg(fc.order0 + t) : t;   // rows past the framework (adsorbates) have no memoised edge
    float2 ds[G], dvx[G], dvy[G], dvz[G];
#pragma unroll
    for (int s = 0; s < G; ++s) ds[s] = dvx[s] = dvy[s] = dvz[s] = dup2(0.f);
    if (il < fc.n0) {
      const float4* mr = reinterpret_cast<const float4*>(fc.mrec0 + (long long)__ldg(fc.rowptr + il) * MREC);
      canon_walk<!FIRST, (T > 640 ? 2 : 4)>(mr, __ldg(fc.nmemo0 + il), lane, wlane, [&](const float4 g, int j, float2 w0, float2 w1, float2 w2) {
#pragma unroll
        for (int s = 0; s < G; ++s) fwd_edge<FIRST>(g, sm[s] + j * PER + 2 * lane, w0, w1, w2, ds[s], dvx[s], dvy[s], dvz[s]);
      });
    }
#pragma unroll
    for (int s = 0; s < G; ++s) {
      if (il >= n[s]) continue;
      const long long i = mA + a0[s] + il;
      const float2 s0 = ld2(s_in + i * F + f0);
      *reinterpret_cast<float2*>(cat + i * 2 * F + f0) = __fadd2_rn(s0, ds[s]);
      float2 ox = dvx[s], oy = dvy[s], oz = dvz[s];
      if (!FIRST) {
        if (il < nst) {
          const float* si = sm[s] + il * PER + 2 * lane;
          ox = __fadd2_rn(ox, ld2(si + 3 * MSG_FC));
          oy = __fadd2_rn(oy, ld2(si + 4 * MSG_FC));
          oz = __fadd2_rn(oz, ld2(si + 5 * MSG_FC));
        } else {      // adsorbate row: not staged, nothing memoised -> v_mid = v_in + 0 (same bits as the staged path)
          const float* vg = v_in + i * 3 * F + f0;
          ox = __fadd2_rn(ox, ldg2(vg)); oy = __fadd2_rn(oy, ldg2(vg + F)); oz = __fadd2_rn(oz, ldg2(vg + 2 * F));
        }
      }
      float* vo = v_mid + i * 3 * F + f0;
      *reinterpret_cast<float2*>(vo) = ox;
      *reinterpret_cast<float2*>(vo + F) = oy;
      *reinterpret_cast<float2*>(vo + 2 * F) = oz;
    }
  }
}

template <int G, int T>
__global__ void __launch_bounds__(T, 1) message_bwd_memo_state_group(
    int layer, int n_atoms, const int32_t* __restrict__ atom_ptr, const int32_t* __restrict__ canonical,
    FilterCacheView fc, const float* __restrict__ phi, const float* __restrict__ v_in,
    const float* __restrict__ ds, const float* __restrict__ dv, float* __restrict__ dphi, float* __restrict__ dv_in,
    int n_struct, int group_atoms) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int row_ctr;
  constexpr int PER = MEMO_STATE_PER;
  const int tid = threadIdx.x, lane = tid & 31;
  const int b0 = G * blockIdx.x, h = blockIdx.y, m = blockIdx.z;
  if (!in_canonical_group(canonical, atom_ptr, b0, n_struct, G, group_atoms)) return;   // left to the one-structure kernel
  if (tid == 0) row_ctr = 0;
  const long long mA = (long long)m * n_atoms;
  int a0[G], n[G];
  float* sm[G];
#pragma unroll
  for (int s = 0; s < G; ++s) { a0[s] = __ldg(atom_ptr + b0 + s); n[s] = __ldg(atom_ptr + b0 + s + 1) - a0[s]; }
  const int nst = fc.n0;
#pragma unroll
  for (int s = 0; s < G; ++s) sm[s] = smem + (size_t)s * nst * PER;
#pragma unroll
  for (int s = 0; s < G; ++s) {
    stage_rows(sm[s], PER, 0, ds + (mA + a0[s]) * F + h * MSG_FC, F, F, 1, nst, tid, T);
    stage_rows(sm[s], PER, MSG_FC, dv + (mA + a0[s]) * 3 * F + h * MSG_FC, 3 * F, F, 3, nst, tid, T);
  }
  const int f0 = h * MSG_FC + 2 * lane;
  const float* __restrict__ wlane = fc.wc + (long long)(m * NCONV + layer) * fc.nslots_cap * F3 + f0;
  int nrows = n[0];
#pragma unroll
  for (int s = 1; s < G; ++s) nrows = max(nrows, n[s]);
  stage_wait();
  __syncthreads();
  for (int t = next_row(&row_ctr, lane); t < nrows; t = next_row(&row_ctr, lane)) {
    const int il = t < fc.n0 ? __ldg(fc.order0 + t) : t;
    float2 pi0[G], vix[G], viy[G], viz[G], dp0[G], dp1[G], dp2n[G], dvx[G], dvy[G], dvz[G];
#pragma unroll
    for (int s = 0; s < G; ++s) {
      dp0[s] = dp1[s] = dp2n[s] = dvx[s] = dvy[s] = dvz[s] = dup2(0.f);
      pi0[s] = vix[s] = viy[s] = viz[s] = dup2(0.f);
      if (il < n[s] && il < fc.n0) {
        const long long i = mA + a0[s] + il;
        pi0[s] = ldg2(phi + i * F3 + f0);
        vix[s] = ldg2(v_in + i * 3 * F + f0); viy[s] = ldg2(v_in + i * 3 * F + F + f0); viz[s] = ldg2(v_in + i * 3 * F + 2 * F + f0);
      }
    }
    if (il < fc.n0) {
      const float4* mr = reinterpret_cast<const float4*>(fc.mrec0 + (long long)__ldg(fc.rowptr + il) * MREC);
      canon_walk<true, (T > 640 ? 2 : 4)>(mr, __ldg(fc.nmemo0 + il), lane, wlane, [&](const float4 g, int j, float2 w0, float2 w1, float2 w2) {
#pragma unroll
        for (int s = 0; s < G; ++s) {
          const float* sj = sm[s] + j * PER + 2 * lane;
          const float2 gsj = ld2(sj), gvjx = ld2(sj + MSG_FC), gvjy = ld2(sj + 2 * MSG_FC), gvjz = ld2(sj + 3 * MSG_FC);
          // same operation order as message_bwd_memo_state / bwd_edge
          const float2 nB2 = __ffma2_rn(gvjz, dup2(g.z), __ffma2_rn(gvjy, dup2(g.y), __fmul2_rn(gvjx, dup2(g.x))));
          const float2 dxB0 = __ffma2_rn(gvjz, viz[s], __ffma2_rn(gvjy, viy[s], __fmul2_rn(gvjx, vix[s])));
          dp0[s] = __ffma2_rn(dxB0, w0, dp0[s]);
          dp1[s] = __ffma2_rn(gsj, w1, dp1[s]);
          dp2n[s] = __ffma2_rn(nB2, w2, dp2n[s]);
          const float2 tv = __fmul2_rn(pi0[s], w0);
          dvx[s] = __ffma2_rn(tv, gvjx, dvx[s]); dvy[s] = __ffma2_rn(tv, gvjy, dvy[s]); dvz[s] = __ffma2_rn(tv, gvjz, dvz[s]);
        }
      });
    }
#pragma unroll
    for (int s = 0; s < G; ++s) {
      if (il >= n[s]) continue;
      const long long i = mA + a0[s] + il;
      float* dpo = dphi + i * F3 + f0;
      float* dvo = dv_in + i * 3 * F + f0;
      *reinterpret_cast<float2*>(dpo) = dp0[s];
      *reinterpret_cast<float2*>(dpo + F) = dp1[s];
      *reinterpret_cast<float2*>(dpo + 2 * F) = neg2(dp2n[s]);
      float2 ox, oy, oz;      // own dv row: staged for framework atoms, from global memory for adsorbate rows
      if (il < nst) {
        const float* si = sm[s] + il * PER + 2 * lane;
        ox = ld2(si + MSG_FC); oy = ld2(si + 2 * MSG_FC); oz = ld2(si + 3 * MSG_FC);
      } else {
        const float* dg = dv + i * 3 * F + f0;
        ox = ldg2(dg); oy = ldg2(dg + F); oz = ldg2(dg + 2 * F);
      }
      *reinterpret_cast<float2*>(dvo) = __fadd2_rn(ox, dvx[s]);
      *reinterpret_cast<float2*>(dvo + F) = __fadd2_rn(oy, dvy[s]);
      *reinterpret_cast<float2*>(dvo + 2 * F) = __fadd2_rn(oz, dvz[s]);
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

// VSSR_FC_CONSTRAINED_GRAD: gradient rows of the framework's frozen atoms are returned as zero
__global__ void zero_frozen_grad_kernel(const int32_t* __restrict__ atom_ptr, int n_struct, int n_atoms, int n0,
                                        const uint8_t* __restrict__ frozen, float* __restrict__ grad) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_atoms) return;
  const int il = a - __ldg(atom_ptr + struct_of_atom(atom_ptr, n_struct, a));
  if (il < n0 && frozen[il]) {
    float* g = grad + ((long long)blockIdx.y * n_atoms + a) * 3;
    g[0] = 0.f; g[1] = 0.f; g[2] = 0.f;
  }
}
