// PaiNN message passing, forward, DIRECT edges with the radial filter on the tensor core (sm_100a).
// Included by painn.cu inside its anonymous namespace, after gemm_tc.cuh and painn_message.cuh.
//
// north_star subsystem 2: "the RBF/cosine-cutoff expansion, filter ... as tcgen05 tensor-core GEMMs".  The filter of
// an edge, w_k[f] = sum_n Wd[k,f,n] rbf_n(d) env(d) + bd[k,f] env(d), is a K = 21 contraction; message_fwd_v2 spends
// 60 of its 72 FFMA2 per edge and lane on it.  Here it is one tcgen05.mma per 32-edge tile:
//
//   D[(block, feature) row, edge] = A[row, 0..23] . B[edge, 0..23]      A = Wd rows (+ bd in column 20), B = rbf*env (+ env)
//
// The per-edge message math separates by filter block (x0 = phi0.w0 feeds dv through v_j, x1 = phi1.w1 feeds ds,
// x2 = phi2.w2 feeds dv through u), so a consumer THREAD owns one (block, feature) pair = one TMEM lane and reads its
// 32 filter values of a tile with one tcgen05.ld.32x32b.x32 -- straight into registers, no transposition.
//
// CTA = (structure, row chunk, 64-feature half, model), 256 threads:
//   warps 0-1 : block 0, features 0..63 of the half   (TMEM lanes   0..63  of accumulator D1)
//   warps 2-3 : block 1                                (TMEM lanes  64..127 of accumulator D1)
//   warps 4-5 : block 2                                (TMEM lanes   0..63  of accumulator D2)
//   warps 6-9 : producers: warp 6+b owns tile buffer b (4 buffers): gathers the tile's edge records (a tile = up to 32
//               consecutive direct edges of ONE receiver row, zero-padded; rows in index order; the records of the warp's
//               NEXT tile are already in flight while the current one is processed), splits rbf*env into TF32 hi + lo, writes the
//               K-major SWIZZLE_128B operand tile and the per-edge metadata (unit vector, sender, row), then one lane
//               issues the 18 tcgen05.mma (3xTF32: A_lo.B_hi + A_hi.B_lo first, then A_hi.B_hi; 3 k-steps; D1, D2)
//               and tcgen05.commit's to the tile's mbarrier.
// All consumer threads walk the same edge stream in the same fixed order (no atomics; results depend on the structure
// only).  The two dv partials of a receiver (blocks 0 and 2) meet in shared memory once per ROW.  Sender windows,
// accum semantics and outputs are those of message_fwd_v2.
#pragma once

namespace mtc {

using tc::fence_after;
using tc::fence_async_smem;
using tc::fence_before;
using tc::make_desc;
using tc::make_idesc;
using tc::mbar_arrive;
using tc::mbar_init;
using tc::mbar_wait;
using tc::mma_commit;
using tc::mma_tf32;
using tc::smem_u32;
using tc::swz;
using tc::tf32_rn;
using tc::tmem_ld32;

// TF32 part of x by integer rounding (nearest, ties away: what cvt.rna.tf32.f32 computes for finite x, without the
// NaN/Inf special-casing the compiler wraps around the cvt -- the operands here are finite by construction)
__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

constexpr int NBUF = 4;               // tile buffers = producer warps
constexpr int THREADS = (6 + NBUF) * 32;
constexpr int TILE = 32;              // edges per tile = MMA N
constexpr int MAXROWS = 128;          // rows of one (structure, chunk)
constexpr int MAXTILES = 512;         // tiles of one (structure, chunk, window)
constexpr int A_BYTES = 128 * 128;    // operand tile: 128 rows x 32 fp32 (K padded 24 -> 32), SWIZZLE_128B
constexpr int A2_BYTES = 64 * 128;    // block 2 has 64 rows; the MMA's rows 64..127 read whatever follows (lanes never read)
constexpr int B_BYTES = TILE * 128;
constexpr int META_FLOATS = 4;        // per edge: ux, uy, uz, sender (local to the window)
// A1hi A1lo A2hi A2lo | B[NBUF][hi,lo] | meta[NBUF]
constexpr int FIXED_BYTES = 2 * A_BYTES + 2 * A2_BYTES + NBUF * 2 * B_BYTES + NBUF * TILE * META_FLOATS * 4;
constexpr int TMEM_COLS = 256;        // NBUF tile buffers x (D1: 32 + D2: 32) columns

__host__ __device__ constexpr size_t fwd_smem_bytes(int rows, bool first) {
  return 1024 + (size_t)FIXED_BYTES + (size_t)rows * (first ? 3 : 6) * MSG_FC * 4;
}

template <bool FIRST>
__global__ void __launch_bounds__(THREADS, 1) message_fwd_tc(
    const float* __restrict__ weights, int layer, int n_atoms, const int32_t* __restrict__ atom_ptr, int n_chunks,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nvalid, const float* __restrict__ erec,
    const float* __restrict__ phi, const float* __restrict__ s_in, const float* __restrict__ v_in,
    float* __restrict__ cat, float* __restrict__ v_mid, int accum, int cap_atoms, int win) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[NBUF], empty_bar[NBUF];
  __shared__ uint8_t s_tile_row[MAXTILES], s_tile_off[MAXTILES];
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_cum[MAXROWS + 1], s_ril[MAXROWS], s_re0[MAXROWS], s_rne[MAXROWS];
  __shared__ float s_dv[2][3][MSG_FC];
  constexpr int PER = MsgFwdLayout<FIRST>::PER;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y / n_chunks, ch = blockIdx.y % n_chunks;
  const int h = blockIdx.x % (F / MSG_FC), m = blockIdx.x / (F / MSG_FC);
  const int a0 = __ldg(atom_ptr + b), n = __ldg(atom_ptr + b + 1) - a0;
  const SenderWindow wn = sender_window(n, cap_atoms, win);
  if (!wn.live) return;
  const bool whole = wn.lo == 0 && wn.hi == n;
  if (win > 0) accum = 1;
  const long long mA = (long long)m * n_atoms;
  phi += (mA + a0) * F3 + h * MSG_FC;
  s_in += (mA + a0) * F + h * MSG_FC;
  cat += (mA + a0) * 2 * F + h * MSG_FC;
  v_mid += (mA + a0) * 3 * F + h * MSG_FC;
  if (!FIRST) v_in += (mA + a0) * 3 * F + h * MSG_FC;

  // (pointer arithmetic on the shared array, not an integer round trip: the compiler must keep the shared address space)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = base;                                   // A1hi, A1lo, A2hi, A2lo
  uint8_t* sA2 = base + 2 * A_BYTES;
  uint8_t* sB = sA2 + 2 * A2_BYTES;                     // [buf][hi, lo]
  float* sMeta = reinterpret_cast<float*>(sB + NBUF * 2 * B_BYTES);   // [buf][TILE][META_FLOATS]
  float* rows = sMeta + NBUF * TILE * META_FLOATS;      // staged phi (, v) rows of the sender window

  // ---- row table of this chunk: rows il = ch, ch + n_chunks, ... ; their (window slice of the) direct records ----
  const int nr = n > ch ? (n - ch + n_chunks - 1) / n_chunks : 0;
  for (int k = warp; k < nr && k < MAXROWS; k += THREADS / 32) {
    const int il = ch + k * n_chunks;
    int e0 = __ldg(rowptr + a0 + il), ne = __ldg(nvalid + a0 + il);
    if (!whole) {
      int e_lo, e_hi;
      window_slice(erec + (long long)e0 * REC, ne, a0, wn.lo, wn.hi, lane, e_lo, e_hi);
      e0 += e_lo;
      ne = e_hi - e_lo;
    }
    if (lane == 0) { s_ril[k] = il; s_re0[k] = e0; s_rne[k] = ne; }
  }
  // ---- stage the window's sender rows ----
  stage_rows(rows, PER, 0, phi + (long long)wn.lo * F3, F3, F, 3, wn.hi - wn.lo, tid, THREADS);
  if (!FIRST) stage_rows(rows, PER, 3 * MSG_FC, v_in + (long long)wn.lo * 3 * F, 3 * F, F, 3, wn.hi - wn.lo, tid, THREADS);
  // ---- weight operand: A1 = [block 0 | block 1] rows, A2 = [block 2 | zero] rows; K = 20 rbf columns + bias column ----
  {
    const float* __restrict__ wl = weights + (long long)m * W_STRIDE + W_LAYER0 + (long long)layer * L_SIZE;
    for (int idx = tid; idx < 192 * 8; idx += THREADS) {
      const int rr = idx >> 3, c = idx & 7;
      const int blk = rr >> 6, mat = blk >> 1, r = mat ? (rr & 63) : rr;   // A1: rows 0..127 = blocks 0,1; A2: rows 0..63 = block 2
      const int f = h * MSG_FC + (rr & 63);
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kx = 4 * c + q;
        v[q] = 0.f;
        if (blk < 3) {
          if (kx < NRBF) v[q] = __ldg(wl + L_WDT + kx * F3 + blk * F + f);
          else if (kx == NRBF) v[q] = __ldg(wl + L_BD + blk * F + f);
        }
      }
      float4 hi, lo;
      hi.x = tf32_hi(v[0]); lo.x = tf32_hi(v[0] - hi.x);
      hi.y = tf32_hi(v[1]); lo.y = tf32_hi(v[1] - hi.y);
      hi.z = tf32_hi(v[2]); lo.z = tf32_hi(v[2] - hi.z);
      hi.w = tf32_hi(v[3]); lo.w = tf32_hi(v[3] - hi.w);
      const uint32_t off = swz(r, c);
      *reinterpret_cast<float4*>((mat ? sA2 : sA) + off) = hi;
      *reinterpret_cast<float4*>((mat ? sA2 + A2_BYTES : sA + A_BYTES) + off) = lo;
    }
  }
  if (tid == 0) {
    for (int q = 0; q < NBUF; ++q) {
      mbar_init(&full_bar[q], 2);      // tcgen05.commit + the producer's own arrive
      mbar_init(&empty_bar[q], 6);     // the six consumer warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  stage_wait();
  fence_async_smem();            // the weight tiles were written through the generic proxy; the tensor core reads them
  __syncthreads();
  const int nrows = nr < MAXROWS ? nr : MAXROWS;
  if (tid == 0) {             // tiles never span rows: row k owns tiles [cum[k], cum[k+1])
    int c = 0;
    s_cum[0] = 0;
    for (int k = 0; k < nrows; ++k) { c += (s_rne[k] + TILE - 1) / TILE; s_cum[k + 1] = c; }
  }
  __syncthreads();
  const int ntiles = s_cum[nrows] < MAXTILES ? s_cum[nrows] : MAXTILES;
  for (int k = tid; k < nrows; k += THREADS)
    for (int t = s_cum[k]; t < s_cum[k + 1] && t < MAXTILES; ++t) { s_tile_row[t] = (uint8_t)k; s_tile_off[t] = (uint8_t)(t - s_cum[k]); }
  // rows without a direct edge in this window: when nothing ran before us their outputs still have to be written
  if (!accum && tid < MSG_FC) {
    for (int k = 0; k < nrows; ++k) {
      if (s_rne[k] != 0) continue;
      const int il = s_ril[k];
      cat[(long long)il * 2 * F + tid] = s_in[(long long)il * F + tid];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        v_mid[(long long)il * 3 * F + c * F + tid] = FIRST ? 0.f : __ldg(v_in + (long long)il * 3 * F + c * F + tid);
    }
  }
  if (ntiles == 0) return;       // (uniform: no TMEM allocated yet)
  __syncthreads();               // tile table complete

  if (warp == 6) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem0 = tmem_base_s;

  if (warp >= 6) {
    // ============================== producers (+ MMA issue) ==============================
    const int bf = warp - 6;
    uint8_t* sBh = sB + (2 * bf) * B_BYTES;
    uint8_t* sBl = sBh + B_BYTES;
    float* meta = sMeta + bf * TILE * META_FLOATS;
    constexpr uint32_t idesc = make_idesc(128, TILE);
    const uint64_t dA1h = make_desc(smem_u32(sA)), dA1l = make_desc(smem_u32(sA + A_BYTES));
    const uint64_t dA2h = make_desc(smem_u32(sA2)), dA2l = make_desc(smem_u32(sA2 + A2_BYTES));
    const uint64_t dBh = make_desc(smem_u32(sBh)), dBl = make_desc(smem_u32(sBl));
    const uint32_t d1 = tmem0 + bf * 64, d2 = d1 + 32;
    // the record of slot (t, lane): issued one tile ahead so that its global-memory latency hides behind the
    // consumers' work on the buffer's previous tile
    float4 g, r4v[NRBF / 2 + 1];
    int jl;
    auto fetch = [&](int t) {
      g = make_float4(0.f, 0.f, 0.f, 1.f);
      jl = 0;
#pragma unroll
      for (int q = 0; q <= NRBF / 2; ++q) r4v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < ntiles) {
        const int k = s_tile_row[t], e = s_tile_off[t] * TILE + lane;
        if (e < s_rne[k]) {
          const float* rec = erec + (long long)(s_re0[k] + e) * REC;
          g = __ldg(reinterpret_cast<const float4*>(rec));
          jl = __float_as_int(__ldg(rec + REC_EJ)) - a0 - wn.lo;
          const float4* r4 = reinterpret_cast<const float4*>(rec + REC_RE);
#pragma unroll
          for (int q = 0; q <= NRBF / 2; ++q) r4v[q] = __ldg(r4 + q);
        }
      }
    };
    fetch(bf);
    uint32_t use = 0;
    for (int t = bf; t < ntiles; t += NBUF, ++use) {
      float vals[24];
#pragma unroll
      for (int q = 0; q < NRBF / 2; ++q) { vals[2 * q] = r4v[q].x; vals[2 * q + 1] = r4v[q].z; }
      vals[NRBF] = r4v[NRBF / 2].x;             // env: multiplies the bias column
      vals[21] = vals[22] = vals[23] = 0.f;
      const float4 gm = make_float4(g.x, g.y, g.z, __int_as_float(jl));
      // hi/lo split BEFORE waiting for the buffer: only the stores and the MMA issue sit on the hand-over path
      float4 hi4[6], lo4[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        hi4[c].x = tf32_hi(vals[4 * c]);     lo4[c].x = tf32_hi(vals[4 * c] - hi4[c].x);
        hi4[c].y = tf32_hi(vals[4 * c + 1]); lo4[c].y = tf32_hi(vals[4 * c + 1] - hi4[c].y);
        hi4[c].z = tf32_hi(vals[4 * c + 2]); lo4[c].z = tf32_hi(vals[4 * c + 2] - hi4[c].z);
        hi4[c].w = tf32_hi(vals[4 * c + 3]); lo4[c].w = tf32_hi(vals[4 * c + 3] - hi4[c].w);
      }
      fetch(t + NBUF);                            // next tile of this buffer: loads in flight from here on
      mbar_wait(&empty_bar[bf], (use & 1) ^ 1);   // consumers are done with this buffer's previous tile
      fence_after();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t off = swz(lane, c);
        *reinterpret_cast<float4*>(sBh + off) = c < 6 ? hi4[c] : z4;
        *reinterpret_cast<float4*>(sBl + off) = c < 6 ? lo4[c] : z4;
      }
      *reinterpret_cast<float4*>(meta + lane * META_FLOATS) = gm;     // padding slots: zero filter, sender 0
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {        // K = 24: three k-steps of 8; small terms first
          const uint64_t adv = (uint64_t)(kk * 32 >> 4);
          mma_tf32(d1, dA1l + adv, dBh + adv, idesc, kk ? 1u : 0u);
          mma_tf32(d1, dA1h + adv, dBl + adv, idesc, 1u);
          mma_tf32(d2, dA2l + adv, dBh + adv, idesc, kk ? 1u : 0u);
          mma_tf32(d2, dA2h + adv, dBl + adv, idesc, 1u);
        }
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const uint64_t adv = (uint64_t)(kk * 32 >> 4);
          mma_tf32(d1, dA1h + adv, dBh + adv, idesc, 1u);
          mma_tf32(d2, dA2h + adv, dBh + adv, idesc, 1u);
        }
        mma_commit(&full_bar[bf]);              // arrives when the MMAs above have completed
        mbar_arrive(&full_bar[bf]);             // releases this warp's metadata writes to the consumers
      }
      __syncwarp();
    }
  } else {
    // ============================== consumers ==============================
    const int blk = warp >> 1;
    const int fl = (warp & 1) * 32 + lane;      // feature within the half
    const uint32_t tbase = tmem0 + ((uint32_t)(32 * (warp & 3)) << 16) + (blk == 2 ? 32u : 0u);
    int cur = -1, nflush = 0;
    float ds = 0.f, dvx = 0.f, dvy = 0.f, dvz = 0.f;

    // Row outputs.  accum (a memo pass or an earlier sender window started the row): ONE fire-and-forget red.add per
    // address and launch -- a single commutative addition onto the stored value, so the result is the same bits as
    // load + add + store, without the load latency that every consumer of this CTA would wait for (all of them walk
    // the same row).  !accum: the base values (s_in, v_in) are requested when the row STARTS and consumed here.
    float base_s = 0.f, base_x = 0.f, base_y = 0.f, base_z = 0.f;
    auto begin_row = [&](int k) {
      if (accum) return;
      const int il = s_ril[k];
      if (blk == 1) base_s = __ldg(s_in + (long long)il * F + fl);
      else if (blk == 0 && !FIRST) {
        const float* vi = v_in + (long long)il * 3 * F + fl;
        base_x = __ldg(vi); base_y = __ldg(vi + F); base_z = __ldg(vi + 2 * F);
      }
    };
    auto flush = [&](int k) {
      const int il = s_ril[k];
      if (blk == 1) {
        float* so = cat + (long long)il * 2 * F + fl;
        if (accum) atomicAdd(so, ds); else *so = base_s + ds;
      } else if (FIRST) {
        if (blk == 2) {
          float* vo = v_mid + (long long)il * 3 * F + fl;
          if (accum) { atomicAdd(vo, dvx); atomicAdd(vo + F, dvy); atomicAdd(vo + 2 * F, dvz); }
          else { vo[0] = dvx; vo[F] = dvy; vo[2 * F] = dvz; }
        }
      } else {
        const int par = nflush & 1;
        if (blk == 2) { s_dv[par][0][fl] = dvx; s_dv[par][1][fl] = dvy; s_dv[par][2][fl] = dvz; }
        asm volatile("bar.sync 1, 128;" ::: "memory");      // warps 0,1 (block 0) and 4,5 (block 2): same stream, same rows
        if (blk == 0) {
          float* vo = v_mid + (long long)il * 3 * F + fl;
          const float tx = dvx + s_dv[par][0][fl], ty = dvy + s_dv[par][1][fl], tz = dvz + s_dv[par][2][fl];
          if (accum) { atomicAdd(vo, tx); atomicAdd(vo + F, ty); atomicAdd(vo + 2 * F, tz); }
          else { vo[0] = base_x + tx; vo[F] = base_y + ty; vo[2 * F] = base_z + tz; }
        }
      }
      ++nflush;
      ds = dvx = dvy = dvz = 0.f;
    };

    for (int t = 0; t < ntiles; ++t) {
      const int bf = t % NBUF;
      const int row = s_tile_row[t];
      if (row != cur) {                          // tiles never span rows: one check per tile, none per edge
        if (cur >= 0) flush(cur);
        cur = row;
        begin_row(row);
      }
      mbar_wait(&full_bar[bf], (t / NBUF) & 1);
      fence_after();
      float wv[TILE];
      tmem_ld32(tbase + bf * 64, wv);
      fence_before();
      const float4* meta = reinterpret_cast<const float4*>(sMeta + bf * TILE * META_FLOATS);
      // one straight-line, fully unrolled loop per block role (a branch on the role inside the loop body would fence
      // the scheduler in: with six consumer warps per SM the 32 independent gathers of a tile are the only latency cover)
      const float* rf = rows + fl;
      if (blk == 1) {
#pragma unroll
        for (int e = 0; e < TILE; ++e) {
          const int j = __float_as_int(meta[e].w);
          ds = fmaf(rf[j * PER + MSG_FC], wv[e], ds);
        }
      } else if (blk == 2) {
#pragma unroll
        for (int e = 0; e < TILE; ++e) {
          const float4 mu = meta[e];
          const float x2 = rf[__float_as_int(mu.w) * PER + 2 * MSG_FC] * wv[e];
          dvx = fmaf(x2, mu.x, dvx); dvy = fmaf(x2, mu.y, dvy); dvz = fmaf(x2, mu.z, dvz);
        }
      } else if (!FIRST) {
#pragma unroll
        for (int e = 0; e < TILE; ++e) {
          const float* sj = rf + __float_as_int(meta[e].w) * PER;
          const float x0 = sj[0] * wv[e];
          dvx = fmaf(x0, sj[3 * MSG_FC], dvx); dvy = fmaf(x0, sj[4 * MSG_FC], dvy); dvz = fmaf(x0, sj[5 * MSG_FC], dvz);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[bf]);
    }
    if (cur >= 0) flush(cur);
  }
  fence_before();
  __syncthreads();
  if (warp == 6) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "n"(TMEM_COLS));
  }
}

}  // namespace mtc
