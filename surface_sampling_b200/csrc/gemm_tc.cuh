// tcgen05 (5th-gen tensor core) GEMM with fp32-grade accuracy via the 3xTF32 split:
//     C[M,N] (op)= T(A)[M,K] . B[K,N]      A = A_hi + A_lo, B = B_hi + B_lo  (hi = TF32 part)
//     D = A_hi.B_hi + A_lo.B_hi + A_hi.B_lo          (fp32 accumulation in TMEM)
// Included by painn.cu inside its anonymous namespace; replaces gemm_kernel for the PaiNN node MLPs
// (same GemmArgs / AMODE / EPI contract), selected at run time (VSSR_GEMM=fma keeps the FFMA2 path).
//
// Blackwell specifics used here (sm_100a only):
//   * tcgen05.mma.cta_group::1.kind::tf32, M=128, N=BN, K=8 per instruction, issued by ONE thread,
//     both operands K-major in shared memory behind 64-bit UMMA descriptors (SWIZZLE_128B atoms:
//     8 rows x 128 B, 16-byte chunk index XOR row%8, SBO = 1024 B);
//   * accumulator tile 128 lanes x BN columns in TMEM (tcgen05.alloc / dealloc by warp 0),
//     read back with tcgen05.ld.32x32b.x32 (one row per thread) for the fused epilogue;
//   * completion tracking with tcgen05.commit -> mbarrier.
// The A operand is produced by the CTA itself (global fp32 -> optional swish / dswish transform ->
// hi/lo split -> swizzled st.shared), so no TMA tensor maps are needed for activations; the weight
// operand B arrives pre-split (hi/lo, K-major [N][K]) from the packed weight block.
// Tiles are walked persistently (tile = blockIdx.x + k*gridDim.x) by one warp-specialised CTA per SM.
#pragma once

namespace tc {

constexpr int BM = 128, BK = 32;     // BK fp32 = 128 bytes = one swizzle row
constexpr int THREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}

// K-major operand tile, SWIZZLE_128B, dense 8-row groups (SBO = 1024 B); LBO unused (one atom along K)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
  d |= (uint64_t)(1024u >> 4) << 32;          // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                     // layout type: SWIZZLE_128B
  return d;
}

// instruction descriptor: D=F32, A=B=TF32, both K-major, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// byte offset of 16-byte chunk c (0..7) of row r inside a [rows][32 fp32] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

struct TcArgs {
  GemmArgs g;            // A, lda, sA, bias, aux, avec, C, ... as for gemm_kernel (g.B/ldb unused)
  const float* Bhi;      // K-major [N][K] TF32 part of the weights
  const float* Blo;      // K-major [N][K] remainder
  long long sBw;         // per-model stride of Bhi/Blo
  int n_models;
  int mode;              // 0: 3xTF32 (default), 1: plain TF32 (accuracy studies)
  unsigned long long* dbg;  // optional [gridDim.x][8] globaltimer stamps (VSSR_TC_DEBUG=1)
};


// ------------------------------------------------------------------------------------------
// Warp-specialised, persistent version (default).  One CTA per SM, 11 warps:
//   warps 0-11 producers: stage s of the 3-deep operand ring is owned by warps 4s..4s+3 (32 rows of A
//              and BN/4 rows of B each); A goes global -> registers -> transform/split -> swizzled
//              st.shared, B (pre-split weights) goes global -> shared with cp.async (no registers)
//   warp  12   MMA issuer: one elected lane waits full[s], issues 12 tcgen05.mma per chunk, and
//              tcgen05.commit's to empty[s] (and to tmem_full[acc] after the last chunk of a tile)
//   warps 13-16 epilogue: TMEM lane quarter (warp&3) -> registers -> fused op -> global
// TMEM holds two accumulator sets (main + correction, 2 x 2BN = 512 columns) so the epilogue of
// tile t overlaps the MMAs of tile t+1; 3 x 64 KB ring stages keep ~3 chunks of loads in flight.
// ------------------------------------------------------------------------------------------
constexpr int WS_STAGES = 3;
constexpr int WS_PW = 4;                       // producer warps per ring stage
constexpr int WS_PRODUCER_WARPS = WS_PW * WS_STAGES;
constexpr int WS_MMA_WARP = WS_PRODUCER_WARPS;
constexpr int WS_EPI_WARP0 = WS_MMA_WARP + 1;
constexpr int WS_THREADS = (WS_EPI_WARP0 + 4) * 32;

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(k) do { if (p.dbg) p.dbg[blockIdx.x * 16 + (k)] = gtime(); } while (0)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BN, int AMODE, int EPI>
__global__ void __launch_bounds__(WS_THREADS, 1) gemm_tc_ws_kernel(TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the shared array (an integer round trip would lose the shared address space: generic LD/ST)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  __shared__ __align__(8) uint64_t full_bar[WS_STAGES], empty_bar[WS_STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const GemmArgs& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) TC_STAMP(0);
  if (warp == WS_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(4 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < WS_STAGES; ++s) { mbar_init(&full_bar[s], WS_PW); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem0 = tmem_base_s;
  if (tid == 0) TC_STAMP(1);

  const int m_tiles = (g.M + BM - 1) / BM, n_tiles = g.N / BN;
  const int total = m_tiles * n_tiles * p.n_models;
  const int nchunks = g.K / BK;
  constexpr uint32_t idesc = make_idesc(BM, BN);

  if (warp < WS_PRODUCER_WARPS) {
    // ================= producers =================
    const int s = warp / WS_PW, part = warp % WS_PW;   // this warp fills rows [part*BM/WS_PW, ...) of stage s
    uint8_t* st = base + s * STAGE_BYTES;
    uint8_t* sAhi = st; uint8_t* sAlo = st + A_BYTES; uint8_t* sBhi = st + 2 * A_BYTES; uint8_t* sBlo = sBhi + B_BYTES;
    int gchunk = 0;   // running chunk counter of this CTA
    uint32_t use = 0; // how many times this stage was filled
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int model = t / (m_tiles * n_tiles);
      const int rem = t % (m_tiles * n_tiles);
      const int m0 = (rem / n_tiles) * BM, n0 = (rem % n_tiles) * BN;
      const float* __restrict__ A = g.A + (long long)model * g.sA;
      const float* __restrict__ Bh = p.Bhi + (long long)model * p.sBw + (long long)n0 * g.K;
      const float* __restrict__ Bl = p.Blo + (long long)model * p.sBw + (long long)n0 * g.K;
      const float* __restrict__ avec = AMODE == 2 ? g.avec + (long long)model * g.sAvec : nullptr;
      for (int c = 0; c < nchunks; ++c, ++gchunk) {
        if (gchunk % WS_STAGES != s) continue;
        const int k0 = c * BK;
        // A (global -> registers) is requested BEFORE waiting for the stage to drain: the loads need no
        // shared memory, so their latency overlaps the MMAs still reading this stage
        constexpr int NQ = (BM / WS_PW) * 8 / 32;   // float4 per lane and chunk (8)
        float4 x[NQ];
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
          const int idx = lane + u * 32;
          const int r = part * (BM / WS_PW) + (idx >> 3), ch = idx & 7;
          x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m0 + r < g.M) x[u] = __ldg(reinterpret_cast<const float4*>(A + (long long)(m0 + r) * g.lda + k0) + ch);
        }
        float4 av[AMODE == 2 ? 1 : 1];
        if (AMODE == 2) av[0] = __ldg(reinterpret_cast<const float4*>(avec + k0) + (lane & 7));
        mbar_wait(&empty_bar[s], (use & 1) ^ 1);
        ++use;
        const bool stamp = (warp == 0 && lane == 0 && use <= 2);
        if (stamp) TC_STAMP(use == 1 ? 8 : 12);
        // B: rows [part*BN/4, +BN/4) of hi and lo, straight to swizzled smem with cp.async
#pragma unroll 4
        for (int q = 0; q < (BN / WS_PW) * 8 / 32; ++q) {
          const int idx = lane + q * 32;
          const int r = part * (BN / WS_PW) + (idx >> 3), ch = idx & 7;
          const uint32_t off = swz(r, ch);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sBhi + off)),
                       "l"(Bh + (long long)r * g.K + k0 + ch * 4));
          if (p.mode != 2)   // mode 2: timing experiment only (wrong results): B_lo not fetched
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sBlo + off)),
                         "l"(Bl + (long long)r * g.K + k0 + ch * 4));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (stamp) TC_STAMP(use == 1 ? 9 : 13);
        // A: registers -> transform -> TF32 split -> swizzled smem
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
          const int idx = lane + u * 32;
          const int r = part * (BM / WS_PW) + (idx >> 3), ch = idx & 7;   // ch == lane & 7
          float4 v = x[u];
          if (AMODE == 1) {
            v.x = swishf_(v.x); v.y = swishf_(v.y); v.z = swishf_(v.z); v.w = swishf_(v.w);
          } else if (AMODE == 2) {
            v.x = dswishf_(v.x) * av[0].x; v.y = dswishf_(v.y) * av[0].y; v.z = dswishf_(v.z) * av[0].z; v.w = dswishf_(v.w) * av[0].w;
          }
          float4 hi, lo;
          hi.x = tf32_rn(v.x); lo.x = tf32_rn(v.x - hi.x);
          hi.y = tf32_rn(v.y); lo.y = tf32_rn(v.y - hi.y);
          hi.z = tf32_rn(v.z); lo.z = tf32_rn(v.z - hi.z);
          hi.w = tf32_rn(v.w); lo.w = tf32_rn(v.w - hi.w);
          const uint32_t off = swz(r, ch);
          *reinterpret_cast<float4*>(sAhi + off) = hi;
          *reinterpret_cast<float4*>(sAlo + off) = lo;
        }
        if (stamp) TC_STAMP(use == 1 ? 10 : 14);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
        if (stamp) TC_STAMP(use == 1 ? 11 : 15);
      }
    }
  } else if (warp == WS_MMA_WARP) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int gchunk = 0, it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t d_main = tmem0 + acc * 2 * BN, d_corr = d_main + BN;
        mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
        fence_after();
        for (int c = 0; c < nchunks; ++c, ++gchunk) {
          const int s = gchunk % WS_STAGES;
          mbar_wait(&full_bar[s], (gchunk / WS_STAGES) & 1);
          if (gchunk == 0) TC_STAMP(2);
          fence_after();
          uint8_t* st = base + s * STAGE_BYTES;
          const uint64_t dAh = make_desc(smem_u32(st)), dAl = make_desc(smem_u32(st + A_BYTES));
          const uint64_t dBh = make_desc(smem_u32(st + 2 * A_BYTES)), dBl = make_desc(smem_u32(st + 2 * A_BYTES + B_BYTES));
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);
            const uint32_t first = (c | kk) ? 1u : 0u;
            mma_tf32(d_main, dAh + adv, dBh + adv, idesc, first);
            if (p.mode != 1) {
              mma_tf32(d_corr, dAl + adv, dBh + adv, idesc, first);
              mma_tf32(d_corr, dAh + adv, dBl + adv, idesc, 1u);
            }
          }
          mma_commit(&empty_bar[s]);
        }
        mma_commit(&tfull_bar[acc]);
        TC_STAMP(3);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue =================
    const int quarter = warp & 3;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const int model = t / (m_tiles * n_tiles);
      const int rem = t % (m_tiles * n_tiles);
      const int m0 = (rem / n_tiles) * BM, n0 = (rem % n_tiles) * BN;
      const uint32_t d_main = tmem0 + acc * 2 * BN + ((uint32_t)(quarter * 32) << 16);
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      if (it == 0 && warp == WS_EPI_WARP0 && lane == 0) TC_STAMP(4);
      fence_after();
      const float* __restrict__ bias = (EPI == 1 || EPI == 4) ? g.bias + (long long)model * g.sBias : nullptr;
      const float* __restrict__ aux = EPI == 2 ? g.aux + (long long)model * g.sAux : nullptr;
      float* __restrict__ C = g.C + (long long)model * g.sC;
      const int r = m0 + quarter * 32 + lane;   // TMEM lane == output row: each thread owns one row
      float* __restrict__ C2 = EPI == 4 ? g.C2 + (long long)model * g.sC : nullptr;
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 32) {
        float v[32], vc[32];
        tmem_ld32(d_main + (uint32_t)cb, v);
        if (p.mode != 1) {
          tmem_ld32(d_main + (uint32_t)(BN + cb), vc);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += vc[j];
        }
        if (r < g.M) {
          float* crow = C + (long long)r * g.ldc + n0 + cb;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (EPI == 1 || EPI == 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n0 + cb + j));
              o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            } else if (EPI == 2) {
              const float4 x = *reinterpret_cast<const float4*>(aux + (long long)r * g.ldaux + n0 + cb + j);
              o.x *= dswishf_(x.x); o.y *= dswishf_(x.y); o.z *= dswishf_(x.z); o.w *= dswishf_(x.w);
            } else if (EPI == 3) {
              const float4 x = *reinterpret_cast<const float4*>(crow + j);
              o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
            }
            *reinterpret_cast<float4*>(crow + j) = o;
            if (EPI == 4) {   // activation for the next GEMM, computed once per element here
              const float4 a = make_float4(swishf_(o.x), swishf_(o.y), swishf_(o.z), swishf_(o.w));
              *reinterpret_cast<float4*>(C2 + (long long)r * g.ldc + n0 + cb + j) = a;
            }
          }
        }
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (warp == WS_EPI_WARP0 && lane == 0) TC_STAMP(5);
    }
  }
  fence_before();
  __syncthreads();
  if (tid == 0) TC_STAMP(6);
  if (warp == WS_MMA_WARP) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "n"(4 * BN));
  }
}

// ------------------------------------------------------------------------------------------
// TMA-fed version (AMODE 0: A used as stored).  The LSU pipe was the busiest unit of the kernel above
// (every operand byte crossed it twice: LDG/LDGSTS in, STS out), so here the three bulk operands of a
// chunk -- A as stored (fp32), B_hi, B_lo -- are fetched by ONE thread with cp.async.bulk.tensor (TMA,
// hardware SWIZZLE_128B, zero fill past M) and never touch registers.  The tensor core reads the fp32 A
// tile as TF32 (low 13 mantissa bits ignored) = A_hi; four "splitter" warps derive A_lo = rn_tf32(A -
// trunc_tf32(A)) in place-compatible layout (same byte offset in a second buffer, so no swizzle math).
//   warp 0      TMA producer (one lane)            warps 1-4  splitters
//   warp 5      MMA issuer (one lane)              warps 6-9  epilogue (TMEM lane quarter = warp & 3)
// Barriers per ring stage: empty (tcgen05.commit), araw (A landed -> splitters), full (B landed + 4 splitter
// arrivals -> MMA).
// ------------------------------------------------------------------------------------------
constexpr int TM_STAGES = 3;
constexpr int TM_SPLIT_WARP0 = 1, TM_SPLIT_WARPS = 4;
constexpr int TM_MMA_WARP = TM_SPLIT_WARP0 + TM_SPLIT_WARPS;
constexpr int TM_EPI_WARP0 = TM_MMA_WARP + 1;
constexpr int TM_THREADS = (TM_EPI_WARP0 + 4) * 32;

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

template <int BN, int EPI>
__global__ void __launch_bounds__(TM_THREADS, 1) gemm_tc_tma_kernel(TcArgs p, const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmBh,
                                                                    const __grid_constant__ CUtensorMap tmBl,
                                                                    const __grid_constant__ CUtensorMap tmC,
                                                                    const __grid_constant__ CUtensorMap tmC2,
                                                                    const __grid_constant__ CUtensorMap tmAux) {
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the shared array (an integer round trip would lose the shared address space: generic LD/ST)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  __shared__ __align__(8) uint64_t full_bar[TM_STAGES], empty_bar[TM_STAGES], araw_bar[TM_STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ __align__(8) uint64_t aux_bar[4];
  __shared__ uint32_t tmem_base_s;

  const GemmArgs& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // (a 2-CTA-cluster variant that multicast the weight tile was measured slower and removed: profiles/r2_notes.md section 4)
  const int wid = blockIdx.x, nw = gridDim.x;                    // persistent CTAs stride over the work items
  if (tid == 0) TC_STAMP(0);
  if (warp == TM_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(4 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBl)) : "memory");
    for (int s = 0; s < TM_STAGES; ++s) {
      mbar_init(&full_bar[s], 1 + TM_SPLIT_WARPS);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&araw_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
    for (int a = 0; a < 4; ++a) mbar_init(&aux_bar[a], 1);
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem0 = tmem_base_s;
  if (tid == 0) TC_STAMP(1);

  // work items: (model, m-tile, n-tile)
  const int m_tiles = (g.M + BM - 1) / BM, n_tiles = g.N / BN;
  const int total = m_tiles * n_tiles * p.n_models;
  const int nchunks = g.K / BK;
  constexpr uint32_t idesc = make_idesc(BM, BN);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int gchunk = 0;
      for (int t = wid; t < total; t += nw) {
        const int model = t / (m_tiles * n_tiles);
        const int rem = t % (m_tiles * n_tiles);
        const int m0 = (rem / n_tiles) * BM, n0 = (rem % n_tiles) * BN;
        for (int c = 0; c < nchunks; ++c, ++gchunk) {
          const int s = gchunk % TM_STAGES;
          mbar_wait(&empty_bar[s], ((gchunk / TM_STAGES) & 1) ^ 1);
          uint8_t* st = base + s * STAGE_BYTES;
          mbar_expect_tx(&araw_bar[s], A_BYTES);
          tma_load_3d(st, &tmA, &araw_bar[s], c * BK, m0, model);
          mbar_expect_tx(&full_bar[s], 2 * B_BYTES);   // bytes that will land in THIS CTA's stage
          tma_load_3d(st + 2 * A_BYTES, &tmBh, &full_bar[s], c * BK, n0, model);
          tma_load_3d(st + 2 * A_BYTES + B_BYTES, &tmBl, &full_bar[s], c * BK, n0, model);
        }
      }
    }
    __syncwarp();
  } else if (warp < TM_MMA_WARP) {
    // ================= splitters: A_lo = rn_tf32(A - trunc_tf32(A)) =================
    const int st_tid = tid - TM_SPLIT_WARP0 * 32;   // 0..127
    int gchunk = 0;
    for (int t = wid; t < total; t += nw) {
      for (int c = 0; c < nchunks; ++c, ++gchunk) {
        const int s = gchunk % TM_STAGES;
        mbar_wait(&araw_bar[s], (gchunk / TM_STAGES) & 1);
        const uint8_t* raw = base + s * STAGE_BYTES;
        uint8_t* lo = base + s * STAGE_BYTES + A_BYTES;
        float4 x[A_BYTES / 16 / (TM_SPLIT_WARPS * 32)];
#pragma unroll
        for (int u = 0; u < A_BYTES / 16 / (TM_SPLIT_WARPS * 32); ++u)
          x[u] = *reinterpret_cast<const float4*>(raw + (st_tid + u * TM_SPLIT_WARPS * 32) * 16);
#pragma unroll
        for (int u = 0; u < A_BYTES / 16 / (TM_SPLIT_WARPS * 32); ++u) {
          float4 v = x[u], l;
          l.x = tf32_rn(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
          l.y = tf32_rn(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
          l.z = tf32_rn(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
          l.w = tf32_rn(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
          *reinterpret_cast<float4*>(lo + (st_tid + u * TM_SPLIT_WARPS * 32) * 16) = l;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
      }
    }
  } else if (warp == TM_MMA_WARP) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int gchunk = 0, it = 0;
      for (int t = wid; t < total; t += nw, ++it) {
        const int acc = it & 1;
        const uint32_t d_main = tmem0 + acc * 2 * BN, d_corr = d_main + BN;
        mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
        fence_after();
        for (int c = 0; c < nchunks; ++c, ++gchunk) {
          const int s = gchunk % TM_STAGES;
          mbar_wait(&araw_bar[s], (gchunk / TM_STAGES) & 1);
          mbar_wait(&full_bar[s], (gchunk / TM_STAGES) & 1);
          if (gchunk == 0) TC_STAMP(2);
          fence_after();
          uint8_t* st = base + s * STAGE_BYTES;
          const uint64_t dAh = make_desc(smem_u32(st)), dAl = make_desc(smem_u32(st + A_BYTES));
          const uint64_t dBh = make_desc(smem_u32(st + 2 * A_BYTES)), dBl = make_desc(smem_u32(st + 2 * A_BYTES + B_BYTES));
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);
            const uint32_t first = (c | kk) ? 1u : 0u;
            mma_tf32(d_main, dAh + adv, dBh + adv, idesc, first);
            if (p.mode != 1) {
              mma_tf32(d_corr, dAl + adv, dBh + adv, idesc, first);
              mma_tf32(d_corr, dAh + adv, dBl + adv, idesc, 1u);
            }
          }
          mma_commit(&empty_bar[s]);
        }
        mma_commit(&tfull_bar[acc]);
        TC_STAMP(3);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue =================
    // Each warp owns 32 accumulator rows (TMEM lane quarter).  A 32x32 block goes TMEM -> registers (thread =
    // row) -> fused op -> a 4 KB SWIZZLE_128B staging tile -> TMA store (reduce-add for EPI 3), so global
    // memory only ever sees full 128-byte lines and the LSU pipe carries no global traffic; the dswish
    // operand of EPI 2 comes in the same way (TMA load of the matching 32x32 block).  Rows past M are
    // clipped by the tensor map.
    const int quarter = warp & 3;
    uint8_t* obuf = base + TM_STAGES * STAGE_BYTES + quarter * 8192;   // 4 KB C tile + 4 KB C2 / aux tile
    uint64_t* abar = &aux_bar[quarter];
    uint32_t aux_uses = 0;
    int it = 0;
    for (int t = wid; t < total; t += nw, ++it) {
      const int acc = it & 1;
      const int model = t / (m_tiles * n_tiles);
      const int rem = t % (m_tiles * n_tiles);
      const int m0 = (rem / n_tiles) * BM, n0 = (rem % n_tiles) * BN;
      const uint32_t d_main = tmem0 + acc * 2 * BN + ((uint32_t)(quarter * 32) << 16);
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      if (it == 0 && quarter == 0 && lane == 0) TC_STAMP(4);
      fence_after();
      const float* __restrict__ bias = (EPI == 1 || EPI == 4) ? g.bias + (long long)model * g.sBias : nullptr;
      const int row0 = m0 + quarter * 32;
      const uint32_t rsw = (uint32_t)(lane & 7);
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 32) {
        if (EPI == 2 && lane == 0) {
          mbar_expect_tx(abar, 4096);
          tma_load_3d(obuf + 4096, &tmAux, abar, n0 + cb, row0, model);
        }
        float v[32], vc[32];
        tmem_ld32(d_main + (uint32_t)cb, v);
        if (p.mode != 1) {
          tmem_ld32(d_main + (uint32_t)(BN + cb), vc);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += vc[j];
        }
        if (EPI == 1 || EPI == 4) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n0 + cb + j));
            v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
          }
        }
        if (EPI == 2) {
          mbar_wait(abar, aux_uses & 1);
          ++aux_uses;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 x = *reinterpret_cast<const float4*>(obuf + 4096 + lane * 128 + ((j ^ rsw) << 4));
            v[4 * j] *= dswishf_(x.x); v[4 * j + 1] *= dswishf_(x.y); v[4 * j + 2] *= dswishf_(x.z); v[4 * j + 3] *= dswishf_(x.w);
          }
        }
        // the TMA store that last used this staging tile must be done READING it; with one output per
        // block the two 4 KB halves alternate, so only the store before the previous one is waited for
        constexpr bool kTwoTiles = (EPI == 2 || EPI == 4);
        uint8_t* ctile = kTwoTiles ? obuf : obuf + ((cb >> 5) & 1) * 4096;
        if (lane == 0) {
          if (kTwoTiles) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncwarp();
        if (p.mode != 3) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            *reinterpret_cast<float4*>(ctile + lane * 128 + ((j ^ rsw) << 4)) = o;
            if (EPI == 4)   // activation for the next GEMM, computed once per element here
              *reinterpret_cast<float4*>(obuf + 4096 + lane * 128 + ((j ^ rsw) << 4)) =
                  make_float4(swishf_(o.x), swishf_(o.y), swishf_(o.z), swishf_(o.w));
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (EPI == 3)
              asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&tmC)), "r"(smem_u32(ctile)), "r"(n0 + cb), "r"(row0), "r"(model) : "memory");
            else
              asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&tmC)), "r"(smem_u32(ctile)), "r"(n0 + cb), "r"(row0), "r"(model) : "memory");
            if (EPI == 4)
              asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&tmC2)), "r"(smem_u32(obuf + 4096)), "r"(n0 + cb), "r"(row0), "r"(model) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (quarter == 0 && lane == 0) TC_STAMP(5);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  fence_before();
  __syncthreads();
  if (tid == 0) TC_STAMP(6);
  if (warp == TM_MMA_WARP) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "n"(4 * BN));
  }
}

template <int BN>
constexpr size_t ws_smem_bytes() { return (size_t)WS_STAGES * (2 * BM * BK * 4 + 2 * BN * BK * 4) + 1024; }
template <int BN>
constexpr size_t tma_smem_bytes() { return (size_t)TM_STAGES * (2 * BM * BK * 4 + 2 * BN * BK * 4) + 4 * 8192 + 1024; }

}  // namespace tc

// ---- TMA tensor maps (host).  cuTensorMapEncodeTiled is fetched through the runtime so that the library
// does not link libcuda; maps are cached by (pointer, shape, strides): workspace and weight pointers are
// stable across evaluations, so steady state does no encoding at all.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct TmapKey {
  const void* ptr; long long d0, d1, d2, s1, s2; int b1;
  bool operator==(const TmapKey& o) const { return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s1 == o.s1 && s2 == o.s2 && b1 == o.b1; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    unsigned long long h = reinterpret_cast<uintptr_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    for (long long v : {k.d0, k.d1, k.d2, k.s1, k.s2, (long long)k.b1}) h = (h ^ (unsigned long long)v) * 0x100000001B3ull;
    return (size_t)(h ^ (h >> 29));
  }
};
// [d2][d1][d0] fp32 tensor, d0 contiguous, row stride s1 floats, slab stride s2 floats; box = 32 x b1 x 1, SWIZZLE_128B
struct TmapCache {
  std::unordered_map<TmapKey, CUtensorMap*, TmapKeyHash> map;
  std::mutex mu;
};
inline TmapCache& tmap_cache() { static TmapCache c; return c; }
// A map is copied into the kernel parameters by the launch that asked for it, so old entries can go at
// any time BETWEEN launches (batch shapes change every MC step, pointers follow): called once per launch.
inline void tmap_cache_trim() {
  TmapCache& c = tmap_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  if (c.map.size() > 8192) {
    for (auto& e : c.map) delete e.second;
    c.map.clear();
  }
}
inline const CUtensorMap* get_tmap(const float* ptr, long long d0, long long d1, long long d2, long long s1, long long s2, int b1) {
  static EncodeTiledFn encode = nullptr;
  static bool tried = false;
  auto& cache = tmap_cache().map;
  std::lock_guard<std::mutex> lock(tmap_cache().mu);
  if (!tried) {
    tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  if (!encode) return nullptr;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (s1 & 3) || (s2 & 3)) return nullptr;
  const TmapKey key{ptr, d0, d1, d2, s1, s2, b1};
  const auto hit = cache.find(key);
  if (hit != cache.end()) return hit->second;
  CUtensorMap* m = new CUtensorMap;
  const cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  const cuuint64_t strides[2] = {(cuuint64_t)s1 * 4, (cuuint64_t)(d2 > 1 ? s2 : s1 * d1) * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)b1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { delete m; return nullptr; }
  cache.emplace(key, m);
  return m;
}

template <int BN, int EPI>
int launch_gemm_tma(const GemmArgs& g, const float* Bhi, const float* Blo, long long sBw, int n_models, cudaStream_t st, bool* done) {
  *done = false;
  tmap_cache_trim();
  const CUtensorMap* mA = get_tmap(g.A, g.K, g.M, n_models, g.lda, g.sA, tc::BM);
  const CUtensorMap* mBh = get_tmap(Bhi, g.K, g.N, n_models, g.K, sBw, BN);
  const CUtensorMap* mBl = get_tmap(Blo, g.K, g.N, n_models, g.K, sBw, BN);
  const CUtensorMap* mC = get_tmap(g.C, g.N, g.M, n_models, g.ldc, g.sC, 32);
  const CUtensorMap* mC2 = EPI == 4 ? get_tmap(g.C2, g.N, g.M, n_models, g.ldc, g.sC, 32) : mC;
  const CUtensorMap* mAux = EPI == 2 ? get_tmap(g.aux, g.N, g.M, n_models, g.ldaux, g.sAux, 32) : mC;
  if (!mA || !mBh || !mBl || !mC || !mC2 || !mAux) return 0;
  tc::TcArgs p{g, Bhi, Blo, sBw, n_models, 0, nullptr};
  const int total = ceil_div(g.M, tc::BM) * (g.N / BN) * n_models;   // work items
  constexpr size_t smem = tc::tma_smem_bytes<BN>();
  static bool configured[64] = {};   // the attribute is per device
  int dev = 0;
  VSSR_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    VSSR_CUDA(cudaFuncSetAttribute(tc::gemm_tc_tma_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev & 63] = true;
  }
  const int grid = total < 148 ? total : 148;   // one persistent CTA per SM
  VSSR_PROF(VSSR_K_GEMM, st, (tc::gemm_tc_tma_kernel<BN, EPI><<<grid, tc::TM_THREADS, smem, st>>>(p, *mA, *mBh, *mBl, *mC, *mC2, *mAux)));
  *done = true;
  return 0;
}

// AMODE 0: operands as stored -> TMA kernel.  AMODE 2 (A transformed on load: dswish(h5)*w6 of the readout backward):
// the cp.async-fed tcgen05 kernel.
template <int BN, int AMODE, int EPI>
int launch_gemm_tc(const GemmArgs& g, const float* Bhi, const float* Blo, long long sBw, int n_models, cudaStream_t st) {
  if (AMODE == 0) {
    bool done = false;
    const int rc = launch_gemm_tma<BN, EPI>(g, Bhi, Blo, sBw, n_models, st, &done);
    if (rc || done) return rc;
  }
  tc::TcArgs p{g, Bhi, Blo, sBw, n_models, 0, nullptr};
  const int total = ceil_div(g.M, tc::BM) * (g.N / BN) * n_models;
  constexpr size_t smem = tc::ws_smem_bytes<BN>();
  static bool configured[64] = {};
  int dev = 0;
  VSSR_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    VSSR_CUDA(cudaFuncSetAttribute(tc::gemm_tc_ws_kernel<BN, AMODE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev & 63] = true;
  }
  const int grid = total < 148 ? total : 148;
  VSSR_PROF(VSSR_K_GEMM, st, (tc::gemm_tc_ws_kernel<BN, AMODE, EPI><<<grid, tc::WS_THREADS, smem, st>>>(p)));
  return 0;
}
