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
// Several CTAs are resident per SM (64 KB smem, 128 TMEM columns each); one CTA's MMAs and epilogue
// overlap another's operand staging.  Tiles are walked persistently (tile = blockIdx.x + k*gridDim.x).
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
};

template <int BN, int AMODE, int EPI>
__global__ void __launch_bounds__(THREADS) gemm_tc_kernel(TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte aligned operand tiles
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sAhi = base;
  uint8_t* sAlo = sAhi + BM * BK * 4;
  uint8_t* sBhi = sAlo + BM * BK * 4;
  uint8_t* sBlo = sBhi + BN * BK * 4;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  const GemmArgs& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_d = tmem_base_s;
  uint32_t phase = 0;

  const int m_tiles = (g.M + BM - 1) / BM, n_tiles = g.N / BN;
  const int total = m_tiles * n_tiles * p.n_models;
  const int nchunks = g.K / BK;
  constexpr uint32_t idesc = make_idesc(BM, BN);

  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int model = t / (m_tiles * n_tiles);
    const int rem = t % (m_tiles * n_tiles);
    const int m0 = (rem / n_tiles) * BM, n0 = (rem % n_tiles) * BN;
    const float* __restrict__ A = g.A + (long long)model * g.sA;
    const float* __restrict__ Bh = p.Bhi + (long long)model * p.sBw + (long long)n0 * g.K;
    const float* __restrict__ Bl = p.Blo + (long long)model * p.sBw + (long long)n0 * g.K;
    const float* __restrict__ avec = AMODE == 2 ? g.avec + (long long)model * g.sAvec : nullptr;

    for (int c = 0; c < nchunks; ++c) {
      const int k0 = c * BK;
      // ---- stage A (transform + hi/lo split) ----
#pragma unroll
      for (int q = 0; q < (BM * 8) / THREADS; ++q) {
        const int idx = tid + q * THREADS;
        const int r = idx >> 3, ch = idx & 7;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + r < g.M) {
          x = __ldg(reinterpret_cast<const float4*>(A + (long long)(m0 + r) * g.lda + k0) + ch);
          if (AMODE == 1) {
            x.x = swishf_(x.x); x.y = swishf_(x.y); x.z = swishf_(x.z); x.w = swishf_(x.w);
          } else if (AMODE == 2) {
            const float4 av = __ldg(reinterpret_cast<const float4*>(avec + k0) + ch);
            x.x = dswishf_(x.x) * av.x; x.y = dswishf_(x.y) * av.y; x.z = dswishf_(x.z) * av.z; x.w = dswishf_(x.w) * av.w;
          }
        }
        // hi = RN_tf32(x), lo = RN_tf32(x - hi): both exactly representable, so the tensor core's
        // operand truncation is a no-op and the split error is unbiased (~2^-22 relative)
        float4 hi, lo;
        hi.x = tf32_rn(x.x); lo.x = tf32_rn(x.x - hi.x);
        hi.y = tf32_rn(x.y); lo.y = tf32_rn(x.y - hi.y);
        hi.z = tf32_rn(x.z); lo.z = tf32_rn(x.z - hi.z);
        hi.w = tf32_rn(x.w); lo.w = tf32_rn(x.w - hi.w);
        const uint32_t off = swz(r, ch);
        *reinterpret_cast<float4*>(sAhi + off) = hi;
        *reinterpret_cast<float4*>(sAlo + off) = lo;
      }
      // ---- stage B (pre-split weights, K-major) ----
#pragma unroll
      for (int q = 0; q < (BN * 8) / THREADS; ++q) {
        const int idx = tid + q * THREADS;
        const int r = idx >> 3, ch = idx & 7;
        const float4 h = __ldg(reinterpret_cast<const float4*>(Bh + (long long)r * g.K + k0) + ch);
        const float4 l = __ldg(reinterpret_cast<const float4*>(Bl + (long long)r * g.K + k0) + ch);
        const uint32_t off = swz(r, ch);
        *reinterpret_cast<float4*>(sBhi + off) = h;
        *reinterpret_cast<float4*>(sBlo + off) = l;
      }
      fence_async_smem();      // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      __syncthreads();
      if (tid == 0) {
        fence_after();
        const uint64_t dAh = make_desc(smem_u32(sAhi)), dAl = make_desc(smem_u32(sAlo));
        const uint64_t dBh = make_desc(smem_u32(sBhi)), dBl = make_desc(smem_u32(sBlo));
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint64_t adv = (uint64_t)(kk * 32 >> 4);   // 8 fp32 = 32 bytes along K inside the swizzle row
          // main product into columns [0,BN); the two small correction products into their own
          // accumulator [BN,2BN): the tensor core truncates on every accumulate, so keeping the
          // 2^-11-sized terms away from the large running sum cuts that bias 3x
          const uint32_t first = (c | kk) ? 1u : 0u;
          mma_tf32(tmem_d, dAh + adv, dBh + adv, idesc, first);
          if (p.mode != 1) {          // mode 1 = plain TF32 (accuracy studies only)
            mma_tf32(tmem_d + BN, dAl + adv, dBh + adv, idesc, first);
            mma_tf32(tmem_d + BN, dAh + adv, dBl + adv, idesc, 1u);
          }
        }
        mma_commit(&bar);
      }
      mbar_wait(&bar, phase);   // MMAs of this chunk retired: smem reusable, accumulator current
      phase ^= 1;
    }
    fence_after();

    // ---- epilogue: TMEM -> registers -> fused op -> global; thread owns row m0 + 32*warp + lane ----
    const int r = m0 + warp * 32 + lane;
    const float* __restrict__ bias = EPI == 1 ? g.bias + (long long)model * g.sBias : nullptr;
    const float* __restrict__ aux = EPI == 2 ? g.aux + (long long)model * g.sAux : nullptr;
    float* __restrict__ C = g.C + (long long)model * g.sC;
#pragma unroll 1
    for (int cb = 0; cb < BN; cb += 32) {
      float v[32], vc[32];
      tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, v);
      if (p.mode != 1) {
        tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(BN + cb), vc);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += vc[j];
      }
      if (r < g.M) {
        float* crow = C + (long long)r * g.ldc + n0 + cb;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (EPI == 1) {
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
        }
      }
    }
    fence_before();
    __syncthreads();   // every warp has drained its TMEM lanes before the next tile overwrites them
    fence_after();
  }

  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(2 * BN));
  }
}

template <int BN>
constexpr size_t smem_bytes() { return (size_t)(2 * BM * BK * 4 + 2 * BN * BK * 4) + 1024; }

}  // namespace tc

template <int BN, int AMODE, int EPI>
int launch_gemm_tc(const GemmArgs& g, const float* Bhi, const float* Blo, long long sBw, int n_models, cudaStream_t st) {
  static bool configured = false;
  constexpr size_t smem = tc::smem_bytes<BN>();
  if (!configured) {
    VSSR_CUDA(cudaFuncSetAttribute(tc::gemm_tc_kernel<BN, AMODE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("VSSR_TC_MODE"); mode = e ? atoi(e) : 0; }
  tc::TcArgs p{g, Bhi, Blo, sBw, n_models, mode};
  const int total = ceil_div(g.M, tc::BM) * (g.N / BN) * n_models;
  const int grid = total < 148 * 2 ? total : 148 * 2;   // 2 CTAs/SM: 256 TMEM columns each
  VSSR_PROF(VSSR_K_GEMM, st, (tc::gemm_tc_kernel<BN, AMODE, EPI><<<grid, tc::THREADS, smem, st>>>(p)));
  return 0;
}
