// C[M,N] = epi(bf16(A[M,K] . W[N,K]^T + bias)) (+ residual) on the 5th-gen tensor
// cores: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared-memory ring ->
// tcgen05.mma (one elected thread, fp32 accumulators in TMEM) -> tcgen05.ld
// epilogue.  Warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer,
// warp 2 = TMEM allocator, warps 4..7 = epilogue (one TMEM lane quarter each).
//
// Replaces nn.Linear / Conv3d(kernel==stride) / Embedding.as_linear of the
// reference (models/qwen2_vl/vision.py:83-102,108-119,132-133,168-169;
// language.py:52-55; mlp.py:9-15).  Rounding points follow
// oracle/mlx_semantics.py::linear (+ activations, residual add).
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace b200 {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;

struct GemmParams {
  const bf16* bias;
  const bf16* residual;
  bf16* C;
  long ldc, ldr;
  int M, N, K;
  int epilogue;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// K-major operand tile, 128B swizzle: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address [0,14)
  d |= (uint64_t)1 << 16;                       // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // SBO = 1024 B
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN>
struct GemmSmem {
  static constexpr int STAGES = (BN == 128) ? 6 : 8;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int RING = STAGES * (A_BYTES + B_BYTES);
  static constexpr int TOTAL = RING + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmParams p) {
  using L = GemmSmem<BN>;
  constexpr int STAGES = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * L::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::RING);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x, m_blk = blockIdx.y;
  const int num_k = (p.K + BK - 1) / BK;

  // The producer thread initialises the barriers itself and puts the first STAGES loads in
  // flight BEFORE the CTA-wide setup barrier: the first TMA latency (cold weights: ~1.5 us)
  // overlaps the TMEM allocation instead of following it.
  const int pre_k = num_k < STAGES ? num_k : STAGES;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int k = 0; k < pre_k; ++k) {
      mbar_expect_tx(&full_bar[k], L::A_BYTES + L::B_BYTES);
      tma_load_2d(sA + k * L::A_BYTES, &tmA, &full_bar[k], k * BK, m_blk * BM);
      tma_load_2d(sB + k * L::B_BYTES, &tmB, &full_bar[k], k * BK, n_blk * BN);
    }
  }
  if (warp == 2) {  // whole warp: tcgen05.alloc is .sync.aligned
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer (the first pre_k stages are already in flight) =====
      for (int k = pre_k; k < num_k; ++k) {
        const int s = k % STAGES;
        const uint32_t ph = (k / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], L::A_BYTES + L::B_BYTES);
        tma_load_2d(sA + s * L::A_BYTES, &tmA, &full_bar[s], k * BK, m_blk * BM);
        tma_load_2d(sB + s * L::B_BYTES, &tmB, &full_bar[s], k * BK, n_blk * BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      for (int k = 0; k < num_k; ++k) {
        const int s = k % STAGES;
        const uint32_t ph = (k / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA + s * L::A_BYTES);
        const uint32_t b0 = smem_u32(sB + s * L::B_BYTES);
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
          const uint64_t ad = make_smem_desc(a0 + kk * UMMA_K * 2);
          const uint64_t bd = make_smem_desc(b0 + kk * UMMA_K * 2);
          umma_bf16(tmem_base, ad, bd, idesc, (k > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
      }
      umma_commit(tmem_full);  // accumulator complete
    }
  }
  __syncwarp();
  {
    // ===== epilogue: TMEM -> registers -> bias/activation/residual -> global =====
    // ALL 8 warps take part once their main-loop role is done (the producer / MMA threads have
    // issued everything by now): warp w reads TMEM lane quarter w % 4 and column half w / 4, so
    // the non-overlapped tail of a CTA (bias, GELU, residual, 16-byte stores) is half as long
    // (ncu, round 1: a 128x128 tile's epilogue is 3-5 us of a 13-24 us CTA).
    const int q = warp & 3;  // TMEM lane quarter == warp % 4
    const int chalf = warp >> 2;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int row = m_blk * BM + q * 32 + lane;
    const bool row_ok = row < p.M;
    const bool vec_ok = ((p.ldc & 7) == 0) && (!p.residual || (p.ldr & 7) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll 1
    for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
      const int n0 = n_blk * BN + c0;
      if (!row_ok || n0 >= p.N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
      const bool full = (n0 + 32 <= p.N);
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (full || n0 + j < p.N) v[j] += bf2f(p.bias[n0 + j]);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = rbf(v[j]);
      if (p.epilogue == B200_EPI_GELU_FAST) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_fast_bf(v[j]);
      } else if (p.epilogue == B200_EPI_GELU_EXACT) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_exact_bf(v[j]);
      }
      bf16* crow = p.C + (long)row * p.ldc + n0;
      if (full && vec_ok) {
        if (p.residual) {
          const uint4* rr = reinterpret_cast<const uint4*>(p.residual + (long)row * p.ldr + n0);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float rf[8];
            unpack8(rr[g], rf);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[g * 8 + j] = rbf(rf[j] + v[g * 8 + j]);
          }
        }
        uint4* cw = reinterpret_cast<uint4*>(crow);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack2(v[g * 8 + 0], v[g * 8 + 1]);
          o.y = pack2(v[g * 8 + 2], v[g * 8 + 3]);
          o.z = pack2(v[g * 8 + 4], v[g * 8 + 5]);
          o.w = pack2(v[g * 8 + 6], v[g * 8 + 7]);
          cw[g] = o;
        }
      } else {
        for (int j = 0; j < 32 && n0 + j < p.N; ++j) {
          float o = v[j];
          if (p.residual) o = rbf(bf2f(p.residual[(long)row * p.ldr + n0 + j]) + o);
          crow[j] = f2bf(o);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)BN)
                 : "memory");
  }
}

// ---- host: tensor maps ------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  long ld;
  int rows, k, box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && ld == o.ld && rows == o.rows && k == o.k && box_rows == o.box_rows;
  }
};
struct TmapHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h = h * 1000003u ^ std::hash<long>()(k.ld);
    h = h * 1000003u ^ (size_t)k.rows;
    h = h * 1000003u ^ (size_t)k.k;
    h = h * 1000003u ^ (size_t)k.box_rows;
    return h;
  }
};

// 2-D K-major bf16 operand (rows x K, row pitch ld elements), box = box_rows x 64.
static int get_tmap(const void* ptr, long ld, int rows, int k, int box_rows, CUtensorMap* out) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapHash> cache;
  static std::mutex mu;
  TmapKey key{ptr, ld, rows, k, box_rows};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return B200_OK;
    }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return B200_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) ptr=%p ld=%ld rows=%d k=%d", (int)r, ptr, ld,
              rows, k);
    return B200_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 8192) cache.clear();
    cache[key] = tm;
  }
  *out = tm;
  return B200_OK;
}

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       cudaStream_t st) {
  static unsigned long long attr_mask = 0ull;
  if (first_use_on_device(&attr_mask)) {
    B200_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   GemmSmem<BN>::TOTAL));
  }
  dim3 grid(cdiv(p.N, BN), cdiv(p.M, BM));
  gemm_tn_kernel<BN><<<grid, 256, GemmSmem<BN>::TOTAL, st>>>(ta, tb, p);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int gemm_bf16_tn(const void* A, long lda, const void* W, const void* bias, const void* residual,
                 long ldr, void* C, long ldc, int M, int N, int K, int epilogue,
                 cudaStream_t st) {
  B200_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  B200_REQUIRE((lda % 8) == 0 && (K % 8) == 0, "gemm: lda (%ld) and K (%d) must be multiples of 8",
               lda, K);
  B200_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0,
               "gemm: A and W must be 16-byte aligned");
  B200_REQUIRE(epilogue >= 0 && epilogue <= 2, "gemm: bad epilogue %d", epilogue);
  // pick the N tile so that the grid covers the 148 SMs when M is small
  const long tiles128 = (long)cdiv(M, BM) * cdiv(N, 128);
  const int bn = (tiles128 >= 120 && (N % 128) == 0) ? 128 : 64;
  CUtensorMap ta, tb;
  int rc = get_tmap(A, lda, M, K, BM, &ta);
  if (rc) return rc;
  rc = get_tmap(W, (long)K, N, K, bn, &tb);
  if (rc) return rc;
  GemmParams p{(const bf16*)bias, (const bf16*)residual, (bf16*)C, ldc, ldr, M, N, K, epilogue};
  return bn == 128 ? launch_gemm<128>(ta, tb, p, st) : launch_gemm<64>(ta, tb, p, st);
}

}  // namespace b200

extern "C" int b200_gemm_bf16_tn(const void* A, long lda, const void* W, const void* bias,
                                 const void* residual, long ldr, void* C, long ldc, int M, int N,
                                 int K, int epilogue, void* stream) {
  return b200::gemm_bf16_tn(A, lda, W, bias, residual, ldr, C, ldc, M, N, K, epilogue,
                            (cudaStream_t)stream);
}
