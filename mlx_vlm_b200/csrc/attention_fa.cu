// Prefill / vision attention, pipelined: TMA-fed, warp-specialised, tcgen05 with double-buffered
// score tiles in TMEM.  Same arithmetic and rounding points as attention_tc.cu (the reference's
// mlx-CPU SDPA: oracle/mlx_semantics.py::sdpa; models/base.py:305-373, qwen2_vl/vision.py:154):
//     qs = bf16(q * bf16(scale))       -- done by the rotary kernels that already touch q
//     s  = bf16(qs . k^T);  p = bf16(softmax_fp32(s));  o = bf16(p . v)
// p is rounded AFTER normalisation with the final row max / sum, so the kernel is two-pass over
// the keys and the second pass recomputes the score tile on the tensor cores.
//
// attention_tc.cu (round 1) staged every tile with the compute threads and ran
// load -> sync -> MMA -> wait -> softmax strictly in sequence (61 us per ViT layer, 2 % of the
// tensor peak).  Here, per CTA = 128 query rows of one head:
//   warp 0      TMA producer: Q once, K tiles for both passes, V^T tiles for pass 2 (3-D tensor
//               maps over the packed qkv buffer / the KV cache, 128B swizzle, zero-filled tails)
//   warp 1      MMA issuer:  S[i % 2] = Qs . K^T  (M128 x N128, TMEM cols 0..255)
//                            O += P . V           (M128 x N=hd, TMEM cols 256..)
//               the P.V of tile j is issued AFTER the Q.K^T of tile j+1, so the tensor core
//               computes the next scores while the softmax warps work on the current ones
//   warps 2..9  softmax: two threads per query row (64 score columns each), TMEM -> registers,
//               pass 1: running max / sum of exp;  pass 2: p -> shared memory (the A operand)
// V is consumed as V^T ([head][dim][key], keys contiguous): a K-major B operand that TMA can
// load directly; the rotary kernels emit it (rowops.cu: *_qkv_post).
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "decode.cuh"

namespace b200 {

namespace {

constexpr int FA_TQ = 128, FA_TK = 128;
constexpr int FA_BLK = 16 * 1024;       // one 128-row x 64-column bf16 operand block
constexpr int FA_THREADS = 320;         // producer warp, MMA warp, 8 softmax warps

struct FaParams {
  bf16* out;
  long o_ts;
  int n_heads, n_kv, hd, hdp, Lq, S, causal;
  int q0, k0;  // token offsets of this segment inside the tensors the maps describe
  const int4* segs;  // optional [gridDim.z]: (q0, Lq, k0, S) of segment blockIdx.z — several independent sequences of one
                     // concatenated batch in ONE launch (the tail wave of a 160-CTA launch per sequence idles half the SMs)
};

struct FaBars {
  uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], s_empty[2], p_full, p_empty,
      o_full;
};

__device__ __forceinline__ uint32_t f_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void f_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f_u32(bar)), "r"(count));
}
__device__ __forceinline__ void f_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(f_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void f_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(f_u32(bar)) : "memory");
}
__device__ __forceinline__ void f_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = f_u32(bar);
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
__device__ __forceinline__ void f_tma_3d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                         int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(f_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(f_u32(bar))
      : "memory");
}
__device__ __forceinline__ void f_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void f_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void f_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void f_umma(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void f_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(f_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t f_desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr uint32_t F_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);

__device__ __forceinline__ void f_tmem_ld32(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void f_tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void f_sbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// 2^x on the SFU (one MUFU instruction, rel. error 2^-22: far below the bf16 rounding of p that follows)
__device__ __forceinline__ float f_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(FA_THREADS, 1)
attention_fa_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const FaParams p_in) {
  FaParams p = p_in;
  if (p_in.segs) {   // static table (written before the producing kernels ran): safe to read ahead of griddepcontrol.wait
    const int4 sg = __ldg(p_in.segs + blockIdx.z);
    p.q0 = sg.x; p.Lq = sg.y; p.k0 = sg.z; p.S = sg.w;
    if ((int)blockIdx.x * FA_TQ >= p.Lq) return;   // the grid is sized for the longest segment
  }
  extern __shared__ uint8_t fa_smem_raw[];
  __shared__ FaBars bars;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 stat[2][FA_TQ];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fa_smem_raw) + 1023) &
                                           ~static_cast<uintptr_t>(1023));
  uint8_t* Qs = sm;                          // [128 q][128 d]            2 k-blocks
  uint8_t* Ks = sm + 2 * FA_BLK;             // 2 slots x [128 keys][128 d]
  uint8_t* Vs = sm + 6 * FA_BLK;             // 2 slots x 2 key-blocks x [hdp d][64 keys]
  uint8_t* Ps = sm + 10 * FA_BLK;            // [128 q][128 keys]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, kvh = h / (p.n_heads / p.n_kv);
  const int row0 = blockIdx.x * FA_TQ;
  const int S = p.S;
  const int q_last = min(row0 + FA_TQ, p.Lq) - 1;
  const int vis_tile = p.causal ? min(S, S - p.Lq + q_last + 1) : S;
  const int n_tiles = (vis_tile + FA_TK - 1) / FA_TK;
  const int kbq = (p.hdp + 63) >> 6;         // 64-column blocks of the head dimension
  const int vblk = p.hdp * 128;              // bytes of one 64-key block of a V^T tile

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    f_init(&bars.q_full, 1);
    for (int i = 0; i < 2; ++i) {
      f_init(&bars.k_full[i], 1);
      f_init(&bars.k_empty[i], 1);
      f_init(&bars.v_full[i], 1);
      f_init(&bars.v_empty[i], 1);
      f_init(&bars.s_full[i], 1);
      f_init(&bars.s_empty[i], 256);
    }
    f_init(&bars.p_full, 256);
    f_init(&bars.p_empty, 1);
    f_init(&bars.o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(f_u32(&tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  f_fence_before();
  __syncthreads();
  f_fence_after();
  const uint32_t tmem = tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      f_expect_tx(&bars.q_full, (uint32_t)kbq * FA_BLK);
      for (int c = 0; c < kbq; ++c) f_tma_3d(Qs + c * FA_BLK, &tmQ, &bars.q_full, c * 64, h, p.q0 + row0);
      for (int i = 0; i < 2 * n_tiles; ++i) {
        const int j = i % n_tiles, s = i & 1;
        f_wait(&bars.k_empty[s], ((i >> 1) & 1) ^ 1);
        f_expect_tx(&bars.k_full[s], (uint32_t)kbq * FA_BLK);
        for (int c = 0; c < kbq; ++c)
          f_tma_3d(Ks + s * 2 * FA_BLK + c * FA_BLK, &tmK, &bars.k_full[s], c * 64, kvh, p.k0 + j * FA_TK);
        if (i >= n_tiles) {
          const int sv = j & 1;
          f_wait(&bars.v_empty[sv], ((j >> 1) & 1) ^ 1);
          f_expect_tx(&bars.v_full[sv], 2u * (uint32_t)vblk);
          for (int c = 0; c < 2; ++c)
            f_tma_3d(Vs + sv * 2 * FA_BLK + c * vblk, &tmV, &bars.v_full[sv], p.k0 + j * FA_TK + c * 64, 0, kvh);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FA_TK >> 3) << 17) | (8u << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.hdp >> 3) << 17) | (8u << 24);
      const uint32_t q_lo = f_desc_lo(f_u32(Qs)), p_lo = f_desc_lo(f_u32(Ps));
      const int ksteps = p.hdp >> 4;
      auto do_pv = [&](int j) {
        const int sv = j & 1;
        f_wait(&bars.p_full, j & 1);
        f_wait(&bars.v_full[sv], (j >> 1) & 1);
        f_fence_after();
        const uint32_t v_lo = f_desc_lo(f_u32(Vs + sv * 2 * FA_BLK));
        for (int ks = 0; ks < FA_TK / 16; ++ks) {
          const uint32_t a = p_lo + (uint32_t)((ks >> 2) * (FA_BLK >> 4) + (ks & 3) * 2);
          const uint32_t b = v_lo + (uint32_t)((ks >> 2) * (vblk >> 4) + (ks & 3) * 2);
          f_umma(tmem + 256, a, b, F_DESC_HI, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        f_commit(&bars.p_empty);
        f_commit(&bars.v_empty[sv]);
      };
      f_wait(&bars.q_full, 0);
      for (int i = 0; i < 2 * n_tiles; ++i) {
        const int s = i & 1;
        f_wait(&bars.k_full[s], (i >> 1) & 1);
        f_wait(&bars.s_empty[s], ((i >> 1) & 1) ^ 1);
        f_fence_after();
        const uint32_t k_lo = f_desc_lo(f_u32(Ks + s * 2 * FA_BLK));
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t off = (uint32_t)((ks >> 2) * (FA_BLK >> 4) + (ks & 3) * 2);
          f_umma(tmem + (uint32_t)s * 128u, q_lo + off, k_lo + off, F_DESC_HI, idesc_s, ks > 0 ? 1u : 0u);
        }
        f_commit(&bars.k_empty[s]);
        f_commit(&bars.s_full[s]);
        if (i > n_tiles) do_pv(i - n_tiles - 1);  // P.V of the previous pass-2 tile, under this Q.K^T
      }
      do_pv(n_tiles - 1);
      f_commit(&bars.o_full);
    }
  } else {
    // ===== softmax warps: thread = (query row, score-column half) =====
    const int q4 = warp & 3, half = (warp - 2) >> 2;
    const int row = q4 * 32 + lane;
    const int qi = row0 + row;
    const int vis = (qi < p.Lq) ? (p.causal ? min(S, S - p.Lq + qi + 1) : S) : S;
    const uint32_t t_row = tmem + ((uint32_t)(q4 * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    // The softmax warps are the instruction-bound part of the kernel (2 warps per scheduler, 288+ score
    // elements per thread and pass), so the inner loops are kept to ~6 instructions per element: bf16
    // rounding by cvt + shift, one FMA into the exponent, ex2 on the SFU, multiplication by 1/l.
    float m = -INFINITY, l = 0.f;
    for (int i = 0; i < n_tiles; ++i) {  // ---- pass 1: row max and sum of exp over bf16 scores ----
      const int s = i & 1;
      f_wait(&bars.s_full[s], (i >> 1) & 1);
      f_fence_after();
      const int jbase = i * FA_TK + half * 64;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t r[32];
        f_tmem_ld32(t_row + (uint32_t)(s * 128 + half * 64 + c0), r);
        float sc[32];
        float cm = -INFINITY;
        if (jbase + c0 + 32 <= vis) {  // whole chunk visible: no per-element mask
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            sc[e] = rbf(__uint_as_float(r[e]));
            cm = fmaxf(cm, sc[e]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            sc[e] = (jbase + c0 + e < vis) ? rbf(__uint_as_float(r[e])) : -INFINITY;
            cm = fmaxf(cm, sc[e]);
          }
        }
        if (cm > -INFINITY) {
          const float mn = fmaxf(m, cm);
          const float nb = -mn * LOG2E;
          float add = 0.f;
#pragma unroll
          for (int e = 0; e < 32; ++e) add += f_ex2(fmaf(sc[e], LOG2E, nb));
          l = l * f_ex2((m - mn) * LOG2E) + add;
          m = mn;
        }
      }
      f_fence_before();
      f_arrive(&bars.s_empty[s]);
    }
    stat[half][row] = make_float2(m, l);
    f_sbar();
    {
      const float2 a = stat[0][row], b = stat[1][row];
      m = fmaxf(a.x, b.x);
      l = (a.y > 0.f ? a.y * f_ex2((a.x - m) * LOG2E) : 0.f) + (b.y > 0.f ? b.y * f_ex2((b.x - m) * LOG2E) : 0.f);
    }
    const float inv_l = 1.0f / l;
    const float nbm = -m * LOG2E;
    for (int j = 0; j < n_tiles; ++j) {  // ---- pass 2: p = bf16(exp(s - m) / l) -> shared memory ----
      const int i = n_tiles + j, s = i & 1;
      f_wait(&bars.s_full[s], (i >> 1) & 1);
      f_fence_after();
      f_wait(&bars.p_empty, (j & 1) ^ 1);  // the previous P.V has consumed the P tile
      uint8_t* prow = Ps + half * FA_BLK + row * 128;
      const int jbase = j * FA_TK + half * 64;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t r[32];
        f_tmem_ld32(t_row + (uint32_t)(s * 128 + half * 64 + c0), r);
        const bool full = jbase + c0 + 32 <= vis;
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          float pv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float pe = f_ex2(fmaf(rbf(__uint_as_float(r[e + u])), LOG2E, nbm)) * inv_l;
            pv[u] = (full || jbase + c0 + e + u < vis) ? pe : 0.f;
          }
          uint4 o;
          o.x = pack2(pv[0], pv[1]); o.y = pack2(pv[2], pv[3]); o.z = pack2(pv[4], pv[5]); o.w = pack2(pv[6], pv[7]);
          const int chunk = (c0 + e) >> 3;
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = o;
        }
      }
      f_fence_before();
      f_arrive(&bars.s_empty[s]);
      f_fence_async();  // generic-proxy writes of P -> visible to the tensor core (async proxy)
      f_arrive(&bars.p_full);
    }
    // ---- epilogue: O row -> bf16 -> global (16-column chunks alternate between the row's two threads) ----
    f_wait(&bars.o_full, 0);
    f_fence_after();
    bf16* orow = p.out + (long)(p.q0 + qi) * p.o_ts + (long)h * p.hd;
    for (int c0 = 16 * half; c0 < p.hd; c0 += 32) {
      uint32_t r[16];
      f_tmem_ld16(t_row + 256u + (uint32_t)c0, r);
      if (qi < p.Lq) {
        if (c0 + 16 <= p.hd) {
          uint4 o0, o1;
          o0.x = pack2(__uint_as_float(r[0]), __uint_as_float(r[1]));
          o0.y = pack2(__uint_as_float(r[2]), __uint_as_float(r[3]));
          o0.z = pack2(__uint_as_float(r[4]), __uint_as_float(r[5]));
          o0.w = pack2(__uint_as_float(r[6]), __uint_as_float(r[7]));
          o1.x = pack2(__uint_as_float(r[8]), __uint_as_float(r[9]));
          o1.y = pack2(__uint_as_float(r[10]), __uint_as_float(r[11]));
          o1.z = pack2(__uint_as_float(r[12]), __uint_as_float(r[13]));
          o1.w = pack2(__uint_as_float(r[14]), __uint_as_float(r[15]));
          *reinterpret_cast<uint4*>(orow + c0) = o0;
          *reinterpret_cast<uint4*>(orow + c0 + 8) = o1;
        } else {
          for (int e = 0; e < 16 && c0 + e < p.hd; ++e) orow[c0 + e] = f2bf(__uint_as_float(r[e]));
        }
      }
    }
  }
  f_fence_before();
  __syncthreads();
  if (warp == 1) {
    f_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---- host: 3-D tensor maps ----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn fa_get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}
struct FaKey {
  const void* ptr;
  long s1, s2;
  int d0, d1, d2, b0, b1, b2;
  bool operator==(const FaKey& o) const {
    return ptr == o.ptr && s1 == o.s1 && s2 == o.s2 && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && b0 == o.b0 &&
           b1 == o.b1 && b2 == o.b2;
  }
};
struct FaHash {
  size_t operator()(const FaKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    for (long v : {k.s1, k.s2, (long)k.d0, (long)k.d1, (long)k.d2, (long)k.b0, (long)k.b1, (long)k.b2})
      h = h * 1000003u ^ std::hash<long>()(v);
    return h;
  }
};
// bf16 tensor: dims (d0 contiguous, d1 with stride s1 elements, d2 with stride s2), box (b0, b1, b2)
int fa_tmap3(const void* ptr, int d0, int d1, long s1, int d2, long s2, int b0, int b1, int b2, CUtensorMap* out) {
  static std::unordered_map<FaKey, CUtensorMap, FaHash> cache;
  static std::mutex mu;
  FaKey key{ptr, s1, s2, d0, d1, d2, b0, b1, b2};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return B200_OK;
    }
  }
  EncodeTiledFn enc = fa_get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return B200_ERR_CUDA;
  }
  cuuint64_t gdim[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t gstr[2] = {(cuuint64_t)s1 * 2, (cuuint64_t)s2 * 2};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap tm;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed (%d) ptr=%p dims=(%d,%d,%d) strides=(%ld,%ld) box=(%d,%d,%d)", (int)r,
              ptr, d0, d1, d2, s1, s2, b0, b1, b2);
    return B200_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 16384) cache.clear();
    cache[key] = tm;
  }
  *out = tm;
  return B200_OK;
}

}  // namespace

bool attention_fa_supported(const void* q, long q_ts, long q_hs, const void* k, long k_ts, long k_hs,
                            const void* vt, long vt_hs, long vt_ds, const void* out, long o_ts, int hd) {
  auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  return hd % 8 == 0 && hd >= 16 && hd <= 128 && al(q) && al(k) && al(vt) && al(out) && (q_ts % 8) == 0 &&
         (q_hs % 8) == 0 && (k_ts % 8) == 0 && (k_hs % 8) == 0 && (vt_hs % 8) == 0 && (vt_ds % 8) == 0 &&
         (o_ts % 8) == 0;
}

// q: PRE-SCALED queries (bf16(q * bf16(scale))), element (t, h, d) at q + t*q_ts + h*q_hs + d
// k: keys, element (s, kvh, d) at k + s*k_ts + kvh*k_hs + d
// vt: values transposed, element (kvh, d, s) at vt + kvh*vt_hs + d*vt_ds + s   (keys contiguous)
// The tensors hold q_tot query tokens / k_tot keys; this call attends queries [q0, q0+Lq) to keys
// [k0, k0+S) (one vision segment, or the whole prompt); out row t is written at out + t*o_ts.
int attention_fa(const void* q, long q_ts, long q_hs, const void* k, long k_ts, long k_hs, const void* vt,
                 long vt_hs, long vt_ds, void* out, long o_ts, int n_heads, int n_kv, int hd, int Lq, int S,
                 int causal, cudaStream_t st, int q0, int q_tot, int k0, int k_tot, const void* segs, int n_seg) {
  if (q_tot <= 0) q_tot = q0 + Lq;
  if (k_tot <= 0) k_tot = k0 + S;
  B200_REQUIRE(!segs || n_seg > 0, "attention_fa: segment table without segments");
  B200_REQUIRE(attention_fa_supported(q, q_ts, q_hs, k, k_ts, k_hs, vt, vt_hs, vt_ds, out, o_ts, hd),
               "attention_fa: unsupported layout (hd=%d)", hd);
  B200_REQUIRE(Lq > 0 && S > 0 && n_heads % n_kv == 0, "attention_fa: bad shape");
  const int hdp = (hd + 15) & ~15;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = fa_tmap3(q, hd, n_heads, q_hs, q_tot, q_ts, 64, 1, FA_TQ, &tq))) return rc;
  if ((rc = fa_tmap3(k, hd, n_kv, k_hs, k_tot, k_ts, 64, 1, FA_TK, &tk))) return rc;
  if ((rc = fa_tmap3(vt, k_tot, hd, vt_ds, n_kv, vt_hs, 64, hdp, 1, &tv))) return rc;
  static unsigned long long set_mask = 0ull;
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  const size_t smem = 12 * FA_BLK + 1024;
  if (!(set_mask >> (dev & 63) & 1ull)) {
    B200_CUDA(cudaFuncSetAttribute(attention_fa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set_mask |= 1ull << (dev & 63);
  }
  FaParams p;
  p.out = (bf16*)out; p.o_ts = o_ts; p.n_heads = n_heads; p.n_kv = n_kv; p.hd = hd; p.hdp = hdp;
  p.Lq = Lq; p.S = S; p.causal = causal; p.q0 = q0; p.k0 = k0;
  p.segs = (const int4*)segs;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(cdiv(Lq, FA_TQ), n_heads, segs ? n_seg : 1);
  lc.blockDim = dim3(FA_THREADS);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  B200_CUDA(cudaLaunchKernelEx(&lc, attention_fa_kernel, tq, tk, tv, p));
  return B200_OK;
}

}  // namespace b200

extern "C" int b200_attention_fa(const void* q, long q_ts, long q_hs, const void* k, long k_ts, long k_hs,
                                 const void* vt, long vt_hs, long vt_ds, void* out, long o_ts, int n_heads,
                                 int n_kv, int hd, int Lq, int S, int causal, void* stream) {
  return b200::attention_fa(q, q_ts, q_hs, k, k_ts, k_hs, vt, vt_hs, vt_ds, out, o_ts, n_heads, n_kv, hd, Lq, S,
                            causal, (cudaStream_t)stream, 0, Lq, 0, S, nullptr, 0);
}
