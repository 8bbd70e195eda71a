// The autoregressive decode step (reference generate/ar.py:496-515 -> _step
// :334-389 -> models/qwen2_vl/language.py:404-518 with L == 1), batch 1.
// Memory-bound: every weight byte is streamed exactly once per token with
// 16-byte no-allocate loads, fp32 accumulation, warp-shuffle reductions.
//
// Per layer:  k_qkv   RMSNorm + [Wq;Wk;Wv] GEMV + bias + M-RoPE + KV append
//             k_attn  GQA attention over the cache (cluster of CTAs per kv head,
//                     softmax statistics / partial outputs exchanged via DSMEM)
//             k_res<1> o_proj GEMV + residual
//             k_gateup RMSNorm + gate/up GEMV + SwiGLU
//             k_res<4> down GEMV + residual
// Tail:       k_head  final RMSNorm + tied-embedding GEMV + logsumexp partials
//             k_sample logprobs (bf16), greedy argmax (lowest index on ties),
//                     token log, state advance, next-token embedding.
// Rounding points follow oracle/qwen2vl.py::lm_layers_forward.
#include <cooperative_groups.h>

#include "common.cuh"
#include "decode.cuh"

namespace cg = cooperative_groups;

namespace b200 {

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
// RMS-normalise `h` (dim) into shared memory as packed bf16 (the oracle's
// rms_norm: two roundings).  All threads of the CTA participate.
__device__ __forceinline__ void cta_rmsnorm_to_smem(const bf16* __restrict__ h,
                                                    const bf16* __restrict__ w, int dim, float eps,
                                                    uint4* xs, float* red) {
  const int nvec = dim >> 3;
  float s = 0.f;
  for (int c = threadIdx.x; c < nvec; c += blockDim.x) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(h + c * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j] * f[j];
  }
  s = warp_sum(s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];  // same order in every CTA
  const float rs = 1.0f / sqrtf(tot / (float)dim + eps);
  for (int c = threadIdx.x; c < nvec; c += blockDim.x) {
    float f[8], wf[8];
    unpack8(*reinterpret_cast<const uint4*>(h + c * 8), f);
    unpack8(*reinterpret_cast<const uint4*>(w + c * 8), wf);
    uint4 o;
    o.x = pack2(rbf(rbf(f[0] * rs) * wf[0]), rbf(rbf(f[1] * rs) * wf[1]));
    o.y = pack2(rbf(rbf(f[2] * rs) * wf[2]), rbf(rbf(f[3] * rs) * wf[3]));
    o.z = pack2(rbf(rbf(f[4] * rs) * wf[4]), rbf(rbf(f[5] * rs) * wf[5]));
    o.w = pack2(rbf(rbf(f[6] * rs) * wf[6]), rbf(rbf(f[7] * rs) * wf[7]));
    xs[c] = o;
  }
  __syncthreads();
}

__device__ __forceinline__ void cta_copy_to_smem(const bf16* __restrict__ x, int dim, uint4* xs) {
  const int nvec = dim >> 3;
  for (int c = threadIdx.x; c < nvec; c += blockDim.x)
    xs[c] = *reinterpret_cast<const uint4*>(x + c * 8);
  __syncthreads();
}

// One warp: NR dot products of weight rows (global, streamed) with the shared
// activation vector over chunk range [c_begin, c_end) (16-byte chunks).  Up to
// 4*NR independent 16-byte loads in flight per lane.
template <int NR>
__device__ __forceinline__ void warp_dot(const bf16* const (&wrow)[NR], const uint4* xs,
                                         int c_begin, int c_end, int lane, float (&acc)[NR]) {
#pragma unroll
  for (int r = 0; r < NR; ++r) acc[r] = 0.f;
  for (int c0 = c_begin + lane; c0 < c_end; c0 += 128) {
    uint4 w[NR][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + 32 * u;
#pragma unroll
      for (int r = 0; r < NR; ++r)
        w[r][u] = (c < c_end) ? ldg_stream(wrow[r] + (long)c * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + 32 * u;
      if (c < c_end) {
        float xf[8];
        unpack8(xs[c], xf);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          float wf[8];
          unpack8(w[r][u], wf);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[r] = fmaf(wf[j], xf[j], acc[r]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) acc[r] = warp_sum(acc[r]);
}

// ---------------------------------------------------------------------------
// k_qkv: one warp per rotary pair (rows j and j+hd/2 of one head slot)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_qkv(const DecodeDims d, const LayerW lw,
                                             const bf16* __restrict__ h, bf16* __restrict__ qbuf,
                                             bf16* __restrict__ kc, bf16* __restrict__ vc,
                                             const DecState* __restrict__ st,
                                             const float* __restrict__ inv_freq) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint4* xs = reinterpret_cast<uint4*>(sm);
  __shared__ float red[8];
  cta_rmsnorm_to_smem(h, lw.ln1, d.hidden, d.eps, xs, red);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = d.hd >> 1;
  const int slots = d.n_heads + 2 * d.n_kv;
  const int ntask = slots * half;
  const int nvec = d.hidden >> 3;
  const int ctx = st->ctx, pos = st->pos;
  for (int task = blockIdx.x * 8 + warp; task < ntask; task += gridDim.x * 8) {
    const int slot = task / half, j = task % half;
    const int r1 = slot * d.hd + j, r2 = r1 + half;
    const bf16* rows[2] = {lw.wqkv + (long)r1 * d.hidden, lw.wqkv + (long)r2 * d.hidden};
    float acc[2];
    warp_dot<2>(rows, xs, 0, nvec, lane, acc);
    if (lane == 0) {
      const float y1 = rbf(acc[0] + bf2f(lw.bqkv[r1]));
      const float y2 = rbf(acc[1] + bf2f(lw.bqkv[r2]));
      if (slot >= d.n_heads + d.n_kv) {
        bf16* dst = vc + ((long)(slot - d.n_heads - d.n_kv) * d.cap + ctx) * d.hd;
        dst[j] = f2bf(y1);
        dst[j + half] = f2bf(y2);
      } else {
        // M-RoPE with identical t/h/w position on decode (language.py:476-509)
        const float ang = (float)pos * inv_freq[j];
        const float c = rbf(cosf(ang)), s = rbf(sinf(ang));
        const float o1 = rbf(rbf(y1 * c) + rbf((-y2) * s));
        const float o2 = rbf(rbf(y2 * c) + rbf(y1 * s));
        bf16* dst = (slot < d.n_heads)
                        ? qbuf + (long)slot * d.hd
                        : kc + ((long)(slot - d.n_heads) * d.cap + ctx) * d.hd;
        dst[j] = f2bf(o1);
        dst[j + half] = f2bf(o2);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// k_attn: cluster of CL CTAs per kv head; rank r owns keys [r*chunk, (r+1)*chunk)
// ---------------------------------------------------------------------------
constexpr int ATT_MAXG = 8;  // q heads per kv head
__global__ void __launch_bounds__(256) k_attn(const DecodeDims d, const bf16* __restrict__ qbuf,
                                              const bf16* __restrict__ kc,
                                              const bf16* __restrict__ vc, bf16* __restrict__ out,
                                              const DecState* __restrict__ st, int chunk_cap) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  extern __shared__ __align__(16) uint8_t sm[];
  const int G = d.n_heads / d.n_kv;
  const int hd = d.hd;
  const int EPL = hd >> 5;  // elements per lane (hd = 64 -> 2, 128 -> 4)
  float* sc = reinterpret_cast<float*>(sm);        // [G][chunk_cap]
  float* stats = sc + (long)G * chunk_cap;         // [G][2]
  float* part = stats + 2 * ATT_MAXG;              // [G][hd]
  float* red = part + ATT_MAXG * hd;               // [8][G][hd]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kvh = blockIdx.y;
  const int nkeys = st->ctx + 1;
  const int chunk = (nkeys + CL - 1) / CL;
  const int k0 = min(nkeys, rank * chunk), k1 = min(nkeys, k0 + chunk);
  const int nloc = k1 - k0;
  const bf16* kb = kc + (long)kvh * d.cap * hd;
  const bf16* vb = vc + (long)kvh * d.cap * hd;

  // scaled + rounded queries of the G heads, EPL elements per lane
  float qs[ATT_MAXG][4];
#pragma unroll
  for (int g = 0; g < ATT_MAXG; ++g) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      qs[g][e] = 0.f;
      if (g < G && e < EPL)
        qs[g][e] = rbf(bf2f(qbuf[(long)(kvh * G + g) * hd + lane * EPL + e]) * d.scale_bf);
    }
  }
  // ---- scores ----
  for (int j = warp; j < nloc; j += 8) {
    const bf16* kr = kb + (long)(k0 + j) * hd + lane * EPL;
    float kf[4] = {0.f, 0.f, 0.f, 0.f};
    if (EPL == 4) {
      unpack4(*reinterpret_cast<const uint2*>(kr), kf);
    } else {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(kr);
      kf[0] = __uint_as_float(w << 16);
      kf[1] = __uint_as_float(w & 0xffff0000u);
    }
#pragma unroll
    for (int g = 0; g < ATT_MAXG; ++g) {
      if (g < G) {
        float s = qs[g][0] * kf[0] + qs[g][1] * kf[1] + qs[g][2] * kf[2] + qs[g][3] * kf[3];
        s = warp_sum(s);
        if (lane == 0) sc[(long)g * chunk_cap + j] = rbf(s);
      }
    }
  }
  __syncthreads();
  // ---- local max / sum(exp) per head: warp g ----
  if (warp < G) {
    float m = -INFINITY;
    for (int j = lane; j < nloc; j += 32) m = fmaxf(m, sc[(long)warp * chunk_cap + j]);
    m = warp_max(m);
    float l = 0.f;
    for (int j = lane; j < nloc; j += 32) l += expf(sc[(long)warp * chunk_cap + j] - m);
    l = warp_sum(l);
    if (lane == 0) {
      stats[warp * 2] = m;
      stats[warp * 2 + 1] = (nloc > 0) ? l : 0.f;
    }
  }
  cluster.sync();
  // ---- global statistics via DSMEM, p = bf16(exp(s - M) / L) in place ----
  if (warp < G) {
    float M = -INFINITY;
    for (int r = 0; r < CL; ++r) M = fmaxf(M, cluster.map_shared_rank(stats, r)[warp * 2]);
    float Ltot = 0.f;
    for (int r = 0; r < CL; ++r) {
      const float* rs = cluster.map_shared_rank(stats, r);
      const float mr = rs[warp * 2], lr = rs[warp * 2 + 1];
      if (lr > 0.f) Ltot += lr * expf(mr - M);
    }
    for (int j = lane; j < nloc; j += 32) {
      float* p = &sc[(long)warp * chunk_cap + j];
      *p = rbf(expf(*p - M) / Ltot);
    }
  }
  __syncthreads();
  // ---- partial output: warp w takes keys w, w+8, ... ; lane owns EPL dims ----
  float acc[ATT_MAXG][4];
#pragma unroll
  for (int g = 0; g < ATT_MAXG; ++g)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[g][e] = 0.f;
  for (int j = warp; j < nloc; j += 8) {
    const bf16* vr = vb + (long)(k0 + j) * hd + lane * EPL;
    float vf[4] = {0.f, 0.f, 0.f, 0.f};
    if (EPL == 4) {
      unpack4(*reinterpret_cast<const uint2*>(vr), vf);
    } else {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(vr);
      vf[0] = __uint_as_float(w << 16);
      vf[1] = __uint_as_float(w & 0xffff0000u);
    }
#pragma unroll
    for (int g = 0; g < ATT_MAXG; ++g) {
      if (g < G) {
        const float p = sc[(long)g * chunk_cap + j];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[g][e] = fmaf(p, vf[e], acc[g][e]);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < ATT_MAXG; ++g)
    if (g < G)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (e < EPL) red[((long)warp * G + g) * hd + lane * EPL + e] = acc[g][e];
  __syncthreads();
  for (int i = threadIdx.x; i < G * hd; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[(long)w * G * hd + i];
    part[i] = s;
  }
  cluster.sync();
  if (rank == 0) {
    for (int i = threadIdx.x; i < G * hd; i += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < CL; ++r) s += cluster.map_shared_rank(part, r)[i];
      out[(long)kvh * G * hd + i] = f2bf(s);
    }
  }
  cluster.sync();  // keep peers' shared memory alive until rank 0 has read it
}

// ---------------------------------------------------------------------------
// k_res<WPR>: h[r] = bf16(h[r] + bf16(W[r,:] . x)),  WPR warps share one row
// ---------------------------------------------------------------------------
template <int WPR>
__global__ void __launch_bounds__(256) k_res(const bf16* __restrict__ W,
                                             const bf16* __restrict__ x, bf16* __restrict__ h,
                                             int N, int K) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint4* xs = reinterpret_cast<uint4*>(sm);
  __shared__ float red[8];
  cta_copy_to_smem(x, K, xs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int RPC = 8 / WPR;  // rows per CTA per iteration
  const int nvec = K >> 3;
  const int sub = warp % WPR, rloc = warp / WPR;
  const int cb = (int)((long)nvec * sub / WPR), ce = (int)((long)nvec * (sub + 1) / WPR);
  for (int r0 = blockIdx.x * RPC; r0 < N; r0 += gridDim.x * RPC) {
    const int r = r0 + rloc;
    float acc[1] = {0.f};
    if (r < N) {
      const bf16* rows[1] = {W + (long)r * K};
      warp_dot<1>(rows, xs, cb, ce, lane, acc);
    }
    if (WPR == 1) {
      if (lane == 0 && r < N) h[r] = f2bf(rbf(bf2f(h[r]) + rbf(acc[0])));
    } else {
      if (lane == 0) red[warp] = acc[0];
      __syncthreads();
      if (sub == 0 && lane == 0 && r < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < WPR; ++i) s += red[warp + i];
        h[r] = f2bf(rbf(bf2f(h[r]) + rbf(s)));
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------
// k_gateup: act[i] = swiglu(bf16(Wg[i,:].x), bf16(Wu[i,:].x)),  x = rmsnorm(h)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gateup(const DecodeDims d, const LayerW lw,
                                                const bf16* __restrict__ h,
                                                bf16* __restrict__ act) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint4* xs = reinterpret_cast<uint4*>(sm);
  __shared__ float red[8];
  cta_rmsnorm_to_smem(h, lw.ln2, d.hidden, d.eps, xs, red);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = d.hidden >> 3;
  for (int i = blockIdx.x * 8 + warp; i < d.inter; i += gridDim.x * 8) {
    const bf16* rows[2] = {lw.wgu + (long)i * d.hidden, lw.wgu + (long)(d.inter + i) * d.hidden};
    float acc[2];
    warp_dot<2>(rows, xs, 0, nvec, lane, acc);
    if (lane == 0) act[i] = f2bf(swiglu_bf(rbf(acc[0]), rbf(acc[1])));
  }
}

// ---------------------------------------------------------------------------
// k_head: logits[v] = bf16(E[v,:] . rmsnorm(h)); per-CTA (max, sum exp) partials
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_head(const DecodeDims d, const bf16* __restrict__ norm_w,
                                              const bf16* __restrict__ E,
                                              const bf16* __restrict__ h,
                                              bf16* __restrict__ logits,
                                              float2* __restrict__ partials) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint4* xs = reinterpret_cast<uint4*>(sm);
  __shared__ float red[8];
  __shared__ float2 wstat[8];
  cta_rmsnorm_to_smem(h, norm_w, d.hidden, d.eps, xs, red);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = d.hidden >> 3;
  float m = -INFINITY, l = 0.f;  // running logsumexp state (lane 0 meaningful)
  for (int v0 = (blockIdx.x * 8 + warp) * 2; v0 < d.vocab; v0 += gridDim.x * 16) {
    const int v1 = min(v0 + 1, d.vocab - 1);
    const bf16* rows[2] = {E + (long)v0 * d.hidden, E + (long)v1 * d.hidden};
    float acc[2];
    warp_dot<2>(rows, xs, 0, nvec, lane, acc);
    const float a = rbf(acc[0]), b = rbf(acc[1]);
    if (lane == 0) {
      logits[v0] = f2bf(a);
      float mn = fmaxf(m, a);
      l = l * expf(m - mn) + expf(a - mn);
      m = mn;
      if (v0 + 1 < d.vocab) {
        logits[v0 + 1] = f2bf(b);
        mn = fmaxf(m, b);
        l = l * expf(m - mn) + expf(b - mn);
        m = mn;
      }
    }
  }
  if (lane == 0) wstat[warp] = make_float2(m, l);
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = -INFINITY;
    for (int i = 0; i < 8; ++i) M = fmaxf(M, wstat[i].x);
    float L = 0.f;
    for (int i = 0; i < 8; ++i)
      if (wstat[i].y > 0.f) L += wstat[i].y * expf(wstat[i].x - M);
    partials[blockIdx.x] = make_float2(M, L);
  }
}

// ---------------------------------------------------------------------------
// k_sample: logprobs = bf16(logits - bf16(logsumexp)), greedy argmax with the
// lowest index on ties (packed 64-bit atomicMax), last CTA finalises the step.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(256) k_sample(const DecodeDims d,
                                                const bf16* __restrict__ logits,
                                                const float2* __restrict__ partials, int n_part,
                                                bf16* __restrict__ logprobs,
                                                const bf16* __restrict__ E, bf16* __restrict__ h,
                                                DecState* __restrict__ st,
                                                int* __restrict__ token_log, int log_cap,
                                                const int* __restrict__ force_tokens,
                                                int advance) {
  __shared__ float s_lse;
  __shared__ unsigned long long s_best;
  __shared__ int s_last, s_feed;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    float M = -INFINITY;
    for (int i = lane; i < n_part; i += 32) M = fmaxf(M, partials[i].x);
    M = warp_max(M);
    float L = 0.f;
    for (int i = lane; i < n_part; i += 32)
      if (partials[i].y > 0.f) L += partials[i].y * expf(partials[i].x - M);
    L = warp_sum(L);
    if (lane == 0) {
      s_lse = rbf(M + logf(L));
      s_best = 0ull;
    }
  }
  __syncthreads();
  const float lse = s_lse;
  unsigned long long best = 0ull;
  const int nvec = d.vocab >> 3;  // vocab % 8 == 0 is checked on the host
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nvec; c += gridDim.x * blockDim.x) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(logits + (long)c * 8), f);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j] = rbf(f[j] - lse);
      const unsigned long long key =
          ((unsigned long long)orderable(o[j]) << 32) | (0xFFFFFFFFu - (uint32_t)(c * 8 + j));
      best = key > best ? key : best;
    }
    uint4 ov;
    ov.x = pack2(o[0], o[1]);
    ov.y = pack2(o[2], o[3]);
    ov.z = pack2(o[4], o[5]);
    ov.w = pack2(o[6], o[7]);
    *reinterpret_cast<uint4*>(logprobs + (long)c * 8) = ov;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(&s_best, best);
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicMax(&st->best_key, s_best);
    __threadfence();
    const unsigned int done = atomicAdd(&st->blocks_done, 1u);
    s_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  // ---- last CTA: finalise the step ----
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long key = atomicMax(&st->best_key, 0ull);  // read
    const int tok = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
    const int n = st->n_out;
    token_log[n % log_cap] = tok;
    int feed = tok;
    if (st->use_force) feed = force_tokens[n % log_cap];
    s_feed = feed;
    st->tok = feed;
    st->n_out = n + 1;
    st->ctx += advance;  // 1 for a decode step, 0 for the prefill call (state pre-armed)
    st->pos += advance;
    st->best_key = 0ull;
    st->blocks_done = 0u;
  }
  __syncthreads();
  const int feed = s_feed;
  const int nv = d.hidden >> 3;
  for (int c = threadIdx.x; c < nv; c += blockDim.x)
    *reinterpret_cast<uint4*>(h + c * 8) =
        *reinterpret_cast<const uint4*>(E + (long)feed * d.hidden + c * 8);
}

// small kernel: arm / overwrite the decode state (after prefill, or explicitly)
__global__ void k_set_state(DecState* st, int tok, int ctx, int pos, int use_force, int set_tok,
                            const bf16* __restrict__ E, bf16* __restrict__ h, int hidden) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (set_tok) st->tok = tok;
    st->ctx = ctx;
    st->pos = pos;
    st->use_force = use_force;
  }
  if (set_tok) {
    const int nv = hidden >> 3;
    for (int c = threadIdx.x; c < nv; c += blockDim.x)
      *reinterpret_cast<uint4*>(h + c * 8) =
          *reinterpret_cast<const uint4*>(E + (long)tok * hidden + c * 8);
  }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
static int g_sm_count = 148;
void decode_set_sm_count(int n) { g_sm_count = n > 0 ? n : 148; }

int launch_qkv(const DecodeDims& d, const LayerW& lw, const bf16* h, bf16* qbuf, bf16* kc,
               bf16* vc, const DecState* st, const float* inv_freq, cudaStream_t s) {
  const int ntask = (d.n_heads + 2 * d.n_kv) * (d.hd / 2);
  const int grid = min(cdiv(ntask, 8), g_sm_count * 2);
  k_qkv<<<grid, 256, (size_t)d.hidden * 2, s>>>(d, lw, h, qbuf, kc, vc, st, inv_freq);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

size_t attn_smem_bytes(const DecodeDims& d, int chunk_cap) {
  const int G = d.n_heads / d.n_kv;
  return ((size_t)G * chunk_cap + 2 * ATT_MAXG + (size_t)ATT_MAXG * d.hd + (size_t)8 * G * d.hd) * 4;
}

int launch_attn(const DecodeDims& d, const bf16* qbuf, const bf16* kc, const bf16* vc, bf16* out,
                const DecState* st, int cluster, cudaStream_t s) {
  const int G = d.n_heads / d.n_kv;
  B200_REQUIRE(G <= ATT_MAXG && (d.hd == 64 || d.hd == 128),
               "decode attention: G=%d (max %d) head_dim=%d (64|128)", G, ATT_MAXG, d.hd);
  const int chunk_cap = cdiv(d.cap, cluster);
  const size_t smem = attn_smem_bytes(d, chunk_cap);
  B200_REQUIRE(smem <= 200 * 1024, "decode attention: cache capacity %d too large", d.cap);
  static size_t set_smem = 0;
  if (smem > set_smem) {
    B200_CUDA(cudaFuncSetAttribute(k_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set_smem = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster, d.n_kv, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200_CUDA(cudaLaunchKernelEx(&cfg, k_attn, d, qbuf, kc, vc, out, st, chunk_cap));
  return B200_OK;
}

int launch_res(const bf16* W, const bf16* x, bf16* h, int N, int K, cudaStream_t s) {
  const size_t smem = (size_t)K * 2;
  if (K <= 2048) {
    k_res<1><<<min(cdiv(N, 8), g_sm_count * 2), 256, smem, s>>>(W, x, h, N, K);
  } else {
    k_res<4><<<min(cdiv(N, 2), g_sm_count * 6), 256, smem, s>>>(W, x, h, N, K);
  }
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int launch_gateup(const DecodeDims& d, const LayerW& lw, const bf16* h, bf16* act,
                  cudaStream_t s) {
  k_gateup<<<min(cdiv(d.inter, 8), g_sm_count * 8), 256, (size_t)d.hidden * 2, s>>>(d, lw, h, act);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int head_grid() { return g_sm_count * 8; }

int launch_head(const DecodeDims& d, const bf16* norm_w, const bf16* E, const bf16* h,
                bf16* logits, float2* partials, cudaStream_t s) {
  k_head<<<head_grid(), 256, (size_t)d.hidden * 2, s>>>(d, norm_w, E, h, logits, partials);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int launch_sample(const DecodeDims& d, const bf16* logits, const float2* partials, bf16* logprobs,
                  const bf16* E, bf16* h, DecState* st, int* token_log, int log_cap,
                  const int* force_tokens, int advance, cudaStream_t s) {
  k_sample<<<g_sm_count, 256, 0, s>>>(d, logits, partials, head_grid(), logprobs, E, h, st,
                                      token_log, log_cap, force_tokens, advance);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int launch_set_state(DecState* st, int tok, int ctx, int pos, int use_force, int set_tok,
                     const bf16* E, bf16* h, int hidden, cudaStream_t s) {
  k_set_state<<<1, 256, 0, s>>>(st, tok, ctx, pos, use_force, set_tok, E, h, hidden);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // namespace b200
