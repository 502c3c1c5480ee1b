// Small fused kernels around the decoder GEMMs.  Every rounding point mirrors the HF bf16 Llama
// path (HF:models/llama/modeling_llama.py) so that the only numeric difference left against the
// reference forward is the accumulation order inside the GEMMs and the attention.
#pragma once
#include "ptx.cuh"
#include "streamk.h"

namespace vlo {

__device__ __forceinline__ float bf16_load(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// block-wide sum for <= 1024 threads
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// ---------------------------------------------------------------------------------------------
// Split-K fix-up + residual add + RMSNorm in one pass over a token row:
//   y      = bf16(sum_s part[s][t][:])                (the Linear output, o_proj / down_proj)
//   h[t]   = bf16(h[t] + y)                           (residual, HF:...llama.py:325,331)
//   xn[t]  = w * bf16(h * rsqrt(mean(h^2) + eps))     (LlamaRMSNorm.forward, HF:...llama.py:62-67)
// n_splits == 0: no partials, plain RMSNorm of h (first layer).  `xn_last` (optional) receives
// compact copies of the rows flagged in last_index (>= 0): the final-norm rows the lm_head needs.
// grid = T rows, block = 256.
struct ResidNormParams {
  const float* part;
  int n_splits;             // 0: no partials; > 0: that many planes; < 0: stream-K planes per 128-column tile (sk)
  long long split_stride;
  SkInfo sk;
  __nv_bfloat16* h;         // [T, H] residual stream, updated in place
  const __nv_bfloat16* w;   // [H] norm weight
  __nv_bfloat16* xn;        // [T, H] normalised output (may be null)
  __nv_bfloat16* xn_last;   // [n_last, H] or null
  const int* last_index;    // [T] or null
  int H;
  float eps;
};

constexpr int kFixMaxPlanes = 8;  // stream-K never needs more (engine checks)

__global__ void __launch_bounds__(1024) resid_rmsnorm_kernel(const ResidNormParams p) {
  // one CTA per token row, 4 contiguous elements per thread per sweep; the plane loads of a sweep are
  // independent (predicated, fully unrolled) so they are all in flight together.
  extern __shared__ float row[];  // H floats
  __shared__ float red[32];
  const int t = blockIdx.x;
  __nv_bfloat16* h = p.h + static_cast<size_t>(t) * p.H;
  pdl_trigger();
  // Everything that does not depend on the producer kernel is fetched BEFORE the grid dependency resolves: the norm
  // weight (a static tensor that the 15 GB weight stream has long evicted from L2 - an HBM miss that used to sit on
  // the critical path after the block reduction), the stream-K plane count of this thread's first column group and
  // the last-token index (tables uploaded by the memcpy at the start of the step).
  const int i0 = threadIdx.x * 4;
  const uint2 w_pre = (i0 < p.H) ? *reinterpret_cast<const uint2*>(p.w + i0) : make_uint2(0u, 0u);
  const int ns_pre = (p.n_splits < 0 && i0 < p.H) ? sk_planes(i0 >> 7, p.sk) : p.n_splits;
  const int li = p.last_index ? p.last_index[t] : -1;
  pdl_wait();
  float ss = 0.f;
  for (int i = threadIdx.x * 4; i < p.H; i += blockDim.x * 4) {
    const uint2 hraw = *reinterpret_cast<const uint2*>(h + i);
    float2 v01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hraw.x));
    float2 v23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hraw.y));
    if (p.n_splits != 0) {
      const int ns = (i == i0) ? ns_pre : (p.n_splits > 0 ? p.n_splits : sk_planes(i >> 7, p.sk));
      const float* pp = p.part + static_cast<size_t>(t) * p.H + i;
      float4 a[kFixMaxPlanes];
#pragma unroll
      for (int s = 0; s < kFixMaxPlanes; ++s)
        a[s] = (s < ns) ? *reinterpret_cast<const float4*>(pp + s * p.split_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
      float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
#pragma unroll
      for (int s = 0; s < kFixMaxPlanes; ++s) {  // plane order: deterministic
        y0 += a[s].x; y1 += a[s].y; y2 += a[s].z; y3 += a[s].w;
      }
      v01.x = bf16_round(v01.x + bf16_round(y0));
      v01.y = bf16_round(v01.y + bf16_round(y1));
      v23.x = bf16_round(v23.x + bf16_round(y2));
      v23.y = bf16_round(v23.y + bf16_round(y3));
      uint2 o;
      *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v01.x, v01.y);
      *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v23.x, v23.y);
      *reinterpret_cast<uint2*>(h + i) = o;
    }
    row[i] = v01.x; row[i + 1] = v01.y; row[i + 2] = v23.x; row[i + 3] = v23.y;
    ss += v01.x * v01.x + v01.y * v01.y + v23.x * v23.x + v23.y * v23.y;
  }
  const float tot = block_sum(ss, red);
  const float rstd = rsqrtf(tot / static_cast<float>(p.H) + p.eps);
  for (int i = threadIdx.x * 4; i < p.H; i += blockDim.x * 4) {
    const uint2 wraw = (i == i0) ? w_pre : *reinterpret_cast<const uint2*>(p.w + i);
    const float2 w01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wraw.x));
    const float2 w23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wraw.y));
    uint2 o;
    *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(bf16_round(w01.x * bf16_round(row[i] * rstd)),
                                                                     bf16_round(w01.y * bf16_round(row[i + 1] * rstd)));
    *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(bf16_round(w23.x * bf16_round(row[i + 2] * rstd)),
                                                                     bf16_round(w23.y * bf16_round(row[i + 3] * rstd)));
    if (p.xn) *reinterpret_cast<uint2*>(p.xn + static_cast<size_t>(t) * p.H + i) = o;
    if (li >= 0) *reinterpret_cast<uint2*>(p.xn_last + static_cast<size_t>(li) * p.H + i) = o;
  }
}

// ---------------------------------------------------------------------------------------------
// QKV split-K fix-up + RoPE + KV-cache append (SURVEY K12/K13/K14 fused):
//   q,k,v = bf16(sum_s part)                              (HF:...llama.py:262-264)
//   q,k   = x*cos + rotate_half(x)*sin in bf16            (apply_rotary_pos_emb, :146-168)
//   K/V rows written in place at the token's position     (replaces DynamicLayer.update's
//                                                          torch.cat, HF:cache_utils.py:119-120)
// cos/sin: bf16 tables [max_pos, head_dim/2] (the two halves of HF's cos/sin are identical).
// grid = (T, n_heads + 2*n_kv_heads), block = 64 (thread d handles dims d and d+64).
struct QkvRopeParams {
  const float* part;       // [S][T][(nh + 2 nkv) * 128]
  int n_splits;            // > 0 planes, < 0 stream-K (head hh == weight tile hh)
  long long split_stride;
  SkInfo sk;
  const __nv_bfloat16* cos_tab;
  const __nv_bfloat16* sin_tab;
  const int* tok_pos;      // [T] absolute position of the token in its stream
  const long long* tok_kvrow;  // [T] cache row of (key 0, kv head 0) of the token's stream
  int kv_head_stride;      // rows between kv heads
  __nv_bfloat16* q_out;    // [T, nh, 128]
  __nv_bfloat16* k_cache;  // layer base, rows of 128
  __nv_bfloat16* v_cache;
  int n_heads, n_kv_heads;
};

__global__ void __launch_bounds__(64) qkv_rope_append_kernel(const QkvRopeParams p) {
  const int t = blockIdx.x, hh = blockIdx.y, d = threadIdx.x;
  pdl_trigger();
  // Before the grid dependency resolves: token position / cache row (tables uploaded by the memcpy at the start of
  // the step), the RoPE table entries (static) and the plane count - three dependent round trips off the critical path.
  const int width = (p.n_heads + 2 * p.n_kv_heads) * 128;
  const float* pp = p.part + static_cast<size_t>(t) * width + hh * 128 + d;
  const int ns = p.n_splits > 0 ? p.n_splits : sk_planes(hh, p.sk);
  const int pos = p.tok_pos[t];
  const long long kvrow = p.tok_kvrow[t];
  const bool rot = hh < p.n_heads + p.n_kv_heads;
  const float c = rot ? __bfloat162float(p.cos_tab[static_cast<size_t>(pos) * 64 + d]) : 1.f;
  const float s = rot ? __bfloat162float(p.sin_tab[static_cast<size_t>(pos) * 64 + d]) : 0.f;
  pdl_wait();
  float x1 = 0.f, x2 = 0.f;
  if (ns <= kFixMaxPlanes) {  // all plane loads in flight together
    float a1[kFixMaxPlanes], a2[kFixMaxPlanes];
#pragma unroll
    for (int k = 0; k < kFixMaxPlanes; ++k) {
      a1[k] = (k < ns) ? pp[k * p.split_stride] : 0.f;
      a2[k] = (k < ns) ? pp[k * p.split_stride + 64] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kFixMaxPlanes; ++k) {  // plane order: deterministic
      x1 += a1[k];
      x2 += a2[k];
    }
  } else {
    for (int k = 0; k < ns; ++k) {
      x1 += pp[k * p.split_stride];
      x2 += pp[k * p.split_stride + 64];
    }
  }
  x1 = bf16_round(x1);
  x2 = bf16_round(x2);
  if (rot) {
    const float o1 = bf16_round(bf16_round(x1 * c) + bf16_round(-x2 * s));
    const float o2 = bf16_round(bf16_round(x2 * c) + bf16_round(x1 * s));
    x1 = o1;
    x2 = o2;
  }
  __nv_bfloat16* dst;
  if (hh < p.n_heads) {
    dst = p.q_out + (static_cast<size_t>(t) * p.n_heads + hh) * 128;
  } else {
    const int kvh = (hh - p.n_heads) % p.n_kv_heads;
    __nv_bfloat16* base = (hh < p.n_heads + p.n_kv_heads) ? p.k_cache : p.v_cache;
    dst = base + (kvrow + static_cast<long long>(kvh) * p.kv_head_stride + pos) * 128;
  }
  dst[d] = __float2bfloat16_rn(x1);
  dst[d + 64] = __float2bfloat16_rn(x2);
}

// ---------------------------------------------------------------------------------------------
// gate/up split-K fix-up + SwiGLU:  act = bf16(bf16(silu(g)) * u)      (LlamaMLP, :182-184)
// The fused gate|up weight is stored tile-interleaved: 128-row tile j = gate rows of features 64j..64j+63 followed by
// the matching up rows (so one tile holds both operands of 64 activations: gemm_wsf.cuh fuses this kernel into the
// GEMM's finisher; this stand-alone version is the VLO_FUSE=0 fallback).  part column of feature i: gate
// (i/64)*128 + i%64, up = gate + 64.   grid-stride over T*I/4 quads.
struct SwigluParams {
  const float* part;
  int n_splits;  // > 0 planes, < 0 stream-K
  long long split_stride;
  SkInfo sk;
  __nv_bfloat16* act;  // [T, I]
  int T, I;
};
__global__ void __launch_bounds__(256) swiglu_kernel(const SwigluParams p) {
  pdl_trigger();
  const long long n4 = static_cast<long long>(p.T) * p.I / 4;
  const long long idx0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  // index arithmetic and stream-K plane counts of the first element group: before the grid dependency resolves
  int t0 = 0, c0 = 0, ng0 = 0;
  if (idx0 < n4) {
    const long long e = idx0 * 4;
    t0 = static_cast<int>(e / p.I), c0 = static_cast<int>(e % p.I);
    ng0 = p.n_splits > 0 ? p.n_splits : sk_planes(c0 >> 6, p.sk);
  }
  pdl_wait();
  for (long long idx = idx0; idx < n4; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = idx * 4;
    const bool first = idx == idx0;
    const int t = first ? t0 : static_cast<int>(e / p.I), i = first ? c0 : static_cast<int>(e % p.I);
    const float* pg = p.part + static_cast<size_t>(t) * 2 * p.I + (i >> 6) * 128 + (i & 63);
    const int ng = first ? ng0 : (p.n_splits > 0 ? p.n_splits : sk_planes(i >> 6, p.sk));
    const int nu = ng;
    float4 ga[kFixMaxPlanes], ua[kFixMaxPlanes];
#pragma unroll
    for (int s = 0; s < kFixMaxPlanes; ++s) {
      ga[s] = (s < ng) ? *reinterpret_cast<const float4*>(pg + s * p.split_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
      ua[s] = (s < nu) ? *reinterpret_cast<const float4*>(pg + s * p.split_stride + 64) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float g[4] = {0.f, 0.f, 0.f, 0.f}, u[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < kFixMaxPlanes; ++s) {
      g[0] += ga[s].x; g[1] += ga[s].y; g[2] += ga[s].z; g[3] += ga[s].w;
      u[0] += ua[s].x; u[1] += ua[s].y; u[2] += ua[s].z; u[3] += ua[s].w;
    }
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = bf16_round(g[k]), uk = bf16_round(u[k]);
      const float a = bf16_round(gk / (1.0f + expf(-gk)));
      o[k] = a * uk;
    }
    uint2 w;
    *reinterpret_cast<__nv_bfloat162*>(&w.x) = __floats2bfloat162_rn(o[0], o[1]);
    *reinterpret_cast<__nv_bfloat162*>(&w.y) = __floats2bfloat162_rn(o[2], o[3]);
    *reinterpret_cast<uint2*>(p.act + e) = w;
  }
}

// ---------------------------------------------------------------------------------------------
// Token-embedding gather (K10) into rows of the packed step input.
//   ids[i] >= 0  -> dst row dst_rows[i] = table[min(id, vocab-1)]   (clamp: models/modeling_live.py:38)
//   ids[i] <  0  -> leave the row alone (it already holds a frame embedding)
__global__ void __launch_bounds__(256) embed_rows_kernel(const long long* ids, const int* dst_rows, int n,
                                                         const __nv_bfloat16* table, int vocab, int H,
                                                         __nv_bfloat16* dst) {
  const int i = blockIdx.x;
  if (i >= n) return;
  long long id = ids[i];
  if (id < 0) return;
  if (id > vocab - 1) id = vocab - 1;
  const int r = dst_rows ? dst_rows[i] : i;
  const uint4* src = reinterpret_cast<const uint4*>(table + static_cast<size_t>(id) * H);
  uint4* out = reinterpret_cast<uint4*>(dst + static_cast<size_t>(r) * H);
  for (int j = threadIdx.x; j < H / 8; j += blockDim.x) out[j] = src[j];
}

// copy packed rows (bf16) src[i] -> dst[rows[i]]
__global__ void __launch_bounds__(256) scatter_rows_kernel(const __nv_bfloat16* src, const int* rows, int n, int H,
                                                           __nv_bfloat16* dst) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const uint4* s = reinterpret_cast<const uint4*>(src + static_cast<size_t>(i) * H);
  uint4* o = reinterpret_cast<uint4*>(dst + static_cast<size_t>(rows[i]) * H);
  for (int j = threadIdx.x; j < H / 8; j += blockDim.x) o[j] = s[j];
}

// ---------------------------------------------------------------------------------------------
// Speak/silent decision + greedy argmax on the device (demo/inference.py:76-81,
// models/modeling_live.py:177-179): one block per sequence over its bf16 logits row.
struct DecisionOut {  // == vlo_decision
  int argmax_id;
  int argmax_excl_id;
  float p_interval;
  float max_logit;
  float top2_margin;
  float lse;
  int argmax_prob_id;
  int reserved1;
};

__global__ void __launch_bounds__(1024) decision_kernel(const __nv_bfloat16* logits, int vocab, int interval_id,
                                                        DecisionOut* out) {
  // Two vectorised sweeps over the row (16-byte loads, 8 logits each):
  //   A: first-index max, second max and the online sum of exponentials;
  //   B: argmax of the bf16-rounded probabilities with / without the interval id.
  __shared__ float s_m1[32], s_m2[32], s_se[32], s_pa[32], s_pe[32];
  __shared__ int s_i1[32], s_ia[32], s_ie[32];
  __shared__ float s_gmax, s_gmax2, s_sum;
  __shared__ int s_gidx;
  pdl_trigger();
  pdl_wait();
  const __nv_bfloat16* x = logits + static_cast<size_t>(blockIdx.x) * vocab;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int nvec = vocab / 8;  // rows are 16-byte aligned (vocab % 8 == 0 checked by the engine); tail handled below

  // ---- sweep A
  float m1 = -INFINITY, m2 = -INFINITY, se = 0.f;
  int i1 = 0x7fffffff;
  auto take = [&](float v, int idx) {
    if (v > m1) {
      se = se * __expf(m1 - v) + 1.f;  // m1 = -inf on the first element: se = 0 * 0 + 1
      m2 = m1; m1 = v; i1 = idx;
    } else {
      se += __expf(v - m1);
      if (v > m2) m2 = v;
    }
  };
  for (int v0 = tid; v0 < nvec; v0 += 4 * blockDim.x) {  // 4 independent 16-byte loads in flight per thread
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int vi = v0 + u * blockDim.x;
      raw[u] = vi < nvec ? *reinterpret_cast<const uint4*>(x + static_cast<size_t>(vi) * 8) : make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int vi = v0 + u * blockDim.x;
      if (vi >= nvec) break;
      const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
        take(f.x, vi * 8 + 2 * k);
        take(f.y, vi * 8 + 2 * k + 1);
      }
    }
  }
  for (int idx = nvec * 8 + tid; idx < vocab; idx += blockDim.x) take(__bfloat162float(x[idx]), idx);
  auto merge = [&](float om1, int oi1, float om2, float ose) {
    if (om1 > m1 || (om1 == m1 && oi1 < i1)) {
      se = ose + se * __expf(m1 - om1);
      m2 = fmaxf(m1, om2); m1 = om1; i1 = oi1;
    } else if (oi1 != 0x7fffffff) {
      se += ose * __expf(om1 - m1);
      m2 = fmaxf(m2, om1);
    }
  };
  for (int o = 16; o > 0; o >>= 1) {
    const float om1 = __shfl_xor_sync(0xffffffffu, m1, o), om2 = __shfl_xor_sync(0xffffffffu, m2, o);
    const float ose = __shfl_xor_sync(0xffffffffu, se, o);
    const int oi1 = __shfl_xor_sync(0xffffffffu, i1, o);
    merge(om1, oi1, om2, ose);
  }
  if (lane == 0) { s_m1[warp] = m1; s_i1[warp] = i1; s_m2[warp] = m2; s_se[warp] = se; }
  __syncthreads();
  if (warp == 0) {
    m1 = lane < nw ? s_m1[lane] : -INFINITY;
    i1 = lane < nw ? s_i1[lane] : 0x7fffffff;
    m2 = lane < nw ? s_m2[lane] : -INFINITY;
    se = lane < nw ? s_se[lane] : 0.f;
    for (int o = 16; o > 0; o >>= 1) {
      const float om1 = __shfl_xor_sync(0xffffffffu, m1, o), om2 = __shfl_xor_sync(0xffffffffu, m2, o);
      const float ose = __shfl_xor_sync(0xffffffffu, se, o);
      const int oi1 = __shfl_xor_sync(0xffffffffu, i1, o);
      merge(om1, oi1, om2, ose);
    }
    if (lane == 0) { s_gmax = m1; s_gidx = i1; s_gmax2 = m2; s_sum = se; }
  }
  __syncthreads();
  const float gmax = s_gmax, sum = s_sum;

  // ---- sweep B: what `next_score.argmax` sees (demo/inference.py:76-79); first index wins ties
  float pa = -1.f, pe = -1.f;
  int ia = 0x7fffffff, ie = 0x7fffffff;
  auto prob = [&](float v, int idx) {
    const float pr = bf16_round(__expf(v - gmax) / sum);
    if (pr > pa) { pa = pr; ia = idx; }
    const float pz = (idx == interval_id) ? 0.f : pr;  // zero_() then argmax
    if (pz > pe) { pe = pz; ie = idx; }
  };
  for (int v0 = tid; v0 < nvec; v0 += 4 * blockDim.x) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int vi = v0 + u * blockDim.x;
      raw[u] = vi < nvec ? *reinterpret_cast<const uint4*>(x + static_cast<size_t>(vi) * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int vi = v0 + u * blockDim.x;
      if (vi >= nvec) break;
      const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
        prob(f.x, vi * 8 + 2 * k);
        prob(f.y, vi * 8 + 2 * k + 1);
      }
    }
  }
  for (int idx = nvec * 8 + tid; idx < vocab; idx += blockDim.x) prob(__bfloat162float(x[idx]), idx);
  for (int o = 16; o > 0; o >>= 1) {
    const float opa = __shfl_xor_sync(0xffffffffu, pa, o), ope = __shfl_xor_sync(0xffffffffu, pe, o);
    const int oia = __shfl_xor_sync(0xffffffffu, ia, o), oie = __shfl_xor_sync(0xffffffffu, ie, o);
    if (opa > pa || (opa == pa && oia < ia)) { pa = opa; ia = oia; }
    if (ope > pe || (ope == pe && oie < ie)) { pe = ope; ie = oie; }
  }
  if (lane == 0) { s_pa[warp] = pa; s_ia[warp] = ia; s_pe[warp] = pe; s_ie[warp] = ie; }
  __syncthreads();
  if (warp == 0) {
    pa = lane < nw ? s_pa[lane] : -1.f; ia = lane < nw ? s_ia[lane] : 0x7fffffff;
    pe = lane < nw ? s_pe[lane] : -1.f; ie = lane < nw ? s_ie[lane] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      const float opa = __shfl_xor_sync(0xffffffffu, pa, o), ope = __shfl_xor_sync(0xffffffffu, pe, o);
      const int oia = __shfl_xor_sync(0xffffffffu, ia, o), oie = __shfl_xor_sync(0xffffffffu, ie, o);
      if (opa > pa || (opa == pa && oia < ia)) { pa = opa; ia = oia; }
      if (ope > pe || (ope == pe && oie < ie)) { pe = ope; ie = oie; }
    }
    if (lane == 0) {
      DecisionOut d;
      d.argmax_id = s_gidx;
      d.argmax_excl_id = ie;
      d.p_interval = (interval_id >= 0 && interval_id < vocab)
                         ? bf16_round(__expf(__bfloat162float(x[interval_id]) - gmax) / sum)
                         : 0.f;
      d.max_logit = gmax;
      d.top2_margin = gmax - s_gmax2;
      d.lse = gmax + logf(sum);
      d.argmax_prob_id = ia;
      d.reserved1 = 0;
      out[blockIdx.x] = d;
    }
  }
}

// KV synthetic fill (bench pre-fill): deterministic finite pseudo-random bf16 in about [-1, 1].
__global__ void kv_fill_kernel(__nv_bfloat16* base, long long row0, int n_rows, unsigned long long seed) {
  const long long n = static_cast<long long>(n_rows) * 128;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * static_cast<unsigned long long>(row0 * 128 + i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = static_cast<float>(z >> 40) * (1.0f / 16777216.0f);  // [0,1)
    base[row0 * 128 + i] = __float2bfloat16_rn(2.f * u - 1.f);
  }
}

}  // namespace vlo
