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

__global__ void __launch_bounds__(256) resid_rmsnorm_kernel(const ResidNormParams p) {
  extern __shared__ float row[];  // H floats
  __shared__ float red[32];
  const int t = blockIdx.x;
  __nv_bfloat16* h = p.h + static_cast<size_t>(t) * p.H;
  float ss = 0.f;
  for (int i = threadIdx.x * 2; i < p.H; i += blockDim.x * 2) {
    float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(h + i));
    if (p.n_splits != 0) {
      float y0 = 0.f, y1 = 0.f;
      const float* pp = p.part + static_cast<size_t>(t) * p.H + i;
      const int ns = p.n_splits > 0 ? p.n_splits : sk_planes(i >> 7, p.sk);
      for (int s = 0; s < ns; ++s) {
        const float2 a = *reinterpret_cast<const float2*>(pp + s * p.split_stride);
        y0 += a.x;
        y1 += a.y;
      }
      v.x = bf16_round(v.x + bf16_round(y0));
      v.y = bf16_round(v.y + bf16_round(y1));
      *reinterpret_cast<__nv_bfloat162*>(h + i) = __floats2bfloat162_rn(v.x, v.y);
    }
    row[i] = v.x;
    row[i + 1] = v.y;
    ss += v.x * v.x + v.y * v.y;
  }
  const float tot = block_sum(ss, red);
  const float rstd = rsqrtf(tot / static_cast<float>(p.H) + p.eps);
  const int li = p.last_index ? p.last_index[t] : -1;
  for (int i = threadIdx.x * 2; i < p.H; i += blockDim.x * 2) {
    const float2 w = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p.w + i));
    const float a = bf16_round(w.x * bf16_round(row[i] * rstd));
    const float b = bf16_round(w.y * bf16_round(row[i + 1] * rstd));
    const __nv_bfloat162 o = __floats2bfloat162_rn(a, b);
    if (p.xn) *reinterpret_cast<__nv_bfloat162*>(p.xn + static_cast<size_t>(t) * p.H + i) = o;
    if (li >= 0) *reinterpret_cast<__nv_bfloat162*>(p.xn_last + static_cast<size_t>(li) * p.H + i) = o;
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
  const int width = (p.n_heads + 2 * p.n_kv_heads) * 128;
  const float* pp = p.part + static_cast<size_t>(t) * width + hh * 128 + d;
  float x1 = 0.f, x2 = 0.f;
  const int ns = p.n_splits > 0 ? p.n_splits : sk_planes(hh, p.sk);
  for (int s = 0; s < ns; ++s) {
    x1 += pp[s * p.split_stride];
    x2 += pp[s * p.split_stride + 64];
  }
  x1 = bf16_round(x1);
  x2 = bf16_round(x2);
  const int pos = p.tok_pos[t];
  if (hh < p.n_heads + p.n_kv_heads) {
    const float c = __bfloat162float(p.cos_tab[static_cast<size_t>(pos) * 64 + d]);
    const float s = __bfloat162float(p.sin_tab[static_cast<size_t>(pos) * 64 + d]);
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
    dst = base + (p.tok_kvrow[t] + static_cast<long long>(kvh) * p.kv_head_stride + pos) * 128;
  }
  dst[d] = __float2bfloat16_rn(x1);
  dst[d + 64] = __float2bfloat16_rn(x2);
}

// ---------------------------------------------------------------------------------------------
// gate/up split-K fix-up + SwiGLU:  act = bf16(bf16(silu(g)) * u)      (LlamaMLP, :182-184)
// part columns [0, I) = gate, [I, 2I) = up.   grid-stride over T*I/2 pairs.
struct SwigluParams {
  const float* part;
  int n_splits;  // > 0 planes, < 0 stream-K
  long long split_stride;
  SkInfo sk;
  __nv_bfloat16* act;  // [T, I]
  int T, I;
};
__global__ void __launch_bounds__(256) swiglu_kernel(const SwigluParams p) {
  const long long n2 = static_cast<long long>(p.T) * p.I / 2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < n2;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = idx * 2;
    const int t = static_cast<int>(e / p.I), i = static_cast<int>(e % p.I);
    const float* pg = p.part + static_cast<size_t>(t) * 2 * p.I + i;
    float g0 = 0.f, g1 = 0.f, u0 = 0.f, u1 = 0.f;
    const int ng = p.n_splits > 0 ? p.n_splits : sk_planes(i >> 7, p.sk);
    const int nu = p.n_splits > 0 ? p.n_splits : sk_planes((p.I + i) >> 7, p.sk);
    for (int s = 0; s < ng; ++s) {
      const float2 g = *reinterpret_cast<const float2*>(pg + s * p.split_stride);
      g0 += g.x; g1 += g.y;
    }
    for (int s = 0; s < nu; ++s) {
      const float2 u = *reinterpret_cast<const float2*>(pg + s * p.split_stride + p.I);
      u0 += u.x; u1 += u.y;
    }
    g0 = bf16_round(g0); g1 = bf16_round(g1); u0 = bf16_round(u0); u1 = bf16_round(u1);
    const float a0 = bf16_round(g0 / (1.0f + expf(-g0)));
    const float a1 = bf16_round(g1 / (1.0f + expf(-g1)));
    *reinterpret_cast<__nv_bfloat162*>(p.act + e) = __floats2bfloat162_rn(a0 * u0, a1 * u1);
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
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  __shared__ float s_val2[32];
  __shared__ float s_bcast[2];
  const __nv_bfloat16* x = logits + static_cast<size_t>(blockIdx.x) * vocab;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;

  // pass 1: max (first index), second max
  float m1 = -INFINITY, m2 = -INFINITY;
  int i1 = 0x7fffffff;
  for (int i = tid; i < vocab; i += blockDim.x) {
    const float v = __bfloat162float(x[i]);
    if (v > m1) { m2 = m1; m1 = v; i1 = i; }
    else if (v > m2) m2 = v;  // a tie with m1 at a later index lands here -> margin 0
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float om1 = __shfl_xor_sync(0xffffffffu, m1, o);
    const float om2 = __shfl_xor_sync(0xffffffffu, m2, o);
    const int oi1 = __shfl_xor_sync(0xffffffffu, i1, o);
    if (om1 > m1 || (om1 == m1 && oi1 < i1)) { m2 = fmaxf(m1, om2); m1 = om1; i1 = oi1; }
    else m2 = fmaxf(m2, om1);
  }
  if (lane == 0) { s_val[warp] = m1; s_idx[warp] = i1; s_val2[warp] = m2; }
  __syncthreads();
  if (warp == 0) {
    m1 = lane < nw ? s_val[lane] : -INFINITY;
    i1 = lane < nw ? s_idx[lane] : 0x7fffffff;
    m2 = lane < nw ? s_val2[lane] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) {
      const float om1 = __shfl_xor_sync(0xffffffffu, m1, o);
      const float om2 = __shfl_xor_sync(0xffffffffu, m2, o);
      const int oi1 = __shfl_xor_sync(0xffffffffu, i1, o);
      if (om1 > m1 || (om1 == m1 && oi1 < i1)) { m2 = fmaxf(m1, om2); m1 = om1; i1 = oi1; }
      else m2 = fmaxf(m2, om1);
    }
    if (lane == 0) { s_val[0] = m1; s_idx[0] = i1; s_val2[0] = m2; }
  }
  __syncthreads();
  const float gmax = s_val[0];
  const int gidx = s_idx[0];
  const float gmax2 = s_val2[0];
  __syncthreads();

  // pass 2: sum of exp
  float se = 0.f;
  for (int i = tid; i < vocab; i += blockDim.x) se += expf(__bfloat162float(x[i]) - gmax);
  se = warp_sum(se);
  if (lane == 0) s_val[warp] = se;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? s_val[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) s_bcast[0] = t;
  }
  __syncthreads();
  const float sum = s_bcast[0];
  __syncthreads();

  // pass 3: argmax over the bf16-rounded probabilities (what the reference's
  // next_score.argmax sees), with and without the interval id; first index wins ties.
  float pa = -1.f, pe = -1.f;
  int ia = 0x7fffffff, ie = 0x7fffffff;
  for (int i = tid; i < vocab; i += blockDim.x) {
    const float pr = bf16_round(expf(__bfloat162float(x[i]) - gmax) / sum);
    if (pr > pa) { pa = pr; ia = i; }
    const float pz = (i == interval_id) ? 0.f : pr;  // zero_() then argmax (demo/inference.py:77-79)
    if (pz > pe) { pe = pz; ie = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float opa = __shfl_xor_sync(0xffffffffu, pa, o), ope = __shfl_xor_sync(0xffffffffu, pe, o);
    const int oia = __shfl_xor_sync(0xffffffffu, ia, o), oie = __shfl_xor_sync(0xffffffffu, ie, o);
    if (opa > pa || (opa == pa && oia < ia)) { pa = opa; ia = oia; }
    if (ope > pe || (ope == pe && oie < ie)) { pe = ope; ie = oie; }
  }
  __shared__ float s_pa[32], s_pe[32];
  __shared__ int s_ia[32], s_ie[32];
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
      d.argmax_id = gidx;
      d.argmax_excl_id = ie;
      d.p_interval = (interval_id >= 0 && interval_id < vocab)
                         ? bf16_round(expf(__bfloat162float(x[interval_id]) - gmax) / sum)
                         : 0.f;
      d.max_logit = gmax;
      d.top2_margin = gmax - gmax2;
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
