// KV-append attention, tcgen05 version 3 ("tc2"): same contract and partial (m, l, O) workspace as attn_tc.cuh,
// re-laid-out so that a CTA keeps (almost) its whole share of the K/V stream in flight:
//
//   * P never touches shared memory.  The softmax threads write P_j as packed bf16 pairs straight into TMEM
//     (tcgen05.st, lane = query row, 32-bit column = two consecutive keys) and O += P_j V_j is issued with the
//     A operand IN TMEM (tcgen05.mma ... [d], [a], b-desc).  TMEM columns: S0 0..127 | S1 128..255 | O 256..383 |
//     P0 384..447 | P1 448..511 — nothing aliases.
//   * the 64 KB of shared memory this frees goes to the K/V ring: 192 KB = 3 stages of 128 keys (or 6 of 64),
//     ALL of them requested before the grid dependency resolves (K/V rows of earlier steps do not depend on any
//     kernel of this step), i.e. while the predecessor kernels (QKV fix-up, or the merge of the previous layer in
//     the micro-loop) are still running.  For the single-stream 12k-token shape that is 192 of the ~335 KB a CTA
//     consumes.
//   * flat grid: one CTA per (item, kv head, split) entry of a host-built table, splits of an item are as even as
//     the block count allows (blk0 = split * nblk / n_splits), and the host deals the splits so that the grid
//     fills the SMs: 8 kv heads x 18 splits = 144 CTAs for one stream (was 16 x 8 = 128 with equal-size splits).
//
//   * key slicing.  The softmax is MUFU-bound: one ex2 per (row, key), 4 lanes/clk per SM sub-partition, and a warp
//     instruction costs the same with 11 live lanes as with 32.  A frame step has 44 live rows (11 tokens x 4 heads),
//     an AR token 4, so with lane = row two (or three) of the four softmax warps idle while the others grind through
//     128 columns.  Here the Q tile is replicated SL = 128 / pad(rows) times down the 128 MMA rows (pad = 32 / 64 / 128)
//     and the softmax thread of lane = slice * pad + row handles only the 128 / SL keys of ITS slice of every block:
//     all four sub-partitions work, each thread issues 128 / SL exponentials per block.  P columns outside a thread's
//     slice stay zero (written once), so the unchanged O += P V accumulates per (slice, row) lane; the SL partial
//     (m, l, O) rows of a query row are combined through shared memory in the epilogue.
//
//   warp 0      TMA producer: Q tile, then the K and V tiles of every block (K and V have their own full/empty pairs)
//   warp 1      MMA issuer:   S_j = Q K_j^T -> TMEM S[j&1];  O += P_j V_j with P_j read from TMEM
//   warps 2-5   softmax:      tcgen05.ld S row -> online softmax (lazy O rescale) -> tcgen05.st P_j -> epilogue
#pragma once
#include <cuda.h>
#include "attn_tc.cuh"

namespace vlo {

template <int BLK>
struct Tc2Cfg {
  static constexpr int kSub = BLK * 128;                         // [BLK keys x 64 dims] bf16, 128B-swizzled
  static constexpr int kHalf = 2 * kSub;                         // one K or V tile
  static constexpr int kStageBytes = 2 * kHalf;                  // K + V of one block
  static constexpr int kStages = (192 * 1024) / kStageBytes;     // 3 (BLK = 128) / 6 (BLK = 64)
  static constexpr int kQBytes = 2 * 128 * 128;                  // Q tile: 128 rows x 128 dims bf16
  static constexpr int kSmemBytes = kQBytes + kStages * kStageBytes + 1024 + 512;
};
constexpr uint32_t kTc2ColS = 0, kTc2ColO = 256, kTc2ColP = 384;

struct AttnTc2Params {
  AttnParams base;
  const int* cta_tab;     // [gridDim.x]: item << 16 | kv head << 8 | split
  uint32_t v_lbo, v_sbo;  // V descriptor strides (bytes)
  long long* dbg;         // optional timeline buffer (VLO_ATTN_TRACE)
};

constexpr int kTc2CombStride = 132;   // floats per (slice, row) record in the combine buffer: O[128] | m | l | pad

struct Tc2Ctx {
  uint64_t *s_full, *s_empty, *p_full, *p_empty;
  uint32_t tS, tO, tP;
  float* comb;            // shared-memory combine buffer (the K/V ring, free by then)
  int nblk, blk0, G, kvh, split;
};

// Softmax / correction / epilogue role of one of the four softmax warps.  NC = keys per thread per block = 128 / SL;
// lane L of the 128 TMEM lanes serves row r = L % NC of the item's tile and key slice L / NC.
template <int NC>
__device__ __forceinline__ void tc2_softmax_role(const AttnTc2Params& pp, const AttnItem& it, const Tc2Ctx& cx, int warp,
                                                 int lane) {
  constexpr int SL = 128 / NC;
  const AttnParams& p = pp.base;
  const int q = warp & 3;
  const int L = q * 32 + lane;                 // TMEM lane
  const int slice = L / NC, r = L % NC;        // key slice, tile row == t * G + g
  const int rows = it.q_count * cx.G;
  const int t = r / cx.G;
  const bool valid = r < rows;
  const bool warp_live = ((q * 32) % NC) < rows;    // any query row in this warp's 32 lanes?
  const int lim = valid ? it.q_pos0 + t : -1;       // last visible key (causal with offset)
  const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
  const uint32_t col0 = static_cast<uint32_t>(slice * NC);
  const float c = p.scale_log2;
  const int nblk = cx.nblk;
  float m_ref = -INFINITY, l_run = 0.f;
  if (SL > 1 && warp_live) {   // P columns of the other slices stay zero for the whole kernel
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0u;
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 32) tmem_st_x32(cx.tP + lane_addr + c0, z);
    tmem_st_wait();
  }
  for (int j = 0; j < nblk; ++j) {
    const int b = j & 1;
    mbar_wait(&cx.s_full[b], (j >> 1) & 1);
    tc_fence_after();
    if (L == 0) VLO_TC_STAMP(2, 4 + 4 * j);
    if (!warp_live) {  // no query row in these lanes: keep the barrier protocol going (their P / O rows stay
                       // garbage; MMA rows are independent, they feed nothing but their own discarded O rows)
      mbar_arrive(&cx.s_empty[b]);
      mbar_wait(&cx.p_empty[b], ((j >> 1) & 1) ^ 1);
      mbar_arrive(&cx.p_full[b]);
      continue;
    }
    const int key0 = (cx.blk0 + j) * 128 + static_cast<int>(col0);
    const bool need_mask = key0 + NC - 1 > it.q_pos0;  // slice reaches past the first query's limit
    float sv[NC];
    {
      uint32_t su[NC];
#pragma unroll
      for (int c0 = 0; c0 < NC; c0 += 32)
        tmem_ld_x32(cx.tS + lane_addr + static_cast<uint32_t>(b * 128) + col0 + c0, *reinterpret_cast<uint32_t(*)[32]>(&su[c0]));
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < NC; ++i) sv[i] = __uint_as_float(su[i]);
    }
    tc_fence_before();
    mbar_arrive(&cx.s_empty[b]);   // S[b] may be overwritten by block j+2
    if (L == 0) VLO_TC_STAMP(2, 5 + 4 * j);
    if (need_mask) {
#pragma unroll
      for (int i = 0; i < NC; ++i)
        if (key0 + i > lim) sv[i] = -INFINITY;
    } else if (!valid) {
#pragma unroll
      for (int i = 0; i < NC; ++i) sv[i] = -INFINITY;
    }
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < NC; i += 4) {
      mx4[0] = fmaxf(mx4[0], sv[i]);
      mx4[1] = fmaxf(mx4[1], sv[i + 1]);
      mx4[2] = fmaxf(mx4[2], sv[i + 2]);
      mx4[3] = fmaxf(mx4[3], sv[i + 3]);
    }
    const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    // lazy rescale: keep the old reference max unless the new one is much larger
    const float m_new = fmaxf(m_ref, mx);
    const bool grow = (m_ref == -INFINITY) ? (m_new != -INFINITY) : ((m_new - m_ref) * c > kTcRescaleLog2);
    const float m_use = grow ? m_new : m_ref;
    const float alpha = (grow && m_ref != -INFINITY) ? exp2f((m_ref - m_new) * c) : 1.f;
    if (__any_sync(0xffffffffu, alpha != 1.f)) {
      // O of this warp's 32 lanes must be rescaled: wait until PV_{j-1} has landed in TMEM.  PV_i commits on
      // p_empty[i & 1]; this thread has observed that barrier up to PV_{j-3} (its wait at block j-1), so the phase of
      // PV_{j-1} is exactly one ahead: no parity aliasing (a free-running per-PV barrier would alias, see epilogue).
      mbar_wait(&cx.p_empty[b ^ 1], ((j - 1) >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(cx.tO + lane_addr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
        tmem_st_x32(cx.tO + lane_addr + c0, v);
      }
      tmem_st_wait();
      l_run *= alpha;
    }
    m_ref = m_use;
    const float moff = (m_ref == -INFINITY) ? 0.f : m_ref * c;
    // P = exp2(S c - m c) -> bf16 pairs -> TMEM; the P buffer must have been consumed by PV_{j-2} first
    mbar_wait(&cx.p_empty[b], ((j >> 1) & 1) ^ 1);
    tc_fence_after();
    if (L == 0) VLO_TC_STAMP(2, 6 + 4 * j);
    float ps4[4] = {0.f, 0.f, 0.f, 0.f};
    const float nmoff = -moff;
    const uint32_t pcol = cx.tP + lane_addr + static_cast<uint32_t>(b * 64) + col0 / 2;
    if constexpr (NC >= 64) {
#pragma unroll
      for (int c0 = 0; c0 < NC; c0 += 64) {
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float p0 = ex2_approx(fmaf(sv[c0 + 2 * i], c, nmoff));
          const float p1 = ex2_approx(fmaf(sv[c0 + 2 * i + 1], c, nmoff));
          ps4[i & 3] += p0 + p1;
          w[i] = pack_bf16(p0, p1);
        }
        tmem_st_x32(pcol + c0 / 2, w);
      }
    } else {
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = ex2_approx(fmaf(sv[2 * i], c, nmoff));
        const float p1 = ex2_approx(fmaf(sv[2 * i + 1], c, nmoff));
        ps4[i & 3] += p0 + p1;
        w[i] = pack_bf16(p0, p1);
      }
      tmem_st_x16(pcol, w);
    }
    l_run += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(&cx.p_full[b]);
    if (L == 0) VLO_TC_STAMP(2, 7 + 4 * j);
  }
  // ---- epilogue: partial (m, l, O) of this split; merged by attn_merge_kernel.
  // Wait for the LAST PV through its p_empty commit: every softmax thread has followed p_empty[b] phase by phase (it
  // waits for PV_{j-2} at block j), so PV_{nblk-1} is exactly the next phase.  A barrier that completes one phase per PV
  // and is only looked at here would alias: parity (nblk-1)&1 is also the parity of PV_{nblk-3}, and with the
  // key-sliced softmax running ahead of the V stream the wait returned two PVs early.
  mbar_wait(&cx.p_empty[(nblk - 1) & 1], ((nblk - 1) >> 1) & 1);
  tc_fence_after();
  if (L == 0) VLO_TC_STAMP(2, 1);
  float w_self = 1.f;
  if constexpr (SL > 1) {
    // slices >= 1 park their (O, m, l) rows in shared memory; slice 0 folds them into its own row
    float* rec = cx.comb + static_cast<size_t>((slice > 0 ? slice - 1 : 0) * NC + r) * kTc2CombStride;
    // NOTE: tcgen05.ld is .sync.aligned - every lane of the warp must execute it: the TMEM loads are guarded by
    // warp-uniform conditions only (slice, warp_live); `valid` guards nothing but the stores.
    if (slice > 0 && warp_live) {
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(cx.tO + lane_addr + c0, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(rec + c0 + 4 * i) = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                                        __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
      }
      if (valid) {
        rec[128] = m_ref;
        rec[129] = l_run;
      }
    }
    asm volatile("bar.sync 1, 128;\n" ::: "memory");   // the four softmax warps
    if (slice == 0 && valid) {
      float m_all = m_ref;
#pragma unroll
      for (int s2 = 1; s2 < SL; ++s2) m_all = fmaxf(m_all, cx.comb[static_cast<size_t>((s2 - 1) * NC + r) * kTc2CombStride + 128]);
      w_self = (m_ref == -INFINITY) ? 0.f : exp2f((m_ref - m_all) * c);
      l_run *= w_self;
#pragma unroll
      for (int s2 = 1; s2 < SL; ++s2) {
        const float* o2 = cx.comb + static_cast<size_t>((s2 - 1) * NC + r) * kTc2CombStride;
        const float m2 = o2[128];
        l_run += (m2 == -INFINITY) ? 0.f : o2[129] * exp2f((m2 - m_all) * c);
      }
      m_ref = m_all;
    }
  }
  const size_t slot = static_cast<size_t>(it.ws_slot0) + (static_cast<size_t>(cx.kvh) * it.n_splits + cx.split) * rows + r;
  if (slice == 0 && warp_live) {   // warp-uniform guard around the TMEM loads (see above)
    float* dst = p.ws_o + slot * kAttnHD;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      tmem_ld_x32(cx.tO + lane_addr + c0, v);
      tmem_ld_wait();
      if (valid) {
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(v[i]) * w_self;
        if constexpr (SL > 1) {
#pragma unroll
          for (int s2 = 1; s2 < SL; ++s2) {
            const float* o2 = cx.comb + static_cast<size_t>((s2 - 1) * NC + r) * kTc2CombStride;
            const float m2 = o2[128];
            const float w2 = (m2 == -INFINITY) ? 0.f : exp2f((m2 - m_ref) * c);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 x = *reinterpret_cast<const float4*>(o2 + c0 + 4 * i);
              o[4 * i] = fmaf(x.x, w2, o[4 * i]);
              o[4 * i + 1] = fmaf(x.y, w2, o[4 * i + 1]);
              o[4 * i + 2] = fmaf(x.z, w2, o[4 * i + 2]);
              o[4 * i + 3] = fmaf(x.w, w2, o[4 * i + 3]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(dst + c0 + 4 * i) = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
      }
    }
    if (valid) {
      p.ws_ml[slot * 2] = m_ref;
      p.ws_ml[slot * 2 + 1] = l_run;
    }
  }
  if (L == 0) VLO_TC_STAMP(2, 2);
  tc_fence_before();
}

// grid = (n_ctas); block = 192.  tm_q32 / tm_q64 / tm_q128: Q maps whose box holds 32 / 64 / 128 tile rows.
template <int BLK>
__global__ void __launch_bounds__(kTcThreads, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                const __grid_constant__ CUtensorMap tm_q32, const __grid_constant__ CUtensorMap tm_q64,
                const __grid_constant__ CUtensorMap tm_q128, const AttnTc2Params pp) {
  static_assert(BLK == 128, "key slicing assumes 128-key blocks");
  using C = Tc2Cfg<BLK>;
  constexpr int NS = C::kStages;
  const AttnParams& p = pp.base;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* q_tile = smem;
  uint8_t* kv_tile = smem + C::kQBytes;                     // NS stages: K tile | V tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_tile + NS * C::kStageBytes);
  uint64_t* k_full = bars;                   // [NS]  TMA -> MMA (K tile landed)
  uint64_t* k_empty = k_full + NS;           // [NS]  S_j done -> TMA
  uint64_t* v_full = k_empty + NS;           // [NS]  TMA -> MMA (V tile landed)
  uint64_t* v_empty = v_full + NS;           // [NS]  PV_j done -> TMA
  uint64_t* s_full = v_empty + NS;           // [2]   S_j in TMEM
  uint64_t* s_empty = s_full + 2;            // [2]   softmax has read S_j
  uint64_t* p_full = s_empty + 2;            // [2]   P_j in TMEM (and O rescaled)
  uint64_t* p_empty = p_full + 2;            // [2]   PV_j has consumed P buffer
  uint64_t* q_ready = p_empty + 2;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // The CTA / item tables were uploaded by a memcpy at the start of the step (a full stream-order dependency of the
  // step's first kernel), so they may be read before pdl_wait().
  const int code = pp.cta_tab[blockIdx.x];
  const int split = code & 255, kvh = (code >> 8) & 255;
  const AttnItem it = p.items[code >> 16];
  if (threadIdx.x == 0) VLO_TC_STAMP(0, 0);
  if (threadIdx.x == 0 && pp.dbg != nullptr) pp.dbg[static_cast<size_t>(blockIdx.x) * 192 + 60] = static_cast<long long>(globaltimer_ns());
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_q32);
    tma_prefetch_desc(&tm_q64);
    tma_prefetch_desc(&tm_q128);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(q_ready, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();  // after the TMEM allocation (see gemm_ws.cuh)

  const int G = p.n_heads / p.n_kv_heads;
  // key slicing: pad = rows of one Q-tile copy (32 / 64 / 128, >= the live rows and >= G), SL = 128 / pad copies
  const int rows_live = it.q_count * G;
  const int need_rows = rows_live > G ? rows_live : G;
  const int pad = need_rows <= 32 ? 32 : (need_rows <= 64 ? 64 : 128);
  const int kv_end = it.q_pos0 + it.q_count;
  const int nblk_total = (kv_end + BLK - 1) / BLK;
  const int blk0 = static_cast<int>((static_cast<long long>(split) * nblk_total) / it.n_splits);
  const int nblk = static_cast<int>((static_cast<long long>(split + 1) * nblk_total) / it.n_splits) - blk0;
  const uint32_t tS = tmem_base + kTc2ColS, tO = tmem_base + kTc2ColO, tP = tmem_base + kTc2ColP;
  const int row_base = it.kv_row0 + kvh * it.kv_head_stride;

  auto load_k = [&](int j) {
    const int s = j % NS;
    uint8_t* st = kv_tile + s * C::kStageBytes;
    const int row = row_base + (blk0 + j) * BLK;
    mbar_arrive_expect_tx(&k_full[s], C::kHalf);
    tma_load_2d(st, &tm_k, &k_full[s], 0, row, kEvictFirst);
    tma_load_2d(st + C::kSub, &tm_k, &k_full[s], 64, row, kEvictFirst);
  };
  auto load_v = [&](int j) {
    const int s = j % NS;
    uint8_t* st = kv_tile + s * C::kStageBytes + C::kHalf;
    const int row = row_base + (blk0 + j) * BLK;
    mbar_arrive_expect_tx(&v_full[s], C::kHalf);
    tma_load_2d(st, &tm_v, &v_full[s], 0, row, kEvictFirst);
    tma_load_2d(st + C::kSub, &tm_v, &v_full[s], 64, row, kEvictFirst);
  };

  // K/V rows of EARLIER steps do not depend on any kernel of this step: fill the whole ring with the blocks that end
  // safely below this step's new tokens before waiting on the grid dependency.
  int pre = 0;
  if (warp == 0 && lane == 0) {
    const int safe_end = it.q_pos0 - 128;  // a step appends at most 128 tokens per stream
    for (; pre < nblk && pre < NS; ++pre) {
      if ((blk0 + pre) * BLK + BLK - 1 >= safe_end) break;
      load_k(pre);
      load_v(pre);
    }
  }
  if (threadIdx.x == 0) VLO_TC_STAMP(0, 1);
  pdl_wait();
  if (threadIdx.x == 0) VLO_TC_STAMP(0, 2);

  if (nblk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        // ------------------------------------------------------------ TMA producer
        // Q tile: 128 / pad copies of the pad rows (row = token * G + head), one 3-D box per 64-dim half and copy
        const CUtensorMap* tq = pad == 32 ? &tm_q32 : (pad == 64 ? &tm_q64 : &tm_q128);
        mbar_arrive_expect_tx(q_ready, C::kQBytes);
        for (int cp = 0; cp < 128 / pad; ++cp) {
          tma_load_3d(q_tile + cp * pad * 128, tq, q_ready, 0, kvh * G, it.q_tok0, kEvictNormal);
          tma_load_3d(q_tile + kTcSub + cp * pad * 128, tq, q_ready, 64, kvh * G, it.q_tok0, kEvictNormal);
        }
        for (int j = pre; j < nblk; ++j) {
          const int s = j % NS;
          const uint32_t ph = (j / NS) & 1;
          mbar_wait(&k_empty[s], ph ^ 1);
          VLO_TC_STAMP(0, 4 + 2 * j);
          load_k(j);
          mbar_wait(&v_empty[s], ph ^ 1);
          VLO_TC_STAMP(0, 5 + 2 * j);
          load_v(j);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, BLK, 0);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, 128, 1);
        const uint32_t q_addr = smem_u32(q_tile);
        mbar_wait(q_ready, 0);
        tc_fence_after();
        auto issue_pv = [&](int i) {
          const int s = i % NS;
          const int b = i & 1;
          mbar_wait(&v_full[s], (i / NS) & 1);
          mbar_wait(&p_full[b], (i >> 1) & 1);
          tc_fence_after();
          VLO_TC_STAMP(1, 3 * i + 2);
          const uint32_t v_addr = smem_u32(kv_tile + s * C::kStageBytes + C::kHalf);
#pragma unroll
          for (int kk = 0; kk < BLK / 16; ++kk) {  // 16 keys per MMA: 8 packed TMEM columns of P
            const uint64_t db = umma_desc_mn_sw128(v_addr + kk * 16 * 128, pp.v_lbo, pp.v_sbo);
            umma_f16_ts(tO, tP + static_cast<uint32_t>(b * 64 + kk * 8), db, idesc_o, (i > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&v_empty[s]);
          umma_commit(&p_empty[b]);
        };
        for (int j = 0; j < nblk; ++j) {
          const int s = j % NS;
          const int b = j & 1;
          mbar_wait(&k_full[s], (j / NS) & 1);
          VLO_TC_STAMP(1, 3 * j);
          mbar_wait(&s_empty[b], ((j >> 1) & 1) ^ 1);
          tc_fence_after();
          VLO_TC_STAMP(1, 3 * j + 1);
          const uint32_t k_addr = smem_u32(kv_tile + s * C::kStageBytes);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {  // 16 dims per MMA
            const uint64_t da = umma_desc_sw128(q_addr + (kk >> 2) * kTcSub + (kk & 3) * 32);
            const uint64_t db = umma_desc_sw128(k_addr + (kk >> 2) * C::kSub + (kk & 3) * 32);
            umma_f16(tS + b * 128, da, db, idesc_s, kk > 0 ? 1u : 0u);
          }
          umma_commit(&k_empty[s]);  // K tile is free as soon as QK^T has retired
          umma_commit(&s_full[b]);
          if (j >= 1) issue_pv(j - 1);
        }
        issue_pv(nblk - 1);
      }
    } else {
      // -------------------------------------------------------------- softmax / correction / epilogue warps
      Tc2Ctx cx;
      cx.s_full = s_full, cx.s_empty = s_empty, cx.p_full = p_full, cx.p_empty = p_empty;
      cx.tS = tS, cx.tO = tO, cx.tP = tP;
      cx.comb = reinterpret_cast<float*>(kv_tile);
      cx.nblk = nblk, cx.blk0 = blk0, cx.G = G, cx.kvh = kvh, cx.split = split;
      if (pad == 32) tc2_softmax_role<32>(pp, it, cx, warp, lane);
      else if (pad == 64) tc2_softmax_role<64>(pp, it, cx, warp, lane);
      else tc2_softmax_role<128>(pp, it, cx, warp, lane);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && pp.dbg != nullptr) {
    pp.dbg[static_cast<size_t>(blockIdx.x) * 192 + 61] = static_cast<long long>(globaltimer_ns());
    pp.dbg[static_cast<size_t>(blockIdx.x) * 192 + 62] = clock64();
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vlo
