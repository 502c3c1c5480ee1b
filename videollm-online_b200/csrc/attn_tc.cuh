// KV-append attention, tcgen05 version (v2) — same contract, work items, split-KV plan and partial
// (m, l, O) workspace as attn.cuh, different engine room:
//
//   S = Q K^T and O += P V run on the 5th-gen tensor cores (tcgen05.mma, M = 128, N = 128, fp32 accumulators in
//   TMEM), so K and V are read from shared memory exactly once by the MMA unit instead of 3-4x through ldmatrix.
//   The <= 128/G query tokens x G heads of a kv head occupy TMEM lanes  lane = t * G + g  (dense: the same row
//   order as the partial workspace), one softmax thread per lane: row max / exp2 / row sum need no shuffles; warps
//   whose 32 lanes hold no query row skip the softmax arithmetic.  The Q tile is ONE 3-D TMA box per 64-d half
//   (tmap_q3d_sw128), issued by the producer as soon as the grid dependency resolves.
//
//   warp 0      TMA producer: Q tile, then per 128-key block one stage = K [128 keys x 128 d] + V [128 keys x 128 d] (64 KB), 2 stages
//   warp 1      MMA issuer:   S_j -> TMEM S[j&1];  after P_j is staged: O += P_j V_j      (software-pipelined: S_{j+1} first)
//   warps 2-5   softmax:      tcgen05.ld S row -> online softmax with lazy (threshold) rescaling of O in TMEM ->
//                             P_j as bf16 into a 128B-swizzled K-major smem tile (double buffered) -> epilogue
//   Measured alternatives (same box, 12k keys, q = 11): two warps per lane quadrant splitting the S row (15.3 vs 15.6 us)
//   and two ping-pong softmax groups with separate O accumulators (16.5 us) do not pay: the steady state of this
//   kernel already streams K/V at ~5.7 TB/s; what is left is ramp (Q staging after the grid dependency) and tail.
//   K tile: K-major B operand (same descriptor as the GEMMs).  V tile: rows = keys, d contiguous -> MN-major B operand.
#pragma once
#include <cuda.h>
#include "attn.cuh"
#include "ptx.cuh"
#include "tc_helpers.cuh"

namespace vlo {

constexpr int kTcBlk = 128;                          // keys per block
constexpr int kTcStages = 2;
constexpr int kTcSub = kTcBlk * 128;                 // one [128 rows x 64 elem] swizzled sub-tile = 16 KB
constexpr int kTcStageBytes = 4 * kTcSub;            // K(2 d-halves) + V(2 d-halves) = 64 KB (K and V rings, 32 KB slots)
constexpr int kTcHalf = 2 * kTcSub;                  // one K or V tile: 32 KB
constexpr int kTcQBytes = 2 * kTcSub;                // Q tile 32 KB
constexpr int kTcPBytes = 2 * kTcSub;                // one P buffer 32 KB
constexpr int kTcSmemBytes = kTcQBytes + 2 * kTcPBytes + kTcStages * kTcStageBytes + 1024 + 256;
constexpr int kTcThreads = 192;

struct AttnTcParams {
  AttnParams base;
  uint32_t v_lbo, v_sbo;  // V descriptor strides (bytes)
  long long* dbg;         // optional timeline buffer (VLO_ATTN_TRACE): [cta][role][64] clock64 stamps
};
#define VLO_TC_STAMP(role, idx)                                                                          \
  do {                                                                                                   \
    if (pp.dbg != nullptr && (idx) < 64)                                                                 \
      pp.dbg[((static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 192 + \
             (role) * 64 + (idx)] = clock64();                                                           \
  } while (0)

// swizzled byte offset of 16-byte chunk `c16` (0..15 along the 128-element row) of row `r` in a
// [2 sub-tiles][128 rows][128 B] K-major tile
__device__ __forceinline__ uint32_t tc_sw_off(int r, int c16) {
  return static_cast<uint32_t>((c16 >> 3) * kTcSub + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
}

// grid = (max_splits, n_kv_heads, n_items); block = 192.
__global__ void __launch_bounds__(kTcThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
               const __grid_constant__ CUtensorMap tm_q, const AttnTcParams pp) {
  const AttnParams& p = pp.base;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* q_tile = smem;
  uint8_t* p_tile = smem + kTcQBytes;                       // 2 buffers
  uint8_t* kv_tile = p_tile + 2 * kTcPBytes;                // kTcStages stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_tile + kTcStages * kTcStageBytes);
  uint64_t* k_full = bars;                   // [stages]  TMA -> MMA (K tile landed)
  uint64_t* k_empty = k_full + kTcStages;    // [stages]  S_j done -> TMA   (K freed as soon as QK^T has run)
  uint64_t* v_full = k_empty + kTcStages;    // [stages]  TMA -> MMA (V tile landed)
  uint64_t* v_empty = v_full + kTcStages;    // [stages]  PV_j done -> TMA
  uint64_t* s_full = v_empty + kTcStages;    // [2]       S_j in TMEM
  uint64_t* s_empty = s_full + 2;            // [2]       softmax has read S_j
  uint64_t* p_full = s_empty + 2;            // [2]       P_j staged in smem (and O rescaled)
  uint64_t* p_empty = p_full + 2;            // [2]       PV_j done reading P buffer
  uint64_t* o_done = p_empty + 2;            // [1]       PV_j complete (phase j)
  uint64_t* q_ready = o_done + 1;            // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Work item.  The item table was uploaded by a memcpy at the start of the step; a memcpy is a full stream-order
  // dependency for the first kernel of the step, so it is complete before ANY kernel of the step runs and may be
  // read before pdl_wait().  Issued first so its latency hides behind the barrier / TMEM set-up.
  const AttnItem it = p.items[blockIdx.z];
  if (threadIdx.x == 0) VLO_TC_STAMP(0, 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_q);
    for (int i = 0; i < kTcStages; ++i) {
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
    mbar_init(o_done, 1);
    mbar_init(q_ready, 1);   // TMA transaction barrier
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();  // after the TMEM allocation (see gemm_ws.cuh); successors may start their prologues

  // split geometry
  const int split = blockIdx.x;
  const int kvh = blockIdx.y;
  const int G = p.n_heads / p.n_kv_heads;
  const int kv_end = it.q_pos0 + it.q_count;
  const int nblk_total = (kv_end + kTcBlk - 1) / kTcBlk;
  const int blk0 = split * it.blocks_per_split;
  const int nblk = (split < it.n_splits) ? min(blk0 + it.blocks_per_split, nblk_total) - blk0 : 0;
  const uint32_t tS = tmem_base, tO = tmem_base + 256;
  const int row_base = it.kv_row0 + kvh * it.kv_head_stride;

  // K/V rows of EARLIER steps do not depend on any kernel of this step (only the rows appended in this step
  // do; the previous step finished before this step's first memcpy): prefetch the first blocks that end safely
  // below this step's new tokens before waiting.
  int pre = 0;
  if (warp == 0 && lane == 0) {
    const int safe_end = it.q_pos0 - 128;  // a step appends at most 128 tokens per stream
    for (; pre < nblk && pre < kTcStages; ++pre) {
      if ((blk0 + pre) * kTcBlk + kTcBlk - 1 >= safe_end) break;
      uint8_t* st = kv_tile + pre * kTcStageBytes;
      const int row = row_base + (blk0 + pre) * kTcBlk;
      mbar_arrive_expect_tx(&k_full[pre], kTcHalf);
      tma_load_2d(st, &tm_k, &k_full[pre], 0, row, kEvictFirst);
      tma_load_2d(st + kTcSub, &tm_k, &k_full[pre], 64, row, kEvictFirst);
      mbar_arrive_expect_tx(&v_full[pre], kTcHalf);
      tma_load_2d(st + 2 * kTcSub, &tm_v, &v_full[pre], 0, row, kEvictFirst);
      tma_load_2d(st + 3 * kTcSub, &tm_v, &v_full[pre], 64, row, kEvictFirst);
    }
  }
  if (threadIdx.x == 0) VLO_TC_STAMP(0, 1);
  pdl_wait();
  if (threadIdx.x == 0) VLO_TC_STAMP(0, 2);

  if (nblk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        // ------------------------------------------------------------ TMA producer
        // Q rows of this kv head (written by the predecessor kernel): two 16 KB boxes, rows = token * G + head
        mbar_arrive_expect_tx(q_ready, kTcQBytes);
        tma_load_3d(q_tile, &tm_q, q_ready, 0, kvh * G, it.q_tok0, kEvictNormal);
        tma_load_3d(q_tile + kTcSub, &tm_q, q_ready, 64, kvh * G, it.q_tok0, kEvictNormal);
        for (int j = pre; j < nblk; ++j) {
          const int s = j % kTcStages;
          const uint32_t ph = (j / kTcStages) & 1;
          uint8_t* st = kv_tile + s * kTcStageBytes;
          const int row = row_base + (blk0 + j) * kTcBlk;
          mbar_wait(&k_empty[s], ph ^ 1);
          VLO_TC_STAMP(0, 4 + 2 * j);
          mbar_arrive_expect_tx(&k_full[s], kTcHalf);
          tma_load_2d(st, &tm_k, &k_full[s], 0, row, kEvictFirst);
          tma_load_2d(st + kTcSub, &tm_k, &k_full[s], 64, row, kEvictFirst);
          mbar_wait(&v_empty[s], ph ^ 1);
          VLO_TC_STAMP(0, 5 + 2 * j);
          mbar_arrive_expect_tx(&v_full[s], kTcHalf);
          tma_load_2d(st + 2 * kTcSub, &tm_v, &v_full[s], 0, row, kEvictFirst);
          tma_load_2d(st + 3 * kTcSub, &tm_v, &v_full[s], 64, row, kEvictFirst);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, 128, 1);
        const uint32_t q_addr = smem_u32(q_tile);
        mbar_wait(q_ready, 0);
        tc_fence_after();
        auto issue_pv = [&](int i) {
          const int s = i % kTcStages;
          const int b = i & 1;
          mbar_wait(&v_full[s], (i / kTcStages) & 1);
          mbar_wait(&p_full[b], (i >> 1) & 1);
          tc_fence_after();
          VLO_TC_STAMP(1, 3 * i + 2);
          const uint32_t p_addr = smem_u32(p_tile + b * kTcPBytes);
          const uint32_t v_addr = smem_u32(kv_tile + s * kTcStageBytes + 2 * kTcSub);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {  // 16 keys per MMA
            const uint64_t da = umma_desc_sw128(p_addr + (kk >> 2) * kTcSub + (kk & 3) * 32);
            const uint64_t db = umma_desc_mn_sw128(v_addr + kk * 16 * 128, pp.v_lbo, pp.v_sbo);
            umma_f16(tO, da, db, idesc_o, (i > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&v_empty[s]);
          umma_commit(&p_empty[b]);
          umma_commit(o_done);
        };
        for (int j = 0; j < nblk; ++j) {
          const int s = j % kTcStages;
          const int b = j & 1;
          mbar_wait(&k_full[s], (j / kTcStages) & 1);
          VLO_TC_STAMP(1, 3 * j);
          mbar_wait(&s_empty[b], ((j >> 1) & 1) ^ 1);
          tc_fence_after();
          VLO_TC_STAMP(1, 3 * j + 1);
          const uint32_t k_addr = smem_u32(kv_tile + s * kTcStageBytes);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {  // 16 dims per MMA
            const uint64_t da = umma_desc_sw128(q_addr + (kk >> 2) * kTcSub + (kk & 3) * 32);
            const uint64_t db = umma_desc_sw128(k_addr + (kk >> 2) * kTcSub + (kk & 3) * 32);
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
      const int q = warp & 3;
      const int r = q * 32 + lane;       // TMEM lane == tile row == t * G + g
      const int t = r / G;
      const bool valid = r < it.q_count * G;
      const bool warp_live = q * 32 < it.q_count * G;   // any query row in this warp's 32 lanes?
      const int lim = valid ? it.q_pos0 + t : -1;       // last visible key (causal with offset)
      const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
      const float c = p.scale_log2;
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nblk; ++j) {
        const int b = j & 1;
        mbar_wait(&s_full[b], (j >> 1) & 1);
        tc_fence_after();
        if (r == 0) VLO_TC_STAMP(2, 4 + 4 * j);
        if (!warp_live) {  // no query row in these lanes: only keep the barrier protocol going (P/O rows stay garbage,
                           // they feed nothing but their own discarded O rows)
          tc_fence_before();
          mbar_arrive(&s_empty[b]);
          mbar_wait(&p_empty[b], ((j >> 1) & 1) ^ 1);
          fence_proxy_async();
          mbar_arrive(&p_full[b]);
          continue;
        }
        const int key0 = (blk0 + j) * kTcBlk;
        const bool need_mask = key0 + kTcBlk - 1 > it.q_pos0;  // block reaches past the first query's limit
        // S row -> registers (128 fp32), masked, row max
        float sv[128];
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t v[32];
          tmem_ld_x32(tS + lane_addr + b * 128 + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) sv[c0 + i] = __uint_as_float(v[i]);
        }
        tc_fence_before();
        mbar_arrive(&s_empty[b]);   // S[b] may be overwritten by block j+2
        if (r == 0) VLO_TC_STAMP(2, 5 + 4 * j);
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (key0 + i > lim) sv[i] = -INFINITY;
        } else if (!valid) {
#pragma unroll
          for (int i = 0; i < 128; ++i) sv[i] = -INFINITY;
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // 4 independent chains
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
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
          // O of this warp's 32 lanes must be rescaled: wait until PV_{j-1} has landed in TMEM (through its p_empty
          // commit, which this thread follows phase by phase; a free-running per-PV barrier aliases, see attn_tc2.cuh)
          mbar_wait(&p_empty[b ^ 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld_x32(tO + lane_addr + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_x32(tO + lane_addr + c0, v);
          }
          tmem_st_wait();
          tc_fence_before();
          l_run *= alpha;
        }
        m_ref = m_use;
        const float moff = (m_ref == -INFINITY) ? 0.f : m_ref * c;
        // P = exp2(S c - m c) -> bf16 -> swizzled smem tile; wait for the P buffer to be free first
        mbar_wait(&p_empty[b], ((j >> 1) & 1) ^ 1);
        if (r == 0) VLO_TC_STAMP(2, 6 + 4 * j);
        uint8_t* pt = p_tile + b * kTcPBytes;
        float ps4[4] = {0.f, 0.f, 0.f, 0.f};
        const float nmoff = -moff;
#pragma unroll
        for (int c16 = 0; c16 < 16; ++c16) {
          uint32_t w[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float p0 = ex2_approx(fmaf(sv[c16 * 8 + 2 * k], c, nmoff));
            const float p1 = ex2_approx(fmaf(sv[c16 * 8 + 2 * k + 1], c, nmoff));
            ps4[k] += p0 + p1;
            w[k] = pack_bf16(p0, p1);
          }
          *reinterpret_cast<uint4*>(pt + tc_sw_off(r, c16)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        const float ps = (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
        l_run += ps;
        fence_proxy_async();        // make the generic-proxy P writes visible to the MMA (async proxy)
        mbar_arrive(&p_full[b]);
        if (r == 0) VLO_TC_STAMP(2, 7 + 4 * j);
      }
      // ---- epilogue: partial (m, l, O) of this split; merged by attn_merge_kernel
      mbar_wait(&p_empty[(nblk - 1) & 1], ((nblk - 1) >> 1) & 1);   // the last PV (see attn_tc2.cuh on parity aliasing)
      tc_fence_after();
      if (r == 0) VLO_TC_STAMP(2, 1);
      const int rows = it.q_count * G;
      const size_t slot = static_cast<size_t>(it.ws_slot0) + (static_cast<size_t>(kvh) * it.n_splits + split) * rows + r;
      if (warp_live) {
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 64) {   // two 64-column reads in flight, then the stores
          uint32_t v0[32], v1[32];
          tmem_ld_x32(tO + lane_addr + c0, v0);
          tmem_ld_x32(tO + lane_addr + c0 + 32, v1);
          tmem_ld_wait();
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(p.ws_o + slot * kAttnHD + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              dst[i] = make_float4(__uint_as_float(v0[4 * i]), __uint_as_float(v0[4 * i + 1]), __uint_as_float(v0[4 * i + 2]),
                                   __uint_as_float(v0[4 * i + 3]));
#pragma unroll
            for (int i = 0; i < 8; ++i)
              dst[8 + i] = make_float4(__uint_as_float(v1[4 * i]), __uint_as_float(v1[4 * i + 1]), __uint_as_float(v1[4 * i + 2]),
                                       __uint_as_float(v1[4 * i + 3]));
          }
        }
        if (valid) {
          p.ws_ml[slot * 2] = m_ref;
          p.ws_ml[slot * 2 + 1] = l_run;
        }
      }
      if (r == 0) VLO_TC_STAMP(2, 2);
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vlo
