// ViT self-attention on tcgen05, second generation: TWO query tiles in flight per CTA (ping-pong) and the K / V of the
// (frame, head) loaded ONCE for all the query tiles the CTA owns.
//   softmax(Q K^T / sqrt(64)) V, non-causal, fp16 operands / fp32 accumulation, head_dim 64, N = 576 tokens per frame
//   (HF:models/siglip/modeling_siglip.py:275-312).
//
// Why: the one-tile kernel (vit_attn_tc.cuh) is a serial chain per CTA - S MMA -> 128 exponentials per thread -> P -> PV MMA -
// with the tensor pipe idle during the exponentials and the MUFU idle during the MMAs, one CTA per SM: 15 % tensor-pipe
// utilisation, 52 us per layer at batch 8 (profiles/r02_vit_b8_singlecta_metrics.csv).  Here two softmax groups (A, B) of
// 128 threads each own a query tile; the single MMA thread serves both, so the MMAs / TMEM traffic of one tile run
// under the exponentials of the other, and a CTA walks through `tiles_per_cta` query tiles (rounds of two) with the
// 160 KB of K / V resident.  The kernel is then MUFU-bound (one ex2 per score).
//
//   warp 0      TMA producer: all K / V blocks up front; the Q tile of (round, group) when that group's Q buffer is free
//   warp 1      MMA issuer:   event-driven over both tiles: S(j) when the tile's softmax group has read S(j-1), PV(j) when P(j) is written
//   warps 2-5   softmax group A, warps 6-9 softmax group B: thread = query row (TMEM lane)
//   TMEM: group g at columns g*256: S 0..127 | O 128..191 | P 192..255 (fp16 pairs).
// qkv: [B*N, 3C] fp16 (q | k | v column blocks, head h at column h*64 inside each); out: [B*N, C] fp16.
#pragma once
#include <cuda.h>
#include "tc_helpers.cuh"
#include "vit_attn_tc.cuh"

namespace vlo {

constexpr int kVit2Threads = 320;
constexpr int kVit2Tile = 128 * 64 * 2;                  // [128 rows x 64 d] fp16, 128B-swizzled: 16 KB
constexpr int kVit2MaxBlk = 5;                           // resident key blocks: N <= 640
constexpr int kVit2SmemBytes = kVit2Tile * (2 + 2 * kVit2MaxBlk) + 1024 + 512;

struct Vit2Bars {
  uint64_t k_full[kVit2MaxBlk], v_full[kVit2MaxBlk];
  uint64_t q_full[2], q_empty[2], s_full[2], s_empty[2], p_full[2], p_empty[2], o_free[2];
};

__global__ void __launch_bounds__(kVit2Threads, 1)
vit_attn_tc2_kernel(const __grid_constant__ CUtensorMap tm_qkv, __half* out, int N, int C, float scale_log2, int tiles_per_cta) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* q_tile = smem;                                  // group g at g * 16 KB
  uint8_t* kv_tile = smem + 2 * kVit2Tile;                 // block j: K at 2j, V at 2j + 1
  Vit2Bars* bars = reinterpret_cast<Vit2Bars*>(kv_tile + 2 * kVit2MaxBlk * kVit2Tile);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);

  const int head = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = (N + 127) / 128;                        // key blocks == query tiles of a frame (host: nblk <= kVit2MaxBlk)
  const int t0 = blockIdx.x * tiles_per_cta;
  const int n_my = min(tiles_per_cta, nblk - t0);          // query tiles of this CTA (>= 1 by grid construction)
  const int n_rounds = (n_my + 1) >> 1;
  const int row_base = b * N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    for (int i = 0; i < kVit2MaxBlk; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->v_full[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bars->q_full[g], 1);
      mbar_init(&bars->q_empty[g], 1);
      mbar_init(&bars->s_full[g], 1);
      mbar_init(&bars->s_empty[g], 128);
      mbar_init(&bars->p_full[g], 128);
      mbar_init(&bars->p_empty[g], 1);
      mbar_init(&bars->o_free[g], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();   // q | k | v come from the QKV GEMM right before

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      auto load_q = [&](int r, int g) {
        const int tile = t0 + 2 * r + g;
        mbar_wait(&bars->q_empty[g], (r & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->q_full[g], kVit2Tile);
        tma_load_2d(q_tile + g * kVit2Tile, &tm_qkv, &bars->q_full[g], head * 64, row_base + tile * 128, kEvictNormal);
      };
      load_q(0, 0);
      if (n_my > 1) load_q(0, 1);
      for (int j = 0; j < nblk; ++j) {
        mbar_arrive_expect_tx(&bars->k_full[j], kVit2Tile);
        tma_load_2d(kv_tile + (2 * j) * kVit2Tile, &tm_qkv, &bars->k_full[j], C + head * 64, row_base + j * 128, kEvictNormal);
      }
      for (int j = 0; j < nblk; ++j) {
        mbar_arrive_expect_tx(&bars->v_full[j], kVit2Tile);
        tma_load_2d(kv_tile + (2 * j + 1) * kVit2Tile, &tm_qkv, &bars->v_full[j], 2 * C + head * 64, row_base + j * 128, kEvictNormal);
      }
      for (int r = 1; r < n_rounds; ++r) {
        load_q(r, 0);
        if (2 * r + 1 < n_my) load_q(r, 1);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // -------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc_s = umma_idesc_f16in(128, 128, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16in(128, 64, 1);
      // Event-driven issue: whichever MMA group of either query tile has its operands ready goes next (non-blocking
      // barrier probes), so the two tiles drift half a block apart and the TMEM loads / stores / barrier hand-offs of one
      // tile run under the exponentials of the other.  (A static order S_A, S_B, PV_A, PV_B keeps both softmax groups in
      // the same phase: 2.6 us per key block for the pair instead of ~1.3.)  The second tile's very first S waits until the
      // first tile's S has been read, which sets the offset.
      for (int r = 0; r < n_rounds; ++r) {
        const int ng = (2 * r + 1 < n_my) ? 2 : 1;       // groups with a tile in this round
        for (int g = 0; g < ng; ++g) mbar_wait(&bars->q_full[g], r & 1);
        tc_fence_after();
        int js[2] = {0, 0}, jp[2] = {0, 0};
        uint32_t idle = 0;
        while (jp[0] < nblk || (ng == 2 && jp[1] < nblk)) {
          bool progress = false;
          for (int g = 0; g < ng; ++g) {
            if (jp[g] < js[g]) {                          // PV(jp): S(jp) was issued; needs P(jp), V(jp) (and a free O at block 0)
              const int i = jp[g], pc = r * nblk + i;
              if (mbar_test_wait(&bars->p_full[g], pc & 1) && mbar_test_wait(&bars->v_full[i], 0) &&
                  (i > 0 || r == 0 || mbar_test_wait(&bars->o_free[g], (r - 1) & 1))) {
                tc_fence_after();
                const uint32_t v_addr = smem_u32(kv_tile + (2 * i + 1) * kVit2Tile);
                const uint32_t tO = tmem_base + static_cast<uint32_t>(g * 256 + 128), tP = tmem_base + static_cast<uint32_t>(g * 256 + 192);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {   // 16 keys per MMA: 8 packed TMEM columns of P
                  const uint64_t db = umma_desc_mn_sw128(v_addr + kk * 16 * 128, kVit2Tile, 1024);
                  umma_f16_ts(tO, tP + static_cast<uint32_t>(kk * 8), db, idesc_o, (i > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&bars->p_empty[g]);
                ++jp[g];
                progress = true;
              }
            }
            if (js[g] < nblk) {                           // S(js): needs K(js) and the softmax group to have read S(js - 1)
              const int j = js[g], sc = r * nblk + j;
              // (js[0] >= 2 implies tile A's first S was read; while js[0] == 1 the parity probe is unambiguous)
              const bool offset_ok = !(g == 1 && r == 0 && j == 0) || js[0] >= 2 || (js[0] == 1 && mbar_test_wait(&bars->s_empty[0], 0));
              if (offset_ok && mbar_test_wait(&bars->k_full[j], 0) && mbar_test_wait(&bars->s_empty[g], (sc & 1) ^ 1)) {
                tc_fence_after();
                const uint32_t q_addr = smem_u32(q_tile + g * kVit2Tile);
                const uint32_t k_addr = smem_u32(kv_tile + (2 * j) * kVit2Tile);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)            // 16 dims per MMA
                  umma_f16(tmem_base + static_cast<uint32_t>(g * 256), umma_desc_sw128(q_addr + kk * 32), umma_desc_sw128(k_addr + kk * 32),
                           idesc_s, kk > 0 ? 1u : 0u);
                umma_commit(&bars->s_full[g]);
                if (j == nblk - 1) umma_commit(&bars->q_empty[g]);   // every S MMA of the round has retired: Q buffer reusable
                ++js[g];
                progress = true;
              }
            }
          }
          if (progress) {
            idle = 0;
          } else if (++idle > (1u << (VLO_MBAR_BOUND_LOG2 + 4))) {
            printf("vlo: vit_attn_tc2 issuer stalled block(%d,%d,%d) round %d S %d/%d PV %d/%d\n", blockIdx.x, blockIdx.y, blockIdx.z, r,
                   js[0], js[1], jp[0], jp[1]);
            __trap();
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue: group g, thread = query row
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;                              // TMEM lane quadrant this warp may access
    const int rr = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + static_cast<uint32_t>(g * 256), tO = tS + 128, tP = tS + 192;
    const float c = scale_log2;
    const int my_rounds = (n_my - g + 1) >> 1;           // tiles t0 + g, t0 + g + 2, ...
    for (int r = 0; r < my_rounds; ++r) {
      const int tile = t0 + 2 * r + g;
      const int row = tile * 128 + rr;                   // token index inside the frame
      const bool valid = row < N;
      const bool warp_live = tile * 128 + q * 32 < N;
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nblk; ++j) {
        const int sc = r * nblk + j;
        mbar_wait(&bars->s_full[g], sc & 1);
        tc_fence_after();
        if (!warp_live) {   // rows past the frame: keep the barrier protocol going (their P / O rows are never stored)
          mbar_arrive(&bars->s_empty[g]);
          mbar_wait(&bars->p_empty[g], (sc & 1) ^ 1);
          mbar_arrive(&bars->p_full[g]);
          continue;
        }
        float sv[128];
        {
          uint32_t su[128];
#pragma unroll
          for (int c0 = 0; c0 < 128; c0 += 32)
            tmem_ld_x32(tS + lane_addr + static_cast<uint32_t>(c0), *reinterpret_cast<uint32_t(*)[32]>(&su[c0]));
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 128; ++i) sv[i] = __uint_as_float(su[i]);
        }
        tc_fence_before();
        mbar_arrive(&bars->s_empty[g]);
        const int key0 = j * 128;
        if (key0 + 128 > N) {   // last block: keys past the frame (next frame's rows / zero fill) are masked
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (key0 + i >= N) sv[i] = -INFINITY;
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          mx4[0] = fmaxf(mx4[0], sv[i]);
          mx4[1] = fmaxf(mx4[1], sv[i + 1]);
          mx4[2] = fmaxf(mx4[2], sv[i + 2]);
          mx4[3] = fmaxf(mx4[3], sv[i + 3]);
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        const float m_new = fmaxf(m_ref, mx);
        const bool grow = (m_ref == -INFINITY) ? (m_new != -INFINITY) : ((m_new - m_ref) * c > kTcRescaleLog2);
        const float m_use = grow ? m_new : m_ref;
        const float alpha = (grow && m_ref != -INFINITY) ? exp2f((m_ref - m_new) * c) : 1.f;
        m_ref = m_use;
        const float nmoff = (m_ref == -INFINITY) ? 0.f : -m_ref * c;
        // exponentials first (registers), so that only the TMEM stores sit behind the wait for PV(sc - 1)
        uint32_t w[64];
        float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float p0 = ex2_approx(fmaf(sv[2 * i], c, nmoff));
          const float p1 = ex2_approx(fmaf(sv[2 * i + 1], c, nmoff));
          ps4[i & 3] += p0 + p1;
          w[i] = pack_f16x2(p0, p1);
        }
        // single P buffer per group: PV(sc - 1) must have retired before P(sc) overwrites it; the same wait makes O safe
        // to rescale.  Followed phase by phase (one completion per PV): no parity aliasing.
        mbar_wait(&bars->p_empty[g], (sc & 1) ^ 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t v[32];
            tmem_ld_x32(tO + lane_addr + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_x32(tO + lane_addr + c0, v);
          }
          l_run *= alpha;
        }
        tmem_st_x32(tP + lane_addr, *reinterpret_cast<const uint32_t(*)[32]>(&w[0]));
        tmem_st_x32(tP + lane_addr + 32, *reinterpret_cast<const uint32_t(*)[32]>(&w[32]));
        l_run += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars->p_full[g]);
      }
      // ---- epilogue of the round: out = O / l (fp16); wait for the last PV of the round (completion #(r + 1) * nblk)
      mbar_wait(&bars->p_empty[g], ((r + 1) * nblk - 1) & 1);
      tc_fence_after();
      if (warp_live) {   // warp-uniform: tcgen05.ld is .sync.aligned
        uint32_t v0[32], v1[32];
        tmem_ld_x32(tO + lane_addr, v0);
        tmem_ld_x32(tO + lane_addr + 32, v1);
        tmem_ld_wait();
        if (valid) {
          const float inv = 1.f / l_run;
          uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(row_base + row) * C + head * 64);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_uint4(pack_f16x2(__uint_as_float(v0[8 * i]) * inv, __uint_as_float(v0[8 * i + 1]) * inv),
                                pack_f16x2(__uint_as_float(v0[8 * i + 2]) * inv, __uint_as_float(v0[8 * i + 3]) * inv),
                                pack_f16x2(__uint_as_float(v0[8 * i + 4]) * inv, __uint_as_float(v0[8 * i + 5]) * inv),
                                pack_f16x2(__uint_as_float(v0[8 * i + 6]) * inv, __uint_as_float(v0[8 * i + 7]) * inv));
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[4 + i] = make_uint4(pack_f16x2(__uint_as_float(v1[8 * i]) * inv, __uint_as_float(v1[8 * i + 1]) * inv),
                                    pack_f16x2(__uint_as_float(v1[8 * i + 2]) * inv, __uint_as_float(v1[8 * i + 3]) * inv),
                                    pack_f16x2(__uint_as_float(v1[8 * i + 4]) * inv, __uint_as_float(v1[8 * i + 5]) * inv),
                                    pack_f16x2(__uint_as_float(v1[8 * i + 6]) * inv, __uint_as_float(v1[8 * i + 7]) * inv));
        }
      }
      tc_fence_before();
      mbar_arrive(&bars->o_free[g]);   // the next round's first PV may overwrite O
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vlo
