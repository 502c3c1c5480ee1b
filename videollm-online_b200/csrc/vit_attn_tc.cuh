// ViT self-attention on the 5th-gen tensor cores (replaces the mma.sync vit_attn_kernel of vit_kernels.cuh):
// non-causal softmax(Q K^T / sqrt(64)) V, fp16 operands / fp32 accumulation, head_dim 64, N = 576 tokens per frame
// (HF:models/siglip/modeling_siglip.py:275-312, the attention inside SiglipEncoderLayer).
//
//   grid = (ceil(N / 128) query tiles, heads, frames); block = 192:
//   warp 0      TMA producer: the Q tile [128 x 64] and ALL key blocks of the (frame, head): K_j, V_j [128 keys x 64 d]
//               (<= kVitTcMaxBlk blocks of 32 KB: the whole 576-token K / V of one head stays resident in shared memory,
//               every load is issued up front);
//   warp 1      MMA issuer:   S_j = Q K_j^T (tcgen05.mma M = N = 128, K = 64) into TMEM S[j & 1];  O += P_j V_j with the
//               A operand P_j read from TMEM (written there by the softmax threads) and V_j as an MN-major B operand;
//   warps 2-5   softmax:      one thread per query row (TMEM lane): tcgen05.ld S row -> online softmax with lazy O
//               rescaling -> fp16 P pairs straight into TMEM (tcgen05.st) -> epilogue O / l -> fp16 out.
//   TMEM columns: S0 0..127 | S1 128..255 | O 256..319 | P0 384..447 | P1 448..511.
// qkv: [B*N, 3C] fp16 (q | k | v column blocks, head h at column h*64 inside each); out: [B*N, C] fp16.
#pragma once
#include <cuda.h>
#include "tc_helpers.cuh"

namespace vlo {

constexpr int kVitTcBlk = 128;                       // keys per block and query rows per tile
constexpr int kVitTcMaxBlk = 5;                      // resident key blocks: N <= 640 (SigLIP-L/16-384: 576)
constexpr int kVitTcTile = kVitTcBlk * 64 * 2;       // [128 rows x 64 d] fp16, 128B-swizzled: 16 KB
constexpr int kVitTcSmemBytes = kVitTcTile * (1 + 2 * kVitTcMaxBlk) + 1024 + 256;
constexpr int kVitTcThreads = 192;

__host__ __device__ constexpr uint32_t umma_idesc_f16in(int m, int n, int b_mn_major) {   // fp16 x fp16 -> fp32
  return (1u << 4) | (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void __launch_bounds__(kVitTcThreads, 1)
vit_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, __half* out, int N, int C, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* q_tile = smem;
  uint8_t* kv_tile = smem + kVitTcTile;                    // block j: K at 2j, V at 2j + 1 (tiles of 16 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_tile + 2 * kVitTcMaxBlk * kVitTcTile);
  uint64_t* k_full = bars;                       // [kVitTcMaxBlk]
  uint64_t* v_full = k_full + kVitTcMaxBlk;      // [kVitTcMaxBlk]
  uint64_t* s_full = v_full + kVitTcMaxBlk;      // [2]
  uint64_t* s_empty = s_full + 2;                // [2]
  uint64_t* p_full = s_empty + 2;                // [2]
  uint64_t* p_empty = p_full + 2;                // [2]
  uint64_t* o_done = p_empty + 2;                // [1]
  uint64_t* q_ready = o_done + 1;                // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);

  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = (N + kVitTcBlk - 1) / kVitTcBlk;        // host guarantees nblk <= kVitTcMaxBlk
  const int row_base = b * N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    for (int i = 0; i < kVitTcMaxBlk; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_done, 1);
    mbar_init(q_ready, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();   // q | k | v come from the QKV GEMM right before
  const uint32_t tS = tmem_base, tO = tmem_base + 256, tP = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer: everything up front
      mbar_arrive_expect_tx(q_ready, kVitTcTile);
      tma_load_2d(q_tile, &tm_qkv, q_ready, head * 64, row_base + qt * kVitTcBlk, kEvictNormal);
      for (int j = 0; j < nblk; ++j) {
        mbar_arrive_expect_tx(&k_full[j], kVitTcTile);
        tma_load_2d(kv_tile + (2 * j) * kVitTcTile, &tm_qkv, &k_full[j], C + head * 64, row_base + j * kVitTcBlk, kEvictNormal);
      }
      for (int j = 0; j < nblk; ++j) {
        mbar_arrive_expect_tx(&v_full[j], kVitTcTile);
        tma_load_2d(kv_tile + (2 * j + 1) * kVitTcTile, &tm_qkv, &v_full[j], 2 * C + head * 64, row_base + j * kVitTcBlk, kEvictNormal);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // -------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc_s = umma_idesc_f16in(128, 128, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16in(128, 64, 1);
      const uint32_t q_addr = smem_u32(q_tile);
      mbar_wait(q_ready, 0);
      tc_fence_after();
      auto issue_pv = [&](int i) {
        const int pb = i & 1;
        mbar_wait(&v_full[i], 0);
        mbar_wait(&p_full[pb], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(kv_tile + (2 * i + 1) * kVitTcTile);
#pragma unroll
        for (int kk = 0; kk < kVitTcBlk / 16; ++kk) {  // 16 keys per MMA: 8 packed TMEM columns of P
          const uint64_t db = umma_desc_mn_sw128(v_addr + kk * 16 * 128, kVitTcTile, 1024);
          umma_f16_ts(tO, tP + static_cast<uint32_t>(pb * 64 + kk * 8), db, idesc_o, (i > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&p_empty[pb]);
      };
      for (int j = 0; j < nblk; ++j) {
        const int sb = j & 1;
        mbar_wait(&k_full[j], 0);
        mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(kv_tile + (2 * j) * kVitTcTile);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {  // 16 dims per MMA
          const uint64_t da = umma_desc_sw128(q_addr + kk * 32);
          const uint64_t db = umma_desc_sw128(k_addr + kk * 32);
          umma_f16(tS + sb * 128, da, db, idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[sb]);
        if (j >= 1) issue_pv(j - 1);
      }
      issue_pv(nblk - 1);
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue: one thread per query row
    const int q = warp & 3;
    const int r = q * 32 + lane;                       // TMEM lane == row of the query tile
    const int row = qt * kVitTcBlk + r;                // token index inside the frame
    const bool valid = row < N;
    const bool warp_live = qt * kVitTcBlk + q * 32 < N;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float c = scale_log2;
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      if (!warp_live) {   // rows past the frame: keep the barrier protocol going (their P / O rows are never stored)
        mbar_arrive(&s_empty[sb]);
        mbar_wait(&p_empty[sb], ((j >> 1) & 1) ^ 1);
        mbar_arrive(&p_full[sb]);
        continue;
      }
      float sv[128];
      {
        uint32_t su[128];
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32)
          tmem_ld_x32(tS + lane_addr + static_cast<uint32_t>(sb * 128 + c0), *reinterpret_cast<uint32_t(*)[32]>(&su[c0]));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 128; ++i) sv[i] = __uint_as_float(su[i]);
      }
      tc_fence_before();
      mbar_arrive(&s_empty[sb]);
      const int key0 = j * kVitTcBlk;
      if (key0 + kVitTcBlk > N) {   // last block: keys past the frame (next frame's rows / zero fill) are masked
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
      if (__any_sync(0xffffffffu, alpha != 1.f)) {
        mbar_wait(&p_empty[sb ^ 1], ((j - 1) >> 1) & 1);   // PV_{j-1} has landed (phase-exact, see attn_tc2.cuh): rescale O
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld_x32(tO + lane_addr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_x32(tO + lane_addr + c0, v);
        }
        tmem_st_wait();
        l_run *= alpha;
      }
      m_ref = m_use;
      const float nmoff = (m_ref == -INFINITY) ? 0.f : -m_ref * c;
      mbar_wait(&p_empty[sb], ((j >> 1) & 1) ^ 1);   // PV_{j-2} has consumed this P buffer
      tc_fence_after();
      float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 64) {
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float p0 = ex2_approx(fmaf(sv[c0 + 2 * i], c, nmoff));
          const float p1 = ex2_approx(fmaf(sv[c0 + 2 * i + 1], c, nmoff));
          ps4[i & 3] += p0 + p1;
          w[i] = pack_f16x2(p0, p1);
        }
        tmem_st_x32(tP + lane_addr + static_cast<uint32_t>(sb * 64 + c0 / 2), w);
      }
      l_run += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[sb]);
    }
    // ---- epilogue: out = O / l  (fp16)
    mbar_wait(&p_empty[(nblk - 1) & 1], ((nblk - 1) >> 1) & 1);   // the last PV, phase-exact (see attn_tc2.cuh)
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
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vlo
