// 2-CTA tcgen05 GEMM (cta_group::2) for the TENSOR-bound ViT trunk at batch >= 3 frames (576 * B token rows):
//   out[t, n] = act(r16(sum_k X[t, k] * W[n, k] + bias[n]))            fp16 operands, fp32 accumulation in TMEM
//
// Orientation: token rows ride MMA-M, features ride MMA-N (the decoder's swap-AB kernel gemm_ws.cuh does the opposite:
// there the problem is HBM-bound and T <= 128).  A CTA PAIR (thread-block cluster of 2, the two SMs of a TPC) owns one
// 256 x BN output tile: UMMA M = 256 (each CTA's TMEM holds its own 128 token rows x BN fp32 columns), and each CTA
// stages only HALF of the W tile (BN / 2 feature rows) - the tensor core reads both halves through the pair's shared
// memory.  Per 64-wide k-block a CTA pulls 16 KB of X + 16 KB of W (BN = 256) for 4.2 MFLOP: half the L2 -> SM bytes per
// flop of the 128 x 192 single-CTA tiles this replaces, which is what bounds a 1.5 PFLOP/s tensor core fed from L2.
//
//   warp 0      TMA producer (both CTAs): X box [128 x 64], W box [BN/2 x 64], 128B-swizzled, 5 (BN = 256) / 7 (BN = 128) stages;
//               cp.async.bulk.tensor...cta_group::2 signals the LEADER CTA's full barrier for both CTAs' bytes
//   warp 1      TMEM allocation (cta_group::2, both CTAs); in the leader CTA one thread issues tcgen05.mma.cta_group::2
//               and commits with .multicast::cluster to the empty / accumulator-full barriers of BOTH CTAs
//   warps 2-5   epilogue (both CTAs): tcgen05.ld 32 columns of the own token row -> bias / activation / rounding -> a
//               128B-swizzled 16 KB panel in shared memory (two, ping-pong) -> TMA store (fp16) or TMA reduce-add (fp32
//               residual stream); two accumulator buffers: the epilogue of tile i overlaps the MMAs of tile i + 1.
//               Persistent: pair p runs tiles p, p + n_pairs, ...
// PDL: launched with the programmatic attribute, but it never triggers its successor early (gemm.cu: VLO_GEMM2_PDL).
// Replaces for the ViT the cuBLAS GEMMs reached from HF:models/siglip/modeling_siglip.py:285-287 (q/k/v), 309 (out_proj),
// 323-327 (fc1 + gelu_pytorch_tanh, fc2).
#pragma once
#include <cuda.h>
#include "gemm.cuh"
#include "tc_helpers.cuh"

namespace vlo {

enum Gemm2Epi : int { G2_STORE16 = 0, G2_RESID32 = 1 };

struct Gemm2Args {
  int rows_x, rows_w, k;
  int m_tiles, n_tiles;   // 256-row token tiles x BN-wide feature tiles
  void* out;              // G2_STORE16: fp16 [rows_x][ld_out];  G2_RESID32: fp32 residual stream [rows_x][ld_out], += r16(y)
  int ld_out;
  const float* bias;      // [rows_w]
  int act;
  int pdl_trigger;        // 1: let the successor kernel start its prologue early (griddepcontrol.launch_dependents)
};

template <int BN, int STAGES>
struct Gemm2Cfg {
  static constexpr int kStages = STAGES;
  static constexpr int kBytesA = 128 * kGemmBK * 2;
  static constexpr int kBytesB = (BN / 2) * kGemmBK * 2;
  static constexpr int kTmemCols = 2 * BN;          // two accumulator buffers (256 or 512 columns)
  static constexpr int kStoreBytes = 128 * 128;     // one output panel: 128 token rows x 128 bytes, 128B-swizzled
  static constexpr int kSmemBytes = kStages * (kBytesA + kBytesB) + 2 * kStoreBytes + 1024 + 512;
};

// ------------------------------------------------------------------ cluster / cta_group::2 wrappers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
// (non-.aligned forms: the role branches leave warps 0 and 1 diverged when they get here)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {   // shared::cluster address of a peer's copy
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols) : "memory");
}
// TMA load whose completion bytes are credited to an mbarrier of the PAIR (here: the leader CTA's), data into this CTA
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const void* tmap, uint32_t mbar_cluster_addr, int32_t c0, int32_t c1,
                                                uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far arrive (once) on the barrier at this shared-memory offset in EVERY CTA of `mask`
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// ---- epilogue through shared memory + TMA store (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void g2_epi_sync() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }   // the 4 epilogue warps
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int BN, int STAGES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
             const __grid_constant__ CUtensorMap tm_out, const Gemm2Args p) {
  using Cfg = Gemm2Cfg<BN, STAGES>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S * Cfg::kBytesA;
  uint8_t* smem_out = smem + S * (Cfg::kBytesA + Cfg::kBytesB);      // two output panels (ping-pong)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_out + 2 * Cfg::kStoreBytes);   // used in the leader only
  uint64_t* empty_bar = full_bar + S;
  uint64_t* acc_full = empty_bar + S;    // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]  used in the leader only: 4 epilogue warps x 2 CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int kb = p.k / kGemmBK;
  const int tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_out);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_cg2<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();     // both CTAs' barriers are initialised and both halves of the TMEM allocation are done
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.pdl_trigger) pdl_trigger();   // after the TMEM allocation (see gemm_ws.cuh)
  pdl_wait();             // X is the predecessor's output (no-op when launched without the PDL attribute)

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer (both CTAs)
      int i = 0;
      for (int t = pair; t < tiles; t += n_pairs) {
        const int mt = t % p.m_tiles, nt = t / p.m_tiles;
        // a half tile that starts past the end of its matrix (token count mod 256 in 1..128: 5, 9, 13 ... frames) loads
        // from the last valid row instead of issuing a fully out-of-range box; its results are never stored (the TMA store
        // clips rows >= rows_x / columns >= rows_w)
        const int row_x = min(mt * 256 + static_cast<int>(rank) * 128, p.rows_x - 1);
        const int row_w = min(nt * BN + static_cast<int>(rank) * (BN / 2), p.rows_w - 1);
        for (int j = 0; j < kb; ++j, ++i) {
          const int s = i % S;
          const uint32_t ph = (i / S) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);    // own slot free (the leader's commit is multicast to both CTAs)
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * (Cfg::kBytesA + Cfg::kBytesB));
          const uint32_t full_leader = mapa_rank(smem_u32(&full_bar[s]), 0);
          tma_load_2d_cg2(smem_a + s * Cfg::kBytesA, &tm_x, full_leader, j * kGemmBK, row_x, kEvictNormal);
          tma_load_2d_cg2(smem_b + s * Cfg::kBytesB, &tm_w, full_leader, j * kGemmBK, row_w, kEvictLast);
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // ------------------------------------------------ MMA issuer (leader CTA)
      constexpr uint32_t idesc = umma_idesc_f16(FMT_F16, 256, BN);
      int i = 0, item = 0;
      for (int t = pair; t < tiles; t += n_pairs, ++item) {
        const int buf = item & 1;
        mbar_wait(&acc_empty[buf], ((item >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int j = 0; j < kb; ++j, ++i) {
          const int s = i % S;
          const uint32_t ph = (i / S) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(smem_a + s * Cfg::kBytesA));
          const uint64_t db = umma_desc_sw128(smem_u32(smem_b + s * Cfg::kBytesB));
#pragma unroll
          for (int kk = 0; kk < kGemmBK / 16; ++kk)
            umma_f16_cg2(tacc, da + 2 * kk, db + 2 * kk, idesc, (j > 0 || kk > 0) ? 1u : 0u);
          umma_commit_cg2(&empty_bar[s], 3);
        }
        umma_commit_cg2(&acc_full[buf], 3);
      }
    }
  } else {
    // -------------------------------------------------- epilogue warps 2..5 (both CTAs): thread = token row
    // TMEM -> registers -> bias / activation / rounding -> 128B-swizzled panel in shared memory -> ONE TMA store (or TMA
    // reduce-add into the fp32 residual stream) per 128-byte-wide panel: full-line writes issued by the copy engine instead
    // of 32 strided 16-byte stores per warp instruction (which made the epilogue, not the MMA, the critical path).
    constexpr int kPanelCols = (EPI == G2_STORE16) ? 64 : 32;       // 128 bytes per row
    const int q = warp & 3;
    const int r = q * 32 + lane;                                     // row inside this CTA's 128-row half tile
    const bool store_thread = (warp == 2 && lane == 0);
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    int item = 0, pbuf = 0;
    for (int t = pair; t < tiles; t += n_pairs, ++item) {
      const int mt = t % p.m_tiles, nt = t / p.m_tiles;
      const int buf = item & 1;
      mbar_wait(&acc_full[buf], (item >> 1) & 1);
      tc_fence_after();
      const int row0 = mt * 256 + static_cast<int>(rank) * 128;
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += kPanelCols, pbuf ^= 1) {
        const uint32_t sb = smem_u32(smem_out + pbuf * Cfg::kStoreBytes) + static_cast<uint32_t>(r) * 128u;
        if (store_thread) bulk_wait_group_read<1>();   // the store issued from this buffer two panels ago has read it
        g2_epi_sync();
#pragma unroll
        for (int hh = 0; hh < kPanelCols / 32; ++hh) {
          uint32_t v[32];
          tmem_ld_x32(tacc + static_cast<uint32_t>(c0 + 32 * hh), v);
          tmem_ld_wait();
          if (c0 + 32 * hh + 32 >= BN) {   // last chunk is in registers: hand the accumulator buffer back to the leader's MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&acc_empty[buf]), 0));
          }
          const int n0 = nt * BN + c0 + 32 * hh;
          float x[32];
          if (n0 < p.rows_w) {              // rows_w is a multiple of 32: a chunk is whole or absent
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + i);
              x[4 * i] = fp16_round(__uint_as_float(v[4 * i]) + b.x);
              x[4 * i + 1] = fp16_round(__uint_as_float(v[4 * i + 1]) + b.y);
              x[4 * i + 2] = fp16_round(__uint_as_float(v[4 * i + 2]) + b.z);
              x[4 * i + 3] = fp16_round(__uint_as_float(v[4 * i + 3]) + b.w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = 0.f;
          }
          if (EPI == G2_STORE16) {
            if (p.act != ACT_NONE) {
#pragma unroll
              for (int i = 0; i < 32; ++i) x[i] = apply_act<FMT_F16>(x[i], p.act);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {   // 16-byte chunk (hh * 4 + i) of the row, XOR-swizzled with the row index
              __half2 h0 = __floats2half2_rn(x[8 * i], x[8 * i + 1]), h1 = __floats2half2_rn(x[8 * i + 2], x[8 * i + 3]);
              __half2 h2 = __floats2half2_rn(x[8 * i + 4], x[8 * i + 5]), h3 = __floats2half2_rn(x[8 * i + 6], x[8 * i + 7]);
              st_shared_v4(sb + ((static_cast<uint32_t>(hh * 4 + i) ^ sw) << 4), *reinterpret_cast<uint32_t*>(&h0),
                           *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
            }
          } else {   // G2_RESID32: fp32 panel, added into the residual stream by the copy engine (HF:...siglip.py:353, 360)
#pragma unroll
            for (int i = 0; i < 8; ++i)
              st_shared_v4(sb + ((static_cast<uint32_t>(i) ^ sw) << 4), __float_as_uint(x[4 * i]), __float_as_uint(x[4 * i + 1]),
                           __float_as_uint(x[4 * i + 2]), __float_as_uint(x[4 * i + 3]));
          }
        }
        fence_proxy_async();      // generic-proxy writes -> visible to the async proxy (TMA)
        g2_epi_sync();
        if (store_thread) {
          if (row0 < p.rows_x && nt * BN + c0 < p.rows_w) {   // (a panel entirely outside the matrix is not issued at all)
            if (EPI == G2_STORE16) tma_store_2d(&tm_out, smem_out + pbuf * Cfg::kStoreBytes, nt * BN + c0, row0);
            else tma_reduce_add_2d(&tm_out, smem_out + pbuf * Cfg::kStoreBytes, nt * BN + c0, row0);
          }
          bulk_commit_group();
        }
      }
    }
    if (store_thread) bulk_wait_group_all();
    tc_fence_before();
  }
  // teardown: nobody may free the pair's TMEM (or exit with a peer still multicasting into its barriers) early
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace vlo
