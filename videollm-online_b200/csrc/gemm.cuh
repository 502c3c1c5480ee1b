// tcgen05 / TMA GEMM for sm_100a:   C[i, j] = sum_k A[i, k] * B[j, k]
//
// Both operands are K-major ("TN"): A is [rows_a, K], B is [rows_b, K], 16-bit
// (bf16 or fp16), fp32 accumulation in TMEM.  One CTA computes one 128 x BN
// tile (optionally one K-split of it):
//   warp 0     TMA producer   (cp.async.bulk.tensor, 128B-swizzled boxes)
//   warp 1     MMA issuer     (tcgen05.mma cta_group::1 kind::f16, one thread)
//   warps 2-5  epilogue       (tcgen05.ld 32x32b -> registers -> global)
// connected by an mbarrier full/empty ring plus one "accumulator ready" barrier.
//
// Two orientations share the kernel:
//   SWAP = false  A = activations [M, K], B = weights [N, K]; out[M, N]   (ViT: M = 576*B rows)
//   SWAP = true   A = weights [N, K],     B = activations [T, K]; out[T, N]
//                 "swap-AB": the weight matrix rides the 128-wide MMA-M dimension and
//                 the (few) tokens are the MMA-N dimension (16..256), so a decoder step
//                 with T = 11..96 tokens streams every weight byte exactly once at full
//                 tensor-core tile efficiency.  The kernel is then HBM-bound.
//
// Replaces, for the hot path, the cuBLAS/cuDNN GEMMs reached from
//   HF:models/llama/modeling_llama.py:262-264,288,182-184,487  (q/k/v/o, MLP, lm_head)
//   HF:models/siglip/modeling_siglip.py:175-179,285-287,309,323-327 (patch embed, attn, MLP)
//   models/live_llama/modeling_live_llama.py:18-22 (connector)
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace vlo {

enum GemmEpi : int {
  EPI_PARTIAL = 0,  // fp32 raw accumulators -> ws[split][out_row][out_col]   (split-K)
  EPI_STORE16 = 1,  // out16 = r16(act(r16(acc + bias)))
  EPI_RESID32 = 2,  // out32 += r16(acc + bias)                (fp32 residual stream, ViT)
  EPI_PATCH32 = 3,  // out32  = r16(acc + bias) + pos[out_row % pos_rows]     (ViT patch embed)
};
enum GemmAct : int { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_GELU_ERF_PY = 2 };
enum GemmFmt : int { FMT_F16 = 0, FMT_BF16 = 1 };

struct GemmArgs {
  int rows_a, rows_b, k;
  int kb_per_split;  // 64-wide k-blocks handled by one split (blockIdx.z)
  void* out;
  int ld_out;  // row stride of the logical output, elements
  const float* bias;  // per output column (feature) or nullptr
  const float* pos;   // EPI_PATCH32: [pos_rows, ld_out] fp32
  int pos_rows;
  int act;
  long long split_stride;  // EPI_PARTIAL: elements between split planes
  unsigned long long hint_a, hint_b;  // TMA L2 cache hints
};

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;  // one 128-byte swizzle atom of 16-bit elements
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = BN <= 64 ? 8 : (BN <= 128 ? 6 : 4);
  static constexpr int kBytesA = kGemmBM * kGemmBK * 2;
  static constexpr int kBytesB = BN * kGemmBK * 2;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  // 1024 B of slack so the tile ring can be aligned for SWIZZLE_128B.
  static constexpr int kSmemBytes = kStages * (kBytesA + kBytesB) + 1024 + 256 + 1024;  // + bias tile
};

template <int FMT>
__device__ __forceinline__ float r16(float x) {
  return FMT == FMT_BF16 ? bf16_round(x) : fp16_round(x);
}
template <int FMT>
__device__ __forceinline__ uint16_t to16(float x) {
  if (FMT == FMT_BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(x));
  return __half_as_ushort(__float2half_rn(x));
}

// Activations evaluated the way the reference evaluates them on 16-bit tensors.
template <int FMT>
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_GELU_TANH) {
    // F.gelu(x, approximate='tanh'), HF:activations.py ("gelu_pytorch_tanh"); fp32 math on
    // the 16-bit input, one rounding at the end.
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float inner = k0 * (v + k1 * v * v * v);
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(inner));  // MUFU.TANH: |err| ~ 2^-11, below fp16 output resolution
    return r16<FMT>(0.5f * v * (1.0f + th));
  }
  if (act == ACT_GELU_ERF_PY) {
    // GELUActivation(use_gelu_python=True): x * 0.5 * (1 + erf(x / sqrt(2))) evaluated as four
    // separate 16-bit tensor ops (HF:activations.py:78-86 as instantiated at
    // models/live_llama/modeling_live_llama.py:20) -> one rounding per op.
    float t1 = r16<FMT>(v * 0.5f);
    float t2 = r16<FMT>(v / 1.4142135623730951f);
    float t3 = r16<FMT>(erff(t2));
    float t4 = r16<FMT>(1.0f + t3);
    return r16<FMT>(t1 * t4);
  }
  return v;
}

template <int FMT, int BN, bool SWAP, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               const GemmArgs p) {
  using Cfg = GemmCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S * Cfg::kBytesA;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * (Cfg::kBytesA + Cfg::kBytesB));
  uint64_t* empty_bar = full_bar + S;
  uint64_t* accum_bar = empty_bar + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, n_tile = blockIdx.y, split = blockIdx.z;
  const int total_kb = p.k / kGemmBK;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(kb0 + p.kb_per_split, total_kb);
  const int nkb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();  // after the TMEM allocation
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], Cfg::kBytesA + Cfg::kBytesB);
        const int kc = (kb0 + i) * kGemmBK;
        tma_load_2d(smem_a + s * Cfg::kBytesA, &tm_a, &full_bar[s], kc, m_tile * kGemmBM, p.hint_a);
        tma_load_2d(smem_b + s * Cfg::kBytesB, &tm_b, &full_bar[s], kc, n_tile * BN, p.hint_b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------ MMA issuer
      constexpr uint32_t idesc = umma_idesc_f16(FMT, kGemmBM, BN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(smem_u32(smem_a + s * Cfg::kBytesA));
        const uint64_t db = umma_desc_sw128(smem_u32(smem_b + s * Cfg::kBytesB));
#pragma unroll
        for (int kk = 0; kk < kGemmBK / 16; ++kk) {
          // advance 16 elements (32 B) along K inside the swizzle atom: +2 in 16-B units
          umma_f16(tmem_base, da + 2 * kk, db + 2 * kk, idesc, (i > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(accum_bar);
    }
  } else {
    // -------------------------------------------------- epilogue warps 2..5
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int a_row = m_tile * kGemmBM + q * 32 + lane;
    // per-column bias of this tile -> shared memory while the mainloop runs (SWAP = false only)
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);
    if (!SWAP && EPI != EPI_PARTIAL) {
      const int et = threadIdx.x - 64;  // 0..127
      for (int j = et; j < BN; j += 128) {
        const int col = n_tile * BN + j;
        s_bias[j] = (p.bias != nullptr && col < p.rows_b) ? __ldg(p.bias + col) : 0.f;
      }
      asm volatile("bar.sync 1, 128;\n" ::: "memory");  // epilogue warps only
    }
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const bool a_ok = a_row < p.rows_a;
    float bias_a = 0.f;
    if (SWAP && EPI != EPI_PARTIAL && p.bias != nullptr && a_ok) bias_a = __ldg(p.bias + a_row);

#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0), v);
      tmem_ld_wait();
      const int b0 = n_tile * BN + c0;
      if (!a_ok || b0 >= p.rows_b) continue;
      if (SWAP) {
        // out[token = b, feature = a]; lanes hold consecutive features -> coalesced.
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int b = b0 + j;
          if (b >= p.rows_b) break;
          const float acc = __uint_as_float(v[j]);
          const size_t idx = static_cast<size_t>(b) * p.ld_out + a_row;
          if (EPI == EPI_PARTIAL) {
            reinterpret_cast<float*>(p.out)[static_cast<size_t>(split) * p.split_stride + idx] = acc;
          } else if (EPI == EPI_STORE16) {
            float x = r16<FMT>(acc + bias_a);
            x = apply_act<FMT>(x, p.act);
            reinterpret_cast<uint16_t*>(p.out)[idx] = to16<FMT>(x);
          } else if (EPI == EPI_RESID32) {
            float* o = reinterpret_cast<float*>(p.out) + idx;
            *o = *o + r16<FMT>(acc + bias_a);
          } else {
            const float pe = __ldg(p.pos + static_cast<size_t>(b % p.pos_rows) * p.ld_out + a_row);
            reinterpret_cast<float*>(p.out)[idx] = r16<FMT>(acc + bias_a) + pe;
          }
        }
      } else {
        // out[row = a, feature = b]; this thread owns 16 consecutive features of one row.
        const int nvalid = min(16, p.rows_b - b0);
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float acc = __uint_as_float(v[j]);
          if (EPI != EPI_PARTIAL) {
            acc = r16<FMT>(acc + s_bias[c0 + j]);
            if (EPI == EPI_STORE16) acc = apply_act<FMT>(acc, p.act);
          }
          x[j] = acc;
        }
        const size_t idx = static_cast<size_t>(a_row) * p.ld_out + b0;
        const bool vec_ok = (nvalid == 16) && ((p.ld_out & 7) == 0);
        if (EPI == EPI_STORE16) {
          uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + idx;
          if (vec_ok) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              w[j] = static_cast<uint32_t>(to16<FMT>(x[2 * j])) |
                     (static_cast<uint32_t>(to16<FMT>(x[2 * j + 1])) << 16);
            reinterpret_cast<uint4*>(o)[0] = make_uint4(w[0], w[1], w[2], w[3]);
            reinterpret_cast<uint4*>(o)[1] = make_uint4(w[4], w[5], w[6], w[7]);
          } else {
            for (int j = 0; j < nvalid; ++j) o[j] = to16<FMT>(x[j]);
          }
        } else {
          float* o = reinterpret_cast<float*>(p.out) + idx +
                     (EPI == EPI_PARTIAL ? static_cast<size_t>(split) * p.split_stride : 0);
          const float* pe = nullptr;
          if (EPI == EPI_PATCH32)
            pe = p.pos + static_cast<size_t>(a_row % p.pos_rows) * p.ld_out + b0;
          if (vec_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 t = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
              if (EPI == EPI_RESID32) {
                const float4 old = reinterpret_cast<float4*>(o + j)[0];
                t.x += old.x; t.y += old.y; t.z += old.z; t.w += old.w;
              } else if (EPI == EPI_PATCH32) {
                const float4 pp = __ldg(reinterpret_cast<const float4*>(pe + j));
                t.x += pp.x; t.y += pp.y; t.z += pp.z; t.w += pp.w;
              }
              reinterpret_cast<float4*>(o + j)[0] = t;
            }
          } else {
            for (int j = 0; j < nvalid; ++j) {
              float t = x[j];
              if (EPI == EPI_RESID32) t += o[j];
              if (EPI == EPI_PATCH32) t += __ldg(pe + j);
              o[j] = t;
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace vlo
