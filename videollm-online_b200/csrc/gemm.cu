// Host launcher for the tcgen05 GEMM: TMA descriptor construction (driver entry point
// resolved at run time, so the library loads on a box without libcuda), tile-shape
// selection and kernel dispatch.
#include "gemm.cuh"
#include "gemm_ws.cuh"
#include "gemm_wsf.cuh"
#include "gemm2.cuh"

#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.h"

namespace vlo {

namespace {

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                              CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  int rows, k, box_rows, fmt;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && k == o.k && box_rows == o.box_rows && fmt == o.fmt;
  }
};
struct TmapHash {
  size_t operator()(const TmapKey& x) const {
    size_t h = reinterpret_cast<size_t>(x.ptr);
    h ^= (static_cast<size_t>(x.rows) * 0x9E3779B97F4A7C15ull) + (h << 6) + (h >> 2);
    h ^= (static_cast<size_t>(x.k) * 0xC2B2AE3D27D4EB4Full) + (h << 6) + (h >> 2);
    h ^= static_cast<size_t>(x.box_rows * 2 + x.fmt) + (h << 6) + (h >> 2);
    return h;
  }
};

std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapHash> g_tmaps;

// [rows, k] row-major 16-bit matrix, box = 64 (k) x box_rows, 128-byte swizzle.
int get_tmap(const void* ptr, int rows, int k, int box_rows, int fmt, CUtensorMap* out) {
  TmapKey key{ptr, rows, k, box_rows, fmt};
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeFn enc = get_encode();
  if (enc == nullptr) return fail("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail("TMA operand not 16-byte aligned");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(k) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kGemmBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, fmt == FMT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed, CUresult " + std::to_string(r));
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    g_tmaps.emplace(key, m);
  }
  *out = m;
  return 0;
}

template <int FMT, int BN, bool SWAP, int EPI>
int launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, dim3 grid,
               cudaStream_t stream) {
  auto kern = gemm_tn_kernel<FMT, BN, SWAP, EPI>;
  if (ensure_max_smem(reinterpret_cast<const void*>(kern), GemmCfg<BN>::kSmemBytes)) return -1;
  if (prof_on()) {
    const double esz = (EPI == EPI_STORE16) ? 2.0 : 4.0;
    const double bytes = 2.0 * args.k * (static_cast<double>(args.rows_a) + args.rows_b) +
                         esz * static_cast<double>(args.rows_a) * args.rows_b * (EPI == EPI_PARTIAL ? grid.z : 1);
    prof_begin(SWAP ? PROF_GEMM_STREAM : PROF_GEMM_VIT, stream, bytes);
  }
  VLO_CUDA(launch_pdl(kern, grid, dim3(kGemmThreads), GemmCfg<BN>::kSmemBytes, stream, ta, tb, args));
  prof_end(stream);
  count_launch();
  return 0;
}

}  // namespace

// Query tensor [n_tok][n_heads][128] bf16 as a 3-D map; one box = (64 dims, G heads of one kv head, 128/G tokens)
// lands in shared memory as 128 rows of 128 B ordered row = token * G + head: the K-major SW128 A tile of the
// attention kernel.  Tokens past n_tok are zero-filled by the TMA unit.  box_rows (32 / 64 / 128, a multiple of G)
// selects a smaller box of box_rows / G tokens: the v3 kernel stacks 128 / box_rows copies of it (key slicing).
int tmap_q3d_sw128(const void* ptr, int n_tok, int n_heads, int G, CUtensorMap* out, int box_rows) {
  if (box_rows <= 0) box_rows = 128;
  TmapKey key{ptr, n_tok, n_heads, -G, 3 + 16 * box_rows};
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeFn enc = get_encode();
  if (enc == nullptr) return fail("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail("TMA operand not 16-byte aligned");
  if (G < 1 || 128 % G != 0 || n_heads % G != 0) return fail("tmap_q3d: GQA group must divide 128 and n_heads");
  if (box_rows % G != 0 || box_rows > 128) return fail("tmap_q3d: box rows must be a multiple of the GQA group, <= 128");
  cuuint64_t dims[3] = {128, static_cast<cuuint64_t>(n_heads), static_cast<cuuint64_t>(n_tok)};
  cuuint64_t strides[2] = {256, static_cast<cuuint64_t>(n_heads) * 256};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(G), static_cast<cuuint32_t>(box_rows / G)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (Q, 3-D) failed, CUresult " + std::to_string(r));
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    g_tmaps.emplace(key, m);
  }
  *out = m;
  return 0;
}

int tmap_2d_sw128(const void* ptr, int rows, int k, int box_rows, int fmt, CUtensorMap* out) {
  return get_tmap(ptr, rows, k, box_rows, fmt, out);
}

// Largest split count <= want for which every split owns at least one 64-wide k-block.
int gemm_fix_splits(int k, int want) {
  const int total_kb = k / kGemmBK;
  int s = want < 1 ? 1 : (want > total_kb ? total_kb : want);
  for (; s > 1; --s) {
    const int per = (total_kb + s - 1) / s;
    if ((total_kb + per - 1) / per == s) break;
  }
  return s;
}

int gemm_launch(const GemmCall& c, cudaStream_t stream) {
  VLO_CHECK(c.k > 0 && c.k % kGemmBK == 0, "K must be a positive multiple of 64");
  VLO_CHECK(c.rows_a > 0 && c.rows_b > 0, "empty GEMM operand");
  VLO_CHECK(c.splits >= 1, "splits >= 1");
  VLO_CHECK(c.splits == 1 || c.epi == EPI_PARTIAL, "split-K requires the fp32 partial epilogue");
  const int total_kb = c.k / kGemmBK;
  const int splits = c.splits > total_kb ? total_kb : c.splits;
  const int kb_per = (total_kb + splits - 1) / splits;
  const int eff_splits = (total_kb + kb_per - 1) / kb_per;  // every split gets >= 1 k-block
  VLO_CHECK(eff_splits == c.splits, "split count leaves an empty split; use gemm_fix_splits()");

  int bn = c.bn;
  if (bn == 0) {
    if (c.swap) bn = c.rows_b <= 16 ? 16 : (c.rows_b <= 32 ? 32 : (c.rows_b <= 64 ? 64 : 128));
    else bn = 128;
  }
  GemmArgs a{};
  a.rows_a = c.rows_a;
  a.rows_b = c.rows_b;
  a.k = c.k;
  a.kb_per_split = kb_per;
  a.out = c.out;
  a.ld_out = c.ld_out;
  a.bias = c.bias;
  a.pos = c.pos;
  a.pos_rows = c.pos_rows > 0 ? c.pos_rows : 1;
  a.act = c.act;
  a.split_stride = c.split_stride;
  a.hint_a = c.stream_weights ? kEvictFirst : kEvictNormal;
  a.hint_b = c.stream_weights ? kEvictLast : kEvictNormal;

  CUtensorMap ta, tb;
  if (get_tmap(c.a, c.rows_a, c.k, kGemmBM, c.fmt, &ta) != 0) return -1;
  if (get_tmap(c.b, c.rows_b, c.k, bn, c.fmt, &tb) != 0) return -1;
  dim3 grid((c.rows_a + kGemmBM - 1) / kGemmBM, (c.rows_b + bn - 1) / bn, eff_splits);

#define VLO_GEMM_CASE(F, N, SW, E)                                                        \
  if (c.fmt == F && bn == N && (c.swap != 0) == SW && c.epi == E)                         \
    return launch_one<F, N, SW, E>(ta, tb, a, grid, stream);

  // decoder / connector: bf16, swap-AB
  VLO_GEMM_CASE(FMT_BF16, 16, true, EPI_PARTIAL)
  VLO_GEMM_CASE(FMT_BF16, 32, true, EPI_PARTIAL)
  VLO_GEMM_CASE(FMT_BF16, 64, true, EPI_PARTIAL)
  VLO_GEMM_CASE(FMT_BF16, 128, true, EPI_PARTIAL)
  VLO_GEMM_CASE(FMT_BF16, 16, true, EPI_STORE16)
  VLO_GEMM_CASE(FMT_BF16, 32, true, EPI_STORE16)
  VLO_GEMM_CASE(FMT_BF16, 64, true, EPI_STORE16)
  VLO_GEMM_CASE(FMT_BF16, 128, true, EPI_STORE16)
  // ViT MAP head (few rows): fp16, swap-AB
  VLO_GEMM_CASE(FMT_F16, 16, true, EPI_STORE16)
  VLO_GEMM_CASE(FMT_F16, 32, true, EPI_STORE16)
  VLO_GEMM_CASE(FMT_F16, 64, true, EPI_STORE16)
  VLO_GEMM_CASE(FMT_F16, 128, true, EPI_STORE16)
  // ViT trunk: fp16, activations on MMA-M
  VLO_GEMM_CASE(FMT_F16, 64, false, EPI_STORE16)
  VLO_GEMM_CASE(FMT_F16, 128, false, EPI_STORE16)
  VLO_GEMM_CASE(FMT_F16, 64, false, EPI_RESID32)
  VLO_GEMM_CASE(FMT_F16, 128, false, EPI_RESID32)
  VLO_GEMM_CASE(FMT_F16, 64, false, EPI_PARTIAL)
  VLO_GEMM_CASE(FMT_F16, 128, false, EPI_PARTIAL)
  VLO_GEMM_CASE(FMT_F16, 64, false, EPI_PATCH32)
  VLO_GEMM_CASE(FMT_F16, 128, false, EPI_PATCH32)
#undef VLO_GEMM_CASE
  return fail("gemm_launch: no kernel instance for fmt=" + std::to_string(c.fmt) + " bn=" +
              std::to_string(bn) + " swap=" + std::to_string(c.swap) + " epi=" + std::to_string(c.epi));
}

// ------------------------------------------------------------------------------------------------
// persistent weight-streaming GEMM (gemm_ws.cuh)
namespace {
int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int FMT, int BN, int STAGES = ws_default_stages(BN)>
int launch_ws(const CUtensorMap& tw, const CUtensorMap& tx, const GemmWsArgs& a, cudaStream_t stream) {
  auto kern = gemm_ws_kernel<FMT, BN, STAGES>;
  if (ensure_max_smem(reinterpret_cast<const void*>(kern), GemmWsCfg<BN, STAGES>::kSmemBytes)) return -1;
  if (prof_on()) {
    const double out_b = a.mode == 0 ? 4.0 * a.rows_x * a.rows_w : 2.0 * a.rows_x * a.rows_w;
    prof_begin(FMT == FMT_BF16 ? PROF_GEMM_STREAM : PROF_GEMM_VIT, stream,
               2.0 * a.k * (static_cast<double>(a.rows_w) + a.rows_x) + out_b);
  }
  VLO_CUDA(launch_pdl(kern, dim3(a.sk.G), dim3(kGemmThreads), GemmWsCfg<BN, STAGES>::kSmemBytes, stream, tw, tx, a));
  prof_end(stream);
  count_launch();
  return 0;
}
}  // namespace

int gemm_ws_plan(int rows_w, int k, int mode, int n_ctas, SkInfo* sk, int* max_planes, int x_tiles) {
  const int tiles = ((rows_w + kGemmBM - 1) / kGemmBM) * (x_tiles > 0 ? x_tiles : 1);
  const int kb = k / kGemmBK;
  sk->kb = kb;
  sk->U = static_cast<long long>(tiles) * kb;
  const long long cap = mode == 0 ? sk->U : tiles;
  int G = n_ctas > 0 ? n_ctas : num_sms();
  if (G > cap) G = static_cast<int>(cap);
  sk->G = G;
  int mp = 1;
  if (mode == 0)
    for (int t = 0; t < tiles; ++t) mp = std::max(mp, sk_planes(t, *sk));
  if (max_planes) *max_planes = mp;
  return 0;
}

int gemm_ws_launch(const GemmWsCall& c, cudaStream_t stream) {
  VLO_CHECK(c.k > 0 && c.k % kGemmBK == 0, "K must be a positive multiple of 64");
  VLO_CHECK(c.rows_w > 0 && c.rows_x > 0, "empty GEMM operand");
  // token-tile width: the smallest instance covering rows_x (a narrower X tile leaves more of the ring to the weight stream:
  // 88 tokens = 8 streams x 11 take the 96-wide instance, 7 x 28 KB stages instead of 6 x 32 KB)
  const int bn = c.bn > 0 ? c.bn
                          : (c.rows_x <= 16 ? 16 : (c.rows_x <= 32 ? 32 : (c.rows_x <= 64 ? 64 : ((c.rows_x <= 96 && c.fmt == FMT_BF16) ? 96 : 128))));
  const int x_tiles = (c.rows_x + bn - 1) / bn;
  VLO_CHECK(c.sk.U == static_cast<long long>((c.rows_w + kGemmBM - 1) / kGemmBM) * x_tiles * (c.k / kGemmBK),
            "gemm_ws: plan does not match the call (x_tiles / bn)");
  GemmWsArgs a{};
  a.rows_w = c.rows_w;
  a.rows_x = c.rows_x;
  a.k = c.k;
  a.tiles = ((c.rows_w + kGemmBM - 1) / kGemmBM) * x_tiles;
  a.x_tiles = x_tiles;
  a.hint_w = c.weights_hot ? kEvictNormal : kEvictFirst;
  a.sk = c.sk;
  a.mode = c.mode;
  a.out = c.out;
  a.ld_out = c.ld_out;
  a.plane_stride = c.plane_stride;
  a.bias = c.bias;
  a.act = c.act;
  CUtensorMap tw, tx;
  if (get_tmap(c.w, c.rows_w, c.k, kGemmBM, c.fmt, &tw) != 0) return -1;
  if (get_tmap(c.x, c.rows_x, c.k, bn, c.fmt, &tx) != 0) return -1;
#define VLO_WS_CASE(F, N) \
  if (c.fmt == F && bn == N) return launch_ws<F, N>(tw, tx, a, stream);
  if (c.fmt == FMT_BF16 && bn == 16) {  // ring depth of the single-stream decoder GEMM: A/B switch (VLO_WS_STAGES=6|8|11)
    static int st = 0;
    if (st == 0) {
      const char* e = getenv("VLO_WS_STAGES");
      st = e ? atoi(e) : 6;
    }
    // 6 x 18 KB (measured best in the two-stream pipelined step: leaves room for a co-resident ViT CTA)
    if (st == 4) return launch_ws<FMT_BF16, 16, 4>(tw, tx, a, stream);
    if (st == 5) return launch_ws<FMT_BF16, 16, 5>(tw, tx, a, stream);
    if (st == 6) return launch_ws<FMT_BF16, 16, 6>(tw, tx, a, stream);
    if (st == 11) return launch_ws<FMT_BF16, 16, 11>(tw, tx, a, stream);
  }
  if (c.fmt == FMT_F16 && c.small_smem && bn == 96) return launch_ws<FMT_F16, 96, 3>(tw, tx, a, stream);    // 84 KB ring
  if (c.fmt == FMT_F16 && bn == 64 && c.small_smem) {
    static int vst = 0;
    if (vst == 0) {
      const char* e = getenv("VLO_VIT_STAGES");
      vst = e ? atoi(e) : 3;
    }
    if (vst == 4) return launch_ws<FMT_F16, 64, 4>(tw, tx, a, stream);
    return launch_ws<FMT_F16, 64, 3>(tw, tx, a, stream);
  }
  VLO_WS_CASE(FMT_BF16, 16)
  VLO_WS_CASE(FMT_BF16, 32)
  VLO_WS_CASE(FMT_BF16, 64)
  VLO_WS_CASE(FMT_BF16, 96)
  VLO_WS_CASE(FMT_BF16, 128)
  VLO_WS_CASE(FMT_F16, 16)
  VLO_WS_CASE(FMT_F16, 32)
  VLO_WS_CASE(FMT_F16, 64)
  VLO_WS_CASE(FMT_F16, 128)
  VLO_WS_CASE(FMT_F16, 96)
  VLO_WS_CASE(FMT_F16, 192)
#undef VLO_WS_CASE
  return fail("gemm_ws_launch: no kernel instance");
}

// ------------------------------------------------------------------------------------------------
// stream-K GEMM with the fused finisher epilogue (gemm_wsf.cuh)
namespace {
template <int BN, int STAGES, int EPI>
int launch_wsf(const CUtensorMap& tw, const CUtensorMap& tx, const GemmWsfArgs& a, cudaStream_t stream) {
  auto kern = gemm_wsf_kernel<BN, STAGES, EPI>;
  using Cfg = GemmWsfCfg<BN, STAGES>;
  if (ensure_max_smem(reinterpret_cast<const void*>(kern), Cfg::kSmemBytes)) return -1;
  if (prof_on())
    prof_begin(PROF_GEMM_STREAM, stream, 2.0 * a.k * (static_cast<double>(a.rows_w) + a.rows_x) + 2.0 * a.rows_x * a.rows_w);
  VLO_CUDA(launch_pdl(kern, dim3(a.sk.G), dim3(kGemmThreads), Cfg::kSmemBytes, stream, tw, tx, a));
  prof_end(stream);
  count_launch();
  return 0;
}
template <int BN, int STAGES>
int launch_wsf_epi(int epi, const CUtensorMap& tw, const CUtensorMap& tx, const GemmWsfArgs& a, cudaStream_t stream) {
  if (epi == WSF_RESID) return launch_wsf<BN, STAGES, WSF_RESID>(tw, tx, a, stream);
  if (epi == WSF_QKV) return launch_wsf<BN, STAGES, WSF_QKV>(tw, tx, a, stream);
  if (epi == WSF_SWIGLU) return launch_wsf<BN, STAGES, WSF_SWIGLU>(tw, tx, a, stream);
  return fail("gemm_wsf_launch: unknown epilogue");
}
}  // namespace

int gemm_wsf_launch(const GemmWsfArgs& a, const void* w, const void* x, int epi, cudaStream_t stream) {
  VLO_CHECK(a.k > 0 && a.k % kGemmBK == 0, "K must be a positive multiple of 64");
  VLO_CHECK(a.rows_w > 0 && a.rows_x > 0 && a.rows_x <= 128, "gemm_wsf: 1..128 token rows");
  VLO_CHECK(epi == WSF_RESID || a.rows_w % kGemmBM == 0, "gemm_wsf: QKV / SwiGLU epilogues need whole 128-row tiles");
  const int bn = a.rows_x <= 16 ? 16 : (a.rows_x <= 32 ? 32 : (a.rows_x <= 64 ? 64 : (a.rows_x <= 96 ? 96 : 128)));
  VLO_CHECK(a.sk.U == static_cast<long long>((a.rows_w + kGemmBM - 1) / kGemmBM) * (a.k / kGemmBK), "gemm_wsf: plan mismatch");
  CUtensorMap tw, tx;
  if (get_tmap(w, a.rows_w, a.k, kGemmBM, FMT_BF16, &tw) != 0) return -1;
  if (get_tmap(x, a.rows_x, a.k, bn, FMT_BF16, &tx) != 0) return -1;
  if (bn == 16) {
    // ring depth (VLO_WSF_STAGES=4|5|6|8|10; default 8 = 144 KB + the 8.7 KB exchange tile).  Same-box A/Bs: the decoder step
    // alone does not care (3.97-3.99 ms for 6 / 8 / 10 stages; 5 and 4 stages, which let the next GEMM's CTA become resident
    // beside this one, are slower: 4.02 / 4.07); next to the co-resident ViT 8 stages measure 211 vs 208 frames/s - the ViT's
    // CTAs then cannot share an SM with the (only) fused GEMM, gate|up, and stop competing with its stream.
    static int st = 0;
    if (st == 0) {
      const char* e = getenv("VLO_WSF_STAGES");
      st = e ? atoi(e) : 8;
    }
    if (st == 4) return launch_wsf_epi<16, 4>(epi, tw, tx, a, stream);
    if (st == 5) return launch_wsf_epi<16, 5>(epi, tw, tx, a, stream);
    if (st == 8) return launch_wsf_epi<16, 8>(epi, tw, tx, a, stream);
    if (st == 10) return launch_wsf_epi<16, 10>(epi, tw, tx, a, stream);
    if (st == 6) return launch_wsf_epi<16, 6>(epi, tw, tx, a, stream);
    return launch_wsf_epi<16, 8>(epi, tw, tx, a, stream);
  }
  if (bn == 32) return launch_wsf_epi<32, ws_default_stages(32)>(epi, tw, tx, a, stream);
  if (bn == 64) return launch_wsf_epi<64, ws_default_stages(64)>(epi, tw, tx, a, stream);
  if (bn == 96) return launch_wsf_epi<96, 6>(epi, tw, tx, a, stream);
  return launch_wsf_epi<128, ws_default_stages(128)>(epi, tw, tx, a, stream);
}

// ------------------------------------------------------------------------------------------------
// 2-CTA tensor-bound GEMM (gemm2.cuh)
namespace {
// output map for the TMA-store epilogue: [rows, cols] row-major, 2- or 4-byte elements, box = 128 rows x 128 bytes, 128B swizzle
int tmap_out_sw128(const void* ptr, int rows, int cols, int elem_bytes, CUtensorMap* out) {
  TmapKey key{ptr, rows, cols, -128, 100 + elem_bytes};
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeFn enc = get_encode();
  if (enc == nullptr) return fail("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail("TMA operand not 16-byte aligned");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elem_bytes), 128};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (output) failed, CUresult " + std::to_string(r));
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    g_tmaps.emplace(key, m);
  }
  *out = m;
  return 0;
}

template <int BN, int STAGES, int EPI>
int launch_gemm2(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& to, const Gemm2Args& a, cudaStream_t stream) {
  auto kern = gemm2_kernel<BN, STAGES, EPI>;
  using Cfg = Gemm2Cfg<BN, STAGES>;
  if (ensure_max_smem(reinterpret_cast<const void*>(kern), Cfg::kSmemBytes)) return -1;
  const int tiles = a.m_tiles * a.n_tiles;
  const int n_pairs = std::min(num_sms() / 2, tiles);
  if (prof_on())
    prof_begin(PROF_GEMM_VIT, stream, 2.0 * a.k * (static_cast<double>(a.rows_w) + a.rows_x) + (EPI == G2_STORE16 ? 2.0 : 8.0) * a.rows_x * a.rows_w);
  // VLO_GEMM2_PDL bit 0: launch with the programmatic-serialization attribute (prologue - barrier init, TMEM allocation of
  // the pair, cluster sync - overlaps the predecessor's tail); bit 1: trigger the successor early.  Default 1.  Measured
  // (tools/gpu_r2_call11.sh, ViT alone at batch 8): 0 -> 5.07 ms, 2 -> 5.10, 1 -> 4.83; 3 DEAD-LOCKS next to the
  // non-cluster kernels of the trunk (three kernels deep: a cluster kernel resident early AND its successor resident early;
  // the gemm2 <-> gemm2 chain alone runs), so a cluster kernel here never triggers early.
  static int pdl_mode = -1;
  if (pdl_mode < 0) {
    const char* e = getenv("VLO_GEMM2_PDL");
    pdl_mode = e ? atoi(e) : 1;
  }
  Gemm2Args a2 = a;
  a2.pdl_trigger = (pdl_mode & 2) ? 1 : 0;
  if (pdl_mode & 1) {
    VLO_CUDA(launch_pdl(kern, dim3(2 * n_pairs), dim3(kGemmThreads), Cfg::kSmemBytes, stream, tx, tw, to, a2));
  } else {
    kern<<<dim3(2 * n_pairs), dim3(kGemmThreads), Cfg::kSmemBytes, stream>>>(tx, tw, to, a2);
    VLO_LAUNCH_CHECK();
  }
  prof_end(stream);
  count_launch();
  return 0;
}
}  // namespace

int gemm2_launch(const Gemm2Call& c, cudaStream_t stream) {
  VLO_CHECK(c.k > 0 && c.k % kGemmBK == 0, "gemm2: K must be a positive multiple of 64");
  VLO_CHECK(c.rows_x > 0 && c.rows_w > 0 && c.rows_w % 64 == 0, "gemm2: rows_w must be a positive multiple of 64");
  VLO_CHECK(c.bn == 256 || c.bn == 128, "gemm2: bn is 256 or 128");
  VLO_CHECK(c.bias != nullptr && (reinterpret_cast<uintptr_t>(c.bias) & 15) == 0, "gemm2: bias required, 16-byte aligned");
  VLO_CHECK(c.ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(c.out) & 15) == 0, "gemm2: output rows must be 16-byte aligned");
  Gemm2Args a{};
  a.rows_x = c.rows_x;
  a.rows_w = c.rows_w;
  a.k = c.k;
  a.m_tiles = (c.rows_x + 255) / 256;
  a.n_tiles = (c.rows_w + c.bn - 1) / c.bn;
  a.out = c.out;
  a.ld_out = c.ld_out;
  a.bias = c.bias;
  a.act = c.act;
  CUtensorMap tx, tw;
  if (get_tmap(c.x, c.rows_x, c.k, 128, FMT_F16, &tx) != 0) return -1;
  if (get_tmap(c.w, c.rows_w, c.k, c.bn / 2, FMT_F16, &tw) != 0) return -1;
  CUtensorMap to;
  VLO_CHECK(c.ld_out == c.rows_w, "gemm2: the output is a dense [rows_x, rows_w] matrix");
  if (tmap_out_sw128(c.out, c.rows_x, c.rows_w, c.epi == G2_STORE16 ? 2 : 4, &to) != 0) return -1;
  // ring: 5 x 32 KB (bn 256) / 7 x 24 KB (bn 128) + two 16 KB output panels
  if (c.bn == 256 && c.epi == G2_STORE16) return launch_gemm2<256, 5, G2_STORE16>(tx, tw, to, a, stream);
  if (c.bn == 128 && c.epi == G2_STORE16) return launch_gemm2<128, 7, G2_STORE16>(tx, tw, to, a, stream);
  if (c.bn == 256 && c.epi == G2_RESID32) return launch_gemm2<256, 5, G2_RESID32>(tx, tw, to, a, stream);
  if (c.bn == 128 && c.epi == G2_RESID32) return launch_gemm2<128, 7, G2_RESID32>(tx, tw, to, a, stream);
  return fail("gemm2_launch: no kernel instance");
}

}  // namespace vlo
