// Host-side declarations shared by the translation units of libvlo_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "streamk.h"

namespace vlo {

// ---- error plumbing: every C-ABI entry returns 0 / negative and leaves a message here.
void set_error(const std::string& msg);
const char* last_error();
int fail(const std::string& msg);  // set_error + return -1

#define VLO_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      return ::vlo::fail(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" +    \
                         __FILE__ + ":" + std::to_string(__LINE__));                    \
  } while (0)

#define VLO_CHECK(cond, msg)                                                            \
  do {                                                                                  \
    if (!(cond)) return ::vlo::fail(std::string("check failed: ") + #cond + ": " + msg); \
  } while (0)

#define VLO_LAUNCH_CHECK()                                                              \
  do {                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess)                                                              \
      return ::vlo::fail(std::string("kernel launch: ") + cudaGetErrorString(_e) + " @" + \
                         __FILE__ + ":" + std::to_string(__LINE__));                    \
  } while (0)

// Kernel launch with the PDL attribute (see ptx.cuh: pdl_wait / pdl_trigger).  Every kernel launched through
// this helper calls pdl_wait() before its first dependent global access.  VLO_NO_PDL=1 disables the attribute.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device function attribute: set it once per (device, kernel).
int ensure_max_smem(const void* func, int bytes);

// Launch counter (the bench reports how many of OUR kernels ran in the timed region).
void count_launch(int n = 1);
long long launch_count();

// ---- per-kernel-class event profiler (runtime.cu); classes are indices into vlo_profile_read's arrays
enum ProfClass : int {
  PROF_GEMM_STREAM = 0,  // swap-AB weight-streaming GEMMs (decoder, connector, lm_head)
  PROF_ATTN = 1,         // attn_kvappend_kernel (the KV-append attention main kernel)
  PROF_ATTN_MERGE = 2,   // split-KV merge
  PROF_GEMM_VIT = 3,     // ViT trunk GEMMs (activations on MMA-M)
  PROF_VIT_ATTN = 4,
  PROF_OTHER = 5,
  PROF_NUM = 6,
};
void prof_enable(bool on);
bool prof_on();
void prof_begin(int cls, cudaStream_t st, double algo_bytes);
void prof_end(cudaStream_t st);
int prof_read(double* ms, long long* n, double* bytes, int ncls);

// ---- GEMM (gemm.cu) ---------------------------------------------------------------
struct GemmCall {
  int fmt;   // 0 fp16, 1 bf16
  int swap;  // 1: A = weights (MMA-M), B = tokens (MMA-N); logical out [rows_b, rows_a]
  int epi;   // GemmEpi
  int act;   // GemmAct
  const void* a;
  int rows_a;
  const void* b;
  int rows_b;
  int k;
  void* out;
  int ld_out;
  const float* bias;
  const float* pos;
  int pos_rows;
  int splits;          // >1 only with EPI_PARTIAL; out is then the fp32 workspace
  long long split_stride;
  int stream_weights;  // 1: operand A is read once (EVICT_FIRST), B is hot (EVICT_LAST)
  int bn;              // 0 = choose
};
int gemm_launch(const GemmCall& c, cudaStream_t stream);
int gemm_fix_splits(int k, int want);

// ---- persistent weight-streaming GEMM (gemm_ws.cuh / gemm.cu)
struct GemmWsCall {
  int fmt;            // 0 fp16, 1 bf16
  int mode;           // 0 stream-K fp32 partial planes, 1 whole tiles + 16-bit epilogue
  const void* w;      // [rows_w, k]
  int rows_w;
  const void* x;      // [rows_x, k], rows_x <= 128
  int rows_x;
  int k;
  void* out;
  int ld_out;
  long long plane_stride;
  const float* bias;
  int act;
  SkInfo sk;          // from gemm_ws_plan
  int bn;             // token-tile width (0 = smallest of 16/32/64/128 covering rows_x); 96 / 192 for the ViT
  int weights_hot;    // 1: weights are re-read by several token tiles (keep them in L2)
  int small_smem;     // 1: shallow 3-stage ring (~75 KB) so the CTA can share an SM with a decoder weight-streaming CTA
};
// decomposition for (rows_w, k): n_ctas <= 0 -> one CTA per SM; max_planes = fp32 planes the partials need
int gemm_ws_plan(int rows_w, int k, int mode, int n_ctas, SkInfo* sk, int* max_planes, int x_tiles = 1);
int gemm_ws_launch(const GemmWsCall& c, cudaStream_t stream);

// ---- 2-CTA (cta_group::2) tensor-bound GEMM for the ViT trunk (gemm2.cuh / gemm.cu): tokens on MMA-M, 256 x bn tiles per CTA pair
struct Gemm2Call {
  const void* x;      // [rows_x, k] fp16 activations
  int rows_x;
  const void* w;      // [rows_w, k] fp16 weights, rows_w a multiple of 32
  int rows_w;
  int k;
  void* out;          // epi 0: fp16 [rows_x, ld_out] = act(fp16(acc + bias));  epi 1: fp32 [rows_x, ld_out] += fp16(acc + bias)
  int ld_out;
  const float* bias;
  int act;
  int epi;
  int bn;             // 256 or 128 features per tile
};
int gemm2_launch(const Gemm2Call& c, cudaStream_t stream);

// ---- stream-K GEMM with the fused finisher epilogue (gemm_wsf.cuh / gemm.cu); args struct defined in gemm_wsf.cuh
struct GemmWsfArgs;
int gemm_wsf_launch(const GemmWsfArgs& a, const void* w, const void* x, int epi, cudaStream_t stream);

// ---- attention (attn.cu) ------------------------------------------------------------
struct AttnSeq {
  int q_tok0;          // first packed token of the sequence
  int q_len;           // new tokens
  int kv_len;          // cache length INCLUDING the new tokens
  long long kv_row0;   // row of (key 0, kv head 0) in the K / V matrices ([rows, head_dim])
  int kv_head_stride;  // rows between kv heads
};
// bytes of scratch attn_launch needs for these sequences (upper bound)
size_t attn_ws_bytes(int total_tokens, int n_seqs, int n_heads, int n_kv_heads);
size_t attn_stage_bytes(int total_tokens, int n_seqs, int n_heads, int n_kv_heads);
int attn_version(int n_heads, int n_kv_heads);
long long* attn_trace_buffer();  // device buffer of the VLO_ATTN_TRACE timeline (nullptr when off)  // 2 = tcgen05 kernel, 1 = mma.sync kernel (VLO_ATTN=1 or G not dividing 128)
struct AttnPlan {
  float* ws_o;
  float* ws_ml;
  void* d_items;
  int* d_tok_item;
  int n_items, max_splits, total_tokens;
  int* d_cta_tab;     // v3: flat (item, kv head, split) table, one entry per CTA
  int n_ctas;
  int blk;            // keys per pipeline block
  int version;        // 1 = mma.sync kernel, 2 = tcgen05 kernel (P in smem), 3 = tcgen05 kernel (P in TMEM, deep ring)
  int skip_merge;     // measurement only: launch the main kernel without the split-KV merge
  double algo_bytes;  // algorithmic HBM bytes of one attn_run over this plan (K+V rows read, Q read, out written)
};
// Build the work-item plan on the host and enqueue its upload.  h_stage (>= attn_stage_bytes())
// must stay untouched until the copies enqueued on `stream` have run.
int attn_plan(AttnPlan* plan, void* d_ws, void* h_stage, const AttnSeq* seqs, int n_seqs, int total_tokens,
              int n_heads, int n_kv_heads, int head_dim, cudaStream_t stream);
// d_k / d_v: one layer's K / V matrices [kv_rows, 128] bf16; d_q [T, n_heads, 128]; d_out [T, n_heads*128].
int attn_run(const AttnPlan& plan, const void* d_q, const void* d_k, const void* d_v, long long kv_rows, void* d_out,
             int n_heads, int n_kv_heads, int head_dim, cudaStream_t stream);

// 2D TMA map over a row-major 16-bit matrix [rows, k], box 64 x box_rows, 128-byte swizzle (cached).
int tmap_2d_sw128(const void* ptr, int rows, int k, int box_rows, int fmt, CUtensorMap* out);
int tmap_q3d_sw128(const void* ptr, int n_tok, int n_heads, int G, CUtensorMap* out, int box_rows = 128);  // attention Q tile (see gemm.cu)

}  // namespace vlo
