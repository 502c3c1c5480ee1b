// C-ABI: library-level and kernel-level entry points (see include/vlo_b200.h).
#include "vlo_b200.h"

#include "common.h"
#include "gemm.cuh"

using namespace vlo;

extern "C" {

const char* vlo_last_error(void) { return last_error(); }

long long vlo_launch_count(void) { return launch_count(); }

int vlo_profile_enable(int on) {
  prof_enable(on != 0);
  return 0;
}
int vlo_profile_read(double* ms, long long* launches, double* algo_bytes, int n_classes) {
  return prof_read(ms, launches, algo_bytes, n_classes);
}

int vlo_device_supported(int device) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return prop.major == 10 ? 1 : 0;
}

int vlo_op_gemm(int fmt, int swap, int epi, int act, const void* d_a, int rows_a, const void* d_b,
                int rows_b, int k, void* d_out, int ld_out, const float* d_bias, const float* d_pos,
                int pos_rows, int splits, long long split_stride, int bn, void* cuda_stream) {
  GemmCall c{};
  c.fmt = fmt;
  c.swap = swap;
  c.epi = epi;
  c.act = act;
  c.a = d_a;
  c.rows_a = rows_a;
  c.b = d_b;
  c.rows_b = rows_b;
  c.k = k;
  c.out = d_out;
  c.ld_out = ld_out;
  c.bias = d_bias;
  c.pos = d_pos;
  c.pos_rows = pos_rows;
  c.splits = splits < 1 ? 1 : splits;
  c.split_stride = split_stride;
  c.stream_weights = swap;
  c.bn = bn;
  return gemm_launch(c, static_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"

#include <vector>

extern "C" {

int vlo_op_gemm2(const void* d_x, int rows_x, const void* d_w, int rows_w, int k, void* d_out, int ld_out, const float* d_bias,
                 int act, int epi, int bn, void* cuda_stream) {
  Gemm2Call c{};
  c.x = d_x;
  c.rows_x = rows_x;
  c.w = d_w;
  c.rows_w = rows_w;
  c.k = k;
  c.out = d_out;
  c.ld_out = ld_out;
  c.bias = d_bias;
  c.act = act;
  c.epi = epi;
  c.bn = bn;
  return gemm2_launch(c, static_cast<cudaStream_t>(cuda_stream));
}

int vlo_op_gemm_ws(int fmt, int mode, const void* d_w, int rows_w, const void* d_x, int rows_x, int k, void* d_out,
                   int ld_out, long long plane_stride, const float* d_bias, int act, int n_ctas, int bn, int* h_max_planes,
                   void* cuda_stream) {
  GemmWsCall c{};
  c.fmt = fmt;
  c.mode = mode;
  c.w = d_w;
  c.rows_w = rows_w;
  c.x = d_x;
  c.rows_x = rows_x;
  c.k = k;
  c.out = d_out;
  c.ld_out = ld_out;
  c.plane_stride = plane_stride;
  c.bias = d_bias;
  c.act = act;
  int planes = 1;
  c.bn = bn;
  c.weights_hot = rows_x > 128;
  const int eff_bn = bn > 0 ? bn : (rows_x <= 16 ? 16 : (rows_x <= 32 ? 32 : (rows_x <= 64 ? 64 : 128)));
  gemm_ws_plan(rows_w, k, mode, n_ctas, &c.sk, &planes, (rows_x + eff_bn - 1) / eff_bn);
  if (h_max_planes) *h_max_planes = planes;
  if (d_out == nullptr) return 0;  // planning query only
  return gemm_ws_launch(c, static_cast<cudaStream_t>(cuda_stream));
}

/* debug: copy the VLO_ATTN_TRACE timeline (n int64 values) to the host; returns -1 when tracing is off */
int vlo_debug_attn_trace(long long* h_out, int n) {
  long long* d = attn_trace_buffer();
  if (d == nullptr) return fail("attention tracing is off (set VLO_ATTN_TRACE=1)");
  VLO_CUDA(cudaMemcpy(h_out, d, sizeof(long long) * n, cudaMemcpyDeviceToHost));
  return 0;
}

int vlo_op_attn_version(int n_heads, int n_kv_heads) { return attn_version(n_heads, n_kv_heads); }

int64_t vlo_op_attn_ws_bytes(int n_tok, int n_heads, int head_dim, int kv_len) {
  (void)head_dim;
  (void)kv_len;
  // n_kv_heads is not part of this query: bound it by n_heads (G = 1 is the worst case)
  return 2 * static_cast<int64_t>(attn_ws_bytes(n_tok, 1, n_heads, n_heads));
}

int vlo_op_attn_kvappend(const void* d_q, const void* d_k, const void* d_v, void* d_out, float* d_ws, int n_tok,
                         int n_heads, int n_kv_heads, int head_dim, int kv_len, long long kv_stride,
                         void* cuda_stream) {
  AttnSeq s{};
  s.q_tok0 = 0;
  s.q_len = n_tok;
  s.kv_len = kv_len;
  s.kv_row0 = 0;
  s.kv_head_stride = static_cast<int>(kv_stride);
  // the staging copy must outlive the async H2D: keep it in a per-thread buffer and sync the copy
  static thread_local std::vector<uint8_t> stage;
  stage.resize(attn_stage_bytes(n_tok, 1, n_heads, n_kv_heads));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  AttnPlan plan{};
  int rc = attn_plan(&plan, d_ws, stage.data(), &s, 1, n_tok, n_heads, n_kv_heads, head_dim, st);
  if (rc != 0) return rc;
  rc = attn_run(plan, d_q, d_k, d_v, static_cast<long long>(n_kv_heads) * kv_stride, d_out, n_heads, n_kv_heads,
                head_dim, st);
  if (rc != 0) return rc;
  VLO_CUDA(cudaStreamSynchronize(st));  // pageable staging buffer: make the copies complete before returning
  return 0;
}

/* Back-to-back launches of the KV-append attention over `n_layers` separate K / V matrices (layer l at
 * d_k + l * layer_stride_rows * head_dim elements), `iters` passes: the micro-loop of vlo_bench_attn without an engine
 * (bench tools time it with one CUDA-event pair).  The plan is built and uploaded once, as in a decoder step. */
int vlo_op_attn_bench(const void* d_q, const void* d_k, const void* d_v, void* d_out, float* d_ws, int n_tok, int n_heads,
                      int n_kv_heads, int head_dim, int kv_len, long long kv_stride, int n_layers,
                      long long layer_stride_rows, int iters, int skip_merge, double* h_algo_bytes, void* cuda_stream) {
  VLO_CHECK(n_layers > 0 && iters > 0 && layer_stride_rows >= static_cast<long long>(n_kv_heads) * kv_stride, "attn_bench: sizes");
  AttnSeq s{};
  s.q_tok0 = 0;
  s.q_len = n_tok;
  s.kv_len = kv_len;
  s.kv_row0 = 0;
  s.kv_head_stride = static_cast<int>(kv_stride);
  static thread_local std::vector<uint8_t> stage;
  stage.resize(attn_stage_bytes(n_tok, 1, n_heads, n_kv_heads));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  AttnPlan plan{};
  if (attn_plan(&plan, d_ws, stage.data(), &s, 1, n_tok, n_heads, n_kv_heads, head_dim, st)) return -1;
  VLO_CUDA(cudaStreamSynchronize(st));
  plan.skip_merge = skip_merge;
  if (h_algo_bytes) *h_algo_bytes = plan.algo_bytes;
  const size_t lstride = static_cast<size_t>(layer_stride_rows) * head_dim * 2;
  for (int it = 0; it < iters; ++it)
    for (int l = 0; l < n_layers; ++l)
      if (attn_run(plan, d_q, static_cast<const uint8_t*>(d_k) + l * lstride, static_cast<const uint8_t*>(d_v) + l * lstride,
                   static_cast<long long>(n_kv_heads) * kv_stride, d_out, n_heads, n_kv_heads, head_dim, st))
        return -1;
  return 0;
}

}  // extern "C"
