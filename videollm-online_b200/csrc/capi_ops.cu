// C-ABI: library-level and kernel-level entry points (see include/vlo_b200.h).
#include "vlo_b200.h"

#include "common.h"
#include "gemm.cuh"

using namespace vlo;

extern "C" {

const char* vlo_last_error(void) { return last_error(); }

long long vlo_launch_count(void) { return launch_count(); }

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
