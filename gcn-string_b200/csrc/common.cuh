// Shared helpers for the gcnstring_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/gcnstring_b200.h"

namespace gcs {

// Thread-local last-error string behind gcs_last_error().
char* error_buffer();
int fail(int status, const char* fmt, ...);

inline cudaStream_t as_stream(gcs_stream s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs on the current device (cached per device).
int sm_count();

int exclusive_scan_i32(const int32_t* cnt, int64_t n, int32_t* out, cudaStream_t st);   // batching.cu
int exclusive_scan_i64(const int32_t* cnt, int64_t n, int64_t* out, cudaStream_t st);   // batching.cu (int64 running total)

// Synchronised-BatchNorm hook (gcs_set_allreduce_hook), per host thread.
struct SyncHook {
  gcs_allreduce_fn fn = nullptr;
  void* user = nullptr;
  int world = 1;
};
SyncHook& sync_hook();

// Running |max| of the tensors the fp16 tensor-core GEMM reads (linear_tc.cu): while `produce` is set, the kernels
// that write activations / gradients (bn_prelu_fwd, the aggregation, the BatchNorm backward apply pass) fold the |max|
// of what they write into *produce (atomicMax on the bit pattern; the caller zeroes the cell); while `consume` is set,
// the dense transforms take it as the |max| of their A operand.  Per host thread, set by model.cu around its calls.
struct AmaxSink {
  float* produce = nullptr;
  const float* consume = nullptr;       // |max| of the A operand of gcs_linear_fwd / gcs_linear_bwd_input / dense_dx_concat,
                                        // and of dH in gcs_linear_bwd_weight
  const float* consume_act = nullptr;   // |max| of the activations A in gcs_linear_bwd_weight
};
AmaxSink& amax_sink();

// Fold a thread's |max| (m >= 0) into *cell (no-op for cell == nullptr): one REDUX over the lanes that are here, then
// one atomic by the first of them, and only if it would raise the cell.  Safe under divergence.
__device__ __forceinline__ void amax_commit(float m, float* cell) {
  if (!cell) return;
  const unsigned mask = __activemask();
  const unsigned bits = __reduce_max_sync(mask, __float_as_uint(m));      // non-negative floats order like their bits
  if ((threadIdx.x & 31) == static_cast<unsigned>(__ffs(mask) - 1) && bits > *reinterpret_cast<volatile unsigned*>(cell))
    atomicMax(reinterpret_cast<unsigned*>(cell), bits);
}
__device__ __forceinline__ float amax4(float m, const float4& v) {
  return fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
}

// Count of kernel launches issued by this library (bench.py's `gpu_launches`).
void count_launch();

// Optional per-op device timing (gcs_debug_profile_*): CUDA events on the launching stream
// around each op of the model entry points.  Off by default; no cost when off.
struct ScopedOpTimer {
  int slot;
  cudaStream_t st;
  ScopedOpTimer(const char* label, gcs_stream stream);
  ~ScopedOpTimer();
};

#define GCS_CHECK_ARG(cond, ...)                                              \
  do {                                                                        \
    if (!(cond)) return ::gcs::fail(GCS_ERR_INVALID_ARGUMENT, __VA_ARGS__);   \
  } while (0)

#define GCS_CHECK_LAUNCH(name)                                                              \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess)                                                                 \
      return ::gcs::fail(GCS_ERR_CUDA, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
    ::gcs::count_launch();                                                                  \
  } while (0)

#define GCS_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return ::gcs::fail(GCS_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__));           \
  } while (0)

#define GCS_TRY(call)               \
  do {                              \
    int st__ = (call);              \
    if (st__ != GCS_OK) return st__; \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// f(x) = prelu(x*scale + shift, alpha): the BatchNorm + PReLU prologue of GeneralConv.
// Keras PReLU: relu(z) - alpha*relu(-z)  ->  z > 0 ? z : alpha*z.
__device__ __forceinline__ float bn_prelu(float x, float sc, float sh, float al) {
  float z = fmaf(x, sc, sh);
  return z > 0.f ? z : al * z;
}

}  // namespace gcs
