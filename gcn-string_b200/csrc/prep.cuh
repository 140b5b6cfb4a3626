// Weights of one model step split ahead of time (fp16 hi / lo parts + scale cells) in two launches instead of three
// small launches in front of every tensor-core GEMM.  model.cu lists the GEMMs of the step (prepare_step_weights), the
// GEMM wrappers in linear.cu look their operand up (find_prepared) and skip their own |max| / split kernels on a hit.
#pragma once
#include <cstdint>

#include "common.cuh"

namespace gcs {
namespace tc {
// One weight block of a batched split: W [rows, cols] with row pitch ldw -> fp16 hi / lo parts at out[r * ldo + c] (or
// [c * ldo + r] when transposed), scaled by f16_scale(cells[0]); cells[1] receives 1 / scale.  hi / lo point at __half.
struct SplitJob {
  const float* W; int rows, cols; int64_t ldw; int transpose; int64_t ldo; void* hi; void* lo; float* cells;
};
constexpr int kMaxSplitJobs = 40;
struct SplitJobs { SplitJob j[kMaxSplitJobs]; int n; };
int split_f16_multi(const SplitJobs& jobs, cudaStream_t st);   // linear_tc.cu; the cells must be zero
int64_t f16_workspace_bytes(int K, int N);
float* f16_cells(void* workspace, int K, int N);
int f16_mode();
}  // namespace tc

// kind: 0 = forward (linear_fwd_fused: W [K, N] -> Bt [N][K]), 1 = concatenated input gradient (dense_dx_concat, keyed by
// its first block), 2 = plain input gradient (gcs_linear_bwd_input: W [K, N] as stored).  ws: a region laid out like the
// per-call workspace of those wrappers (hi | lo | cells).
struct PreparedWeights { int kind; const float* w; int kred, nout; void* ws; };
struct PreparedTable {
  PreparedWeights e[tc::kMaxSplitJobs];
  int n = 0;
};
PreparedTable*& prepared_table();                       // per host thread; set for the duration of a model entry point
void* find_prepared(int kind, const float* w, int kred, int nout);
struct PreparedScope {                                  // RAII: the table is never left set between ABI calls
  PreparedTable* saved;
  explicit PreparedScope(PreparedTable* t) : saved(prepared_table()) { prepared_table() = t; }
  ~PreparedScope() { prepared_table() = saved; }
};
}  // namespace gcs
