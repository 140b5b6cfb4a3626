// K3 / K7 — sum aggregation  Y = pattern(A) . f(X),  f = fused BatchNorm + PReLU prologue.
//
// Replaces MessagePassing.propagate as GeneralConv.call invokes it:
//   messages = tf.gather(x, a.indices[:,1])          -> [nnz, H] materialised in TF
//   out      = tf.math.unsorted_segment_sum(messages, a.indices[:,0], N)
// Here the messages are never materialised, BN+PReLU are applied to X on the fly (they are
// a PROLOGUE of the aggregation in GeneralConv: transform -> BN -> PReLU -> aggregate), and
// Y is written straight into its slice of the 'cat' buffer (ldy = concat width).
// Determinism: each output element is owned by one thread which adds its neighbours in
// ascending column order — no atomics, no cross-lane reduction.
//
// Two kernels:
//  * spmm_graph_kernel  — one CTA per (graph, 32-column slab): the slab of f(X) for the
//    whole graph is staged ONCE in shared memory (each X element is read from HBM exactly
//    once and transformed once), neighbours are then gathered from shared memory with
//    128-bit loads (a quarter-warp per row: 8 lanes x float4 = 32 columns = 128 B,
//    conflict-free), Y rows are written with 128 B coalesced stores.  Graphs are served by
//    size class (256/512/1024/1600 staged rows -> 32/64/128/200 KB of shared memory and
//    128/256/512/1024 threads) so that small graphs keep several CTAs per SM resident; a
//    CTA whose graph is not in the launch's class exits at once.
//  * spmm_rows_kernel   — row-parallel (no graph structure given, odd widths, or graphs
//    above the largest class): H/4 lanes per row gather 128-bit pieces straight from
//    global memory (L1/L2 serve the re-reads), 4 neighbours in flight per lane.
#include "common.cuh"

namespace gcs {

constexpr int kSlab = 32;   // columns per CTA slab: one 128 B line per row

template <int THREADS, bool kTransform>
__global__ void __launch_bounds__(THREADS) spmm_graph_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
    const int32_t* __restrict__ graph_ptr, const float* __restrict__ X, int64_t ldx,
    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ alpha,
    float* __restrict__ Y, int64_t ldy, int lo_rows, int hi_rows) {
  extern __shared__ __align__(16) float4 s_x[];   // [rows][8] float4
  const int g = blockIdx.x;
  const int r0 = __ldg(graph_ptr + g);
  const int n = __ldg(graph_ptr + g + 1) - r0;
  if (n <= lo_rows || n > hi_rows) return;         // another size class serves this graph
  const int c0 = blockIdx.y * kSlab;
  const int lane8 = threadIdx.x & 7;               // float4 index inside the slab
  const int sub = threadIdx.x >> 3;                // row slot
  constexpr int ROWS = THREADS / 8;
  float4 sc, sh, al;
  if (kTransform) {
    sc = __ldg(reinterpret_cast<const float4*>(scale + c0) + lane8);
    sh = __ldg(reinterpret_cast<const float4*>(shift + c0) + lane8);
    al = __ldg(reinterpret_cast<const float4*>(alpha + c0) + lane8);
  }
  const float* xb = X + static_cast<int64_t>(r0) * ldx + c0;
#pragma unroll 4
  for (int r = sub; r < n; r += ROWS) {
    float4 v = __ldg(reinterpret_cast<const float4*>(xb + static_cast<int64_t>(r) * ldx) + lane8);
    if (kTransform) {
      v.x = bn_prelu(v.x, sc.x, sh.x, al.x);
      v.y = bn_prelu(v.y, sc.y, sh.y, al.y);
      v.z = bn_prelu(v.z, sc.z, sh.z, al.z);
      v.w = bn_prelu(v.w, sc.w, sh.w, al.w);
    }
    s_x[r * 8 + lane8] = v;
  }
  __syncthreads();
  float* yb = Y + static_cast<int64_t>(r0) * ldy + c0;
  const int32_t* rp = rowptr + r0;
  for (int r = sub; r < n; r += ROWS) {
    const int eb = __ldg(rp + r), ee = __ldg(rp + r + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int e = eb;
    for (; e + 4 <= ee; e += 4) {
      const int j0 = __ldg(colidx + e) - r0, j1 = __ldg(colidx + e + 1) - r0;
      const int j2 = __ldg(colidx + e + 2) - r0, j3 = __ldg(colidx + e + 3) - r0;
      const float4 v0 = s_x[j0 * 8 + lane8], v1 = s_x[j1 * 8 + lane8];
      const float4 v2 = s_x[j2 * 8 + lane8], v3 = s_x[j3 * 8 + lane8];
      acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
      acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
      acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
      acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
    }
    for (; e < ee; ++e) {
      const float4 v = s_x[(__ldg(colidx + e) - r0) * 8 + lane8];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(yb + static_cast<int64_t>(r) * ldy)[lane8] = acc;
  }
}

// Row-parallel kernel.  VEC = 4: H % 4 == 0 and 16 B-aligned rows, lanes = H/4 threads per
// row.  VEC = 1: any H, lanes = H threads per row.  With graph_ptr != nullptr only graphs
// with more than min_rows rows are processed (blockIdx.y strides over graphs); with
// graph_ptr == nullptr all n_rows rows are.
template <int VEC, bool kTransform>
__global__ void __launch_bounds__(256) spmm_rows_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int64_t n_rows,
    const float* __restrict__ X, int64_t ldx, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ alpha, float* __restrict__ Y,
    int64_t ldy, int H, int lanes, int rows_per_block, const int32_t* __restrict__ graph_ptr,
    int n_graphs, int min_rows) {
  const int lane = threadIdx.x % lanes;
  const int slot = threadIdx.x / lanes;
  const int c = lane * VEC;
  if (slot >= rows_per_block || c >= H) return;
  float sc[VEC], sh[VEC], al[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    sc[k] = kTransform ? __ldg(scale + c + k) : 1.f;
    sh[k] = kTransform ? __ldg(shift + c + k) : 0.f;
    al[k] = kTransform ? __ldg(alpha + c + k) : 1.f;
  }
  auto load = [&](int j, float* v) {
    const float* p = X + static_cast<int64_t>(j) * ldx + c;
    if (VEC == 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
      v[0] = __ldg(p);
    }
    if (kTransform) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] = bn_prelu(v[k], sc[k], sh[k], al[k]);
    }
  };
  auto do_row = [&](int64_t r) {
    const int eb = __ldg(rowptr + r), ee = __ldg(rowptr + r + 1);
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    int e = eb;
    for (; e + 4 <= ee; e += 4) {
      const int j0 = __ldg(colidx + e), j1 = __ldg(colidx + e + 1);
      const int j2 = __ldg(colidx + e + 2), j3 = __ldg(colidx + e + 3);
      float v0[VEC], v1[VEC], v2[VEC], v3[VEC];
      load(j0, v0); load(j1, v1); load(j2, v2); load(j3, v3);
#pragma unroll
      for (int k = 0; k < VEC; ++k) { acc[k] += v0[k]; acc[k] += v1[k]; acc[k] += v2[k]; acc[k] += v3[k]; }
    }
    for (; e < ee; ++e) {
      float v[VEC];
      load(__ldg(colidx + e), v);
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] += v[k];
    }
    float* q = Y + r * ldy + c;
    if (VEC == 4) *reinterpret_cast<float4*>(q) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else q[0] = acc[0];
  };
  if (graph_ptr == nullptr) {
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * rows_per_block + slot; r < n_rows;
         r += static_cast<int64_t>(gridDim.x) * rows_per_block)
      do_row(r);
  } else {
    for (int g = blockIdx.y; g < n_graphs; g += gridDim.y) {
      const int rb = __ldg(graph_ptr + g), re = __ldg(graph_ptr + g + 1);
      if (re - rb <= min_rows) continue;
      for (int64_t r = rb + static_cast<int64_t>(blockIdx.x) * rows_per_block + slot; r < re;
           r += static_cast<int64_t>(gridDim.x) * rows_per_block)
        do_row(r);
    }
  }
}

}  // namespace gcs

using namespace gcs;

namespace {

struct SpmmArgs {
  const int32_t* rowptr; const int32_t* colidx; const int32_t* graph_ptr; int n_graphs; int64_t n_rows;
  const float* X; int64_t ldx; const float* scale; const float* shift; const float* alpha;
  float* Y; int64_t ldy; int H; cudaStream_t st;
};

template <int VEC>
int launch_rows(const SpmmArgs& a, const int32_t* graph_ptr, int min_rows) {
  const int lanes = (a.H + VEC - 1) / VEC;
  if (lanes > 256) return fail(GCS_ERR_UNSUPPORTED, "gcs_spmm_sum: H=%d too wide for the row kernel", a.H);
  const int rows_per_block = 256 / lanes;
  dim3 grid;
  if (graph_ptr == nullptr) {
    int64_t blocks = ceil_div(a.n_rows, rows_per_block);
    const int64_t cap = 16LL * sm_count();
    grid = dim3(static_cast<unsigned>(blocks > cap ? cap : blocks));
  } else {
    grid = dim3(8, a.n_graphs > 65535 ? 65535 : a.n_graphs);   // 8 row-blocks per oversize graph
  }
  if (a.scale)
    spmm_rows_kernel<VEC, true><<<grid, 256, 0, a.st>>>(a.rowptr, a.colidx, a.n_rows, a.X, a.ldx, a.scale, a.shift, a.alpha, a.Y, a.ldy, a.H, lanes, rows_per_block, graph_ptr, a.n_graphs, min_rows);
  else
    spmm_rows_kernel<VEC, false><<<grid, 256, 0, a.st>>>(a.rowptr, a.colidx, a.n_rows, a.X, a.ldx, a.scale, a.shift, a.alpha, a.Y, a.ldy, a.H, lanes, rows_per_block, graph_ptr, a.n_graphs, min_rows);
  GCS_CHECK_LAUNCH("spmm_rows_kernel");
  return GCS_OK;
}

template <int THREADS>
int launch_graph_class(const SpmmArgs& a, int lo_rows, int hi_rows) {
  const int smem = hi_rows * kSlab * static_cast<int>(sizeof(float));
  auto kt = spmm_graph_kernel<THREADS, true>;
  auto kf = spmm_graph_kernel<THREADS, false>;
  static bool attr_set = false;   // per template instantiation
  if (!attr_set) {
    GCS_CUDA(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    GCS_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid(a.n_graphs, a.H / kSlab);
  if (a.scale)
    kt<<<grid, THREADS, smem, a.st>>>(a.rowptr, a.colidx, a.graph_ptr, a.X, a.ldx, a.scale, a.shift, a.alpha, a.Y, a.ldy, lo_rows, hi_rows);
  else
    kf<<<grid, THREADS, smem, a.st>>>(a.rowptr, a.colidx, a.graph_ptr, a.X, a.ldx, a.scale, a.shift, a.alpha, a.Y, a.ldy, lo_rows, hi_rows);
  GCS_CHECK_LAUNCH("spmm_graph_kernel");
  return GCS_OK;
}

int g_spmm_mode = 0;   // 0 = auto, 1 = force row kernel, 2 = force staged kernel

}  // namespace

// Test / sweep hook (not part of the drop-in surface).
extern "C" void gcs_debug_set_spmm_mode(int mode) { g_spmm_mode = mode; }

extern "C" int gcs_spmm_sum(const int32_t* rowptr, const int32_t* colidx, const int32_t* graph_ptr,
                            int32_t n_graphs, int32_t max_graph_rows, int64_t n_rows, const float* X,
                            int64_t ldx, const float* scale, const float* shift, const float* alpha,
                            float* Y, int64_t ldy, int32_t H, gcs_stream stream) {
  GCS_CHECK_ARG(n_rows >= 0 && H > 0, "gcs_spmm_sum: bad size (n_rows=%lld, H=%d)", (long long)n_rows, H);
  if (n_rows == 0) return GCS_OK;
  GCS_CHECK_ARG(rowptr && colidx && X && Y, "gcs_spmm_sum: null pointer");
  GCS_CHECK_ARG(ldx >= H && ldy >= H, "gcs_spmm_sum: leading dimension smaller than H");
  GCS_CHECK_ARG((scale != nullptr) == (shift != nullptr) && (scale != nullptr) == (alpha != nullptr),
                "gcs_spmm_sum: scale/shift/alpha must be all NULL or all set");
  GCS_CHECK_ARG(X != Y, "gcs_spmm_sum: in-place aggregation is not defined");
  GCS_CHECK_ARG(n_rows < INT32_MAX, "gcs_spmm_sum: n_rows exceeds int32 CSR range");
  SpmmArgs a{rowptr, colidx, graph_ptr, n_graphs, n_rows, X, ldx, scale, shift, alpha, Y, ldy, H, as_stream(stream)};
  const bool vec_ok = (H % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(X) && aligned16(Y) &&
                      (!scale || (aligned16(scale) && aligned16(shift) && aligned16(alpha)));
  const bool staged = g_spmm_mode != 1 && graph_ptr && n_graphs > 0 && n_graphs <= 0x7fffffff / 1 &&
                      vec_ok && H % kSlab == 0 && H / kSlab <= 65535;
  if (!staged) {
    if (vec_ok) return launch_rows<4>(a, nullptr, 0);
    return launch_rows<1>(a, nullptr, 0);
  }
  // Size classes; max_graph_rows (0 = unknown) lets the host skip classes no graph is in.
  const int mx = max_graph_rows > 0 ? max_graph_rows : INT32_MAX;
  GCS_TRY(launch_graph_class<128>(a, 0, 256));
  if (mx > 256) GCS_TRY(launch_graph_class<256>(a, 256, 512));
  if (mx > 512) GCS_TRY(launch_graph_class<512>(a, 512, 1024));
  if (mx > 1024) GCS_TRY(launch_graph_class<1024>(a, 1024, 1600));
  if (mx > 1600) GCS_TRY(launch_rows<4>(a, graph_ptr, 1600));
  return GCS_OK;
}
