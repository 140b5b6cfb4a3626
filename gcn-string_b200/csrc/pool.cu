// K4 / K6 — GlobalSumPool: tf.math.segment_sum(X, i) over sorted graph ids and its gradient
// dX[n] = dOut[i[n]]  (reference call path gcn.py:320 -> GeneralGNN.pool).
// One streaming pass over [N, W]; HBM-bound.  Deterministic: fixed row->warp assignment and
// a fixed-order cross-warp sum.
#include "common.cuh"

namespace gcs {

constexpr int kPoolWarps = 8;

// grid (graph, column tile of 128 floats).  Lane l owns float4 column (tile*32 + l); warp w
// sums rows w, w+8, ... of the graph; the 8 partials are added in warp order.
__global__ void __launch_bounds__(kPoolWarps * 32) segment_sum_fwd_kernel(
    const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ graph_ptr, int W,
    float* __restrict__ out, int64_t ldo) {
  __shared__ float4 part[kPoolWarps][32];
  const int g = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (blockIdx.y * 32 + lane) * 4;
  const int r0 = __ldg(graph_ptr + g), r1 = __ldg(graph_ptr + g + 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < W) {
    const float* p = X + c;
    int r = r0 + warp;
    for (; r + 3 * kPoolWarps < r1; r += 4 * kPoolWarps) {   // 4 independent loads in flight
      const float4 a = __ldg(reinterpret_cast<const float4*>(p + static_cast<int64_t>(r) * ldx));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p + static_cast<int64_t>(r + kPoolWarps) * ldx));
      const float4 d = __ldg(reinterpret_cast<const float4*>(p + static_cast<int64_t>(r + 2 * kPoolWarps) * ldx));
      const float4 e = __ldg(reinterpret_cast<const float4*>(p + static_cast<int64_t>(r + 3 * kPoolWarps) * ldx));
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
      acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
      acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
    }
    for (; r < r1; r += kPoolWarps) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p + static_cast<int64_t>(r) * ldx));
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < W) {
    float4 s = part[0][lane];
#pragma unroll
    for (int w = 1; w < kPoolWarps; ++w) {
      const float4 t = part[w][lane];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    *reinterpret_cast<float4*>(out + static_cast<int64_t>(g) * ldo + c) = s;
  }
}

// Scalar variant for widths / alignments the float4 path cannot take.
__global__ void __launch_bounds__(kPoolWarps * 32) segment_sum_fwd_scalar_kernel(
    const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ graph_ptr, int W,
    float* __restrict__ out, int64_t ldo) {
  __shared__ float part[kPoolWarps][32];
  const int g = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + lane;
  const int r0 = __ldg(graph_ptr + g), r1 = __ldg(graph_ptr + g + 1);
  float acc = 0.f;
  if (c < W)
    for (int r = r0 + warp; r < r1; r += kPoolWarps) acc += __ldg(X + static_cast<int64_t>(r) * ldx + c);
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < W) {
    float s = part[0][lane];
    for (int w = 1; w < kPoolWarps; ++w) s += part[w][lane];
    out[static_cast<int64_t>(g) * ldo + c] = s;
  }
}

// grid (graph, column tile): broadcast dOut[g] to every row of the graph.
template <int VEC>
__global__ void __launch_bounds__(256) segment_sum_bwd_kernel(
    const float* __restrict__ dout, int64_t ldo, const int32_t* __restrict__ graph_ptr, int W,
    float* __restrict__ dX, int64_t ldx) {
  const int g = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (blockIdx.y * 32 + lane) * VEC;
  if (c >= W) return;
  const int r0 = __ldg(graph_ptr + g), r1 = __ldg(graph_ptr + g + 1);
  if (VEC == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(dout + static_cast<int64_t>(g) * ldo + c));
    for (int r = r0 + warp; r < r1; r += 8)
      *reinterpret_cast<float4*>(dX + static_cast<int64_t>(r) * ldx + c) = v;
  } else {
    const float v = __ldg(dout + static_cast<int64_t>(g) * ldo + c);
    for (int r = r0 + warp; r < r1; r += 8) dX[static_cast<int64_t>(r) * ldx + c] = v;
  }
}

}  // namespace gcs

using namespace gcs;

extern "C" int gcs_segment_sum_fwd(const float* X, int64_t ldx, const int32_t* graph_ptr, int32_t n_graphs,
                                   int32_t W, float* out, int64_t ldo, gcs_stream stream) {
  GCS_CHECK_ARG(n_graphs >= 0 && W > 0, "gcs_segment_sum_fwd: bad size");
  if (n_graphs == 0) return GCS_OK;
  GCS_CHECK_ARG(X && graph_ptr && out && ldx >= W && ldo >= W, "gcs_segment_sum_fwd: bad pointer / leading dimension");
  const bool vec = (W % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && aligned16(X) && aligned16(out);
  if (vec) {
    dim3 grid(n_graphs, static_cast<unsigned>(ceil_div(W, 128)));
    segment_sum_fwd_kernel<<<grid, kPoolWarps * 32, 0, as_stream(stream)>>>(X, ldx, graph_ptr, W, out, ldo);
  } else {
    dim3 grid(n_graphs, static_cast<unsigned>(ceil_div(W, 32)));
    segment_sum_fwd_scalar_kernel<<<grid, kPoolWarps * 32, 0, as_stream(stream)>>>(X, ldx, graph_ptr, W, out, ldo);
  }
  GCS_CHECK_LAUNCH("segment_sum_fwd_kernel");
  return GCS_OK;
}

extern "C" int gcs_segment_sum_bwd(const float* dout, int64_t ldo, const int32_t* graph_ptr, int32_t n_graphs,
                                   int32_t W, float* dX, int64_t ldx, gcs_stream stream) {
  GCS_CHECK_ARG(n_graphs >= 0 && W > 0, "gcs_segment_sum_bwd: bad size");
  if (n_graphs == 0) return GCS_OK;
  GCS_CHECK_ARG(dout && graph_ptr && dX && ldx >= W && ldo >= W, "gcs_segment_sum_bwd: bad pointer / leading dimension");
  const bool vec = (W % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && aligned16(dX) && aligned16(dout);
  if (vec) {
    dim3 grid(n_graphs, static_cast<unsigned>(ceil_div(W, 128)));
    segment_sum_bwd_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(dout, ldo, graph_ptr, W, dX, ldx);
  } else {
    dim3 grid(n_graphs, static_cast<unsigned>(ceil_div(W, 32)));
    segment_sum_bwd_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(dout, ldo, graph_ptr, W, dX, ldx);
  }
  GCS_CHECK_LAUNCH("segment_sum_bwd_kernel");
  return GCS_OK;
}
