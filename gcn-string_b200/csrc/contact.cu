// K11 / K12 — residue contact maps and inter-protein pair graphs on the device.
//
// Replaces, for a whole batch of chains / pairs per launch:
//   GraphMaker.calculate_residue_dist / calculate_dist_matrix / generate_proximity_matrix
//     (src/utilities/gcn_utills.py:161-238): an O(n^2) Python double loop over Bio.PDB residues,
//       diff = seq_1["CA"].coord - seq_2["CA"].coord           (float32 arrays of 3)
//       d    = np.sqrt(np.sum(diff * diff))                    (float32: ((dx*dx + dy*dy) + dz*dz), then sqrt)
//       adjacency[d < angstroms] = 1                           (diagonal: d = 0 -> self-loops)
//   GraphMaker.generate_graphs (nx.from_numpy_matrix, :240-270) and link_graphs (:319-377):
//       U = nx.union(G_1, G_2, rename=('a-', 'b-')); U.add_edge('a-' + b1, 'b-' + b2) per DCA bridge
//   and the adjacency gcn.py:104-117,184-197 takes from U (node order a then b, 0/1 pattern, both directions).
// The distance is evaluated with the reference's float32 operation order and NO fused multiply-add, so the
// `d < angstroms` decision - and with it every integer output - is bit-identical to NumPy's.
// Output layout = the packed dataset K0 consumes (node_off / rowptr int64, graph-local int32 columns, ascending).
#include "common.cuh"

namespace gcs {

__device__ __forceinline__ float ca_distance(float xi, float yi, float zi, float xj, float yj, float zj) {
  const float dx = __fsub_rn(xi, xj), dy = __fsub_rn(yi, yj), dz = __fsub_rn(zi, zj);
  const float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  return __fsqrt_rn(s);
}

// Largest k in [0, n) with ptr[k] <= v (ptr non-decreasing, ptr[0] <= v < ptr[n]).
template <typename T>
__device__ __forceinline__ int segment_of(const T* __restrict__ ptr, int n, T v) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(ptr + mid) <= v) lo = mid; else hi = mid;
  }
  return lo;
}

// One warp per residue (row); lanes stride over the residues of the same chain, a ballot keeps the columns ascending.
template <bool kFill>
__global__ void __launch_bounds__(256) contact_kernel(const float* __restrict__ ca, const int32_t* __restrict__ chain_ptr,
                                                      int n_chains, int64_t n_res, float thr, int32_t* __restrict__ count,
                                                      const int64_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                                      float* __restrict__ dist) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_res) return;
  const int c = segment_of<int32_t>(chain_ptr, n_chains, static_cast<int32_t>(row));
  const int c0 = __ldg(chain_ptr + c), c1 = __ldg(chain_ptr + c + 1);
  const float xi = __ldg(ca + 3 * row), yi = __ldg(ca + 3 * row + 1), zi = __ldg(ca + 3 * row + 2);
  const int64_t base = kFill ? __ldg(rowptr + row) : 0;
  int n = 0;
  for (int j0 = c0; j0 < c1; j0 += 32) {
    const int j = j0 + lane;
    float d = 0.f;
    bool hit = false;
    if (j < c1) {
      const float* q = ca + 3 * static_cast<int64_t>(j);
      d = ca_distance(xi, yi, zi, __ldg(q), __ldg(q + 1), __ldg(q + 2));
      hit = d < thr;                                    // NaN coordinates never make a contact, as in NumPy
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (kFill && hit) {
      const int64_t pos = base + n + __popc(m & ((1u << lane) - 1u));
      col[pos] = j - c0;
      if (dist) dist[pos] = d;
    }
    n += __popc(m);
  }
  if (!kFill && lane == 0) count[row] = n;
}

__global__ void pair_nodes_kernel(const int32_t* __restrict__ chain_ptr, int n_chains, const int32_t* __restrict__ pair_a,
                                  const int32_t* __restrict__ pair_b, int n_pairs, int32_t* __restrict__ n_nodes,
                                  int32_t* __restrict__ status) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const int a = pair_a[p], b = pair_b[p];
  if (a < 0 || a >= n_chains || b < 0 || b >= n_chains) {
    atomicOr(status, 1);
    n_nodes[p] = 0;
    return;
  }
  n_nodes[p] = (chain_ptr[a + 1] - chain_ptr[a]) + (chain_ptr[b + 1] - chain_ptr[b]);
}

// One thread per row of the pair graphs.  A row of chain a keeps its contacts and gains, after them, its bridge
// partners n_a + b2 (ascending, duplicates merged); a row of chain b gains its partners b1 in front of its contacts.
template <bool kFill>
__global__ void __launch_bounds__(256) link_kernel(
    const int64_t* __restrict__ c_rowptr, const int32_t* __restrict__ c_col, const int32_t* __restrict__ chain_ptr,
    const int32_t* __restrict__ pair_a, const int32_t* __restrict__ pair_b, int n_pairs,
    const int32_t* __restrict__ bridge_ptr, const int32_t* __restrict__ bridge_a, const int32_t* __restrict__ bridge_b,
    const int64_t* __restrict__ node_off, int64_t n_rows, int32_t* __restrict__ count, const int64_t* __restrict__ rowptr,
    int32_t* __restrict__ col, int32_t* __restrict__ status) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const int p = segment_of<int64_t>(node_off, n_pairs, row);
  const int a = pair_a[p], b = pair_b[p];
  const int a0 = chain_ptr[a], na = chain_ptr[a + 1] - a0, b0 = chain_ptr[b], nb = chain_ptr[b + 1] - b0;
  const int r = static_cast<int>(row - node_off[p]);
  const bool in_a = r < na;
  const int64_t crow = in_a ? a0 + r : b0 + (r - na);
  const int64_t e0 = c_rowptr[crow], e1 = c_rowptr[crow + 1];
  const int q0 = bridge_ptr[p], q1 = bridge_ptr[p + 1];
  const int me = in_a ? r : r - na;
  const int32_t* mine = in_a ? bridge_a : bridge_b;     // this row's side of a bridge
  const int32_t* other = in_a ? bridge_b : bridge_a;
  const int n_other = in_a ? nb : na;
  // bridge partners in ascending order without duplicates: repeated selection of the next larger one (<= ~20 bridges)
  int64_t out = kFill ? rowptr[row] : 0;
  int n = 0;
  if (kFill && in_a)
    for (int64_t e = e0; e < e1; ++e) col[out++] = c_col[e];
  int last = -1;
  while (true) {
    int best = INT32_MAX;
    for (int q = q0; q < q1; ++q) {
      const int o = other[q];
      if (mine[q] < 0 || mine[q] >= (in_a ? na : nb) || o < 0 || o >= n_other) {
        atomicOr(status, 2);                            // networkx would silently add a NEW node here
        continue;
      }
      if (mine[q] == me && o > last && o < best) best = o;
    }
    if (best == INT32_MAX) break;
    if (kFill) col[out++] = in_a ? na + best : best;
    last = best;
    ++n;
  }
  if (kFill && !in_a)
    for (int64_t e = e0; e < e1; ++e) col[out++] = na + c_col[e];
  if (!kFill) count[row] = static_cast<int32_t>(e1 - e0) + n;
}

}  // namespace gcs

using namespace gcs;

extern "C" int64_t gcs_contact_workspace_bytes(int64_t n_rows) {
  return round_up((n_rows > 0 ? n_rows : 1) * static_cast<int64_t>(sizeof(int32_t)), 256);
}

extern "C" int gcs_contact_map_rowptr(const float* ca, const int32_t* chain_ptr, int32_t n_chains, int64_t n_residues,
                                      float angstroms, int64_t* rowptr, void* workspace, int64_t workspace_bytes,
                                      gcs_stream stream) {
  GCS_CHECK_ARG(chain_ptr && rowptr && n_chains >= 0 && n_residues >= 0, "gcs_contact_map_rowptr: bad argument");
  GCS_CHECK_ARG(n_residues < INT32_MAX, "gcs_contact_map_rowptr: residue offsets are int32");
  GCS_CHECK_ARG(n_residues == 0 || (ca && workspace && n_chains > 0), "gcs_contact_map_rowptr: null pointer");
  if (workspace_bytes < gcs_contact_workspace_bytes(n_residues))
    return fail(GCS_ERR_WORKSPACE, "gcs_contact_map_rowptr: workspace too small");
  cudaStream_t st = as_stream(stream);
  if (n_residues == 0) {
    GCS_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int64_t), st));
    return GCS_OK;
  }
  int32_t* cnt = static_cast<int32_t*>(workspace);
  contact_kernel<false><<<static_cast<unsigned>(ceil_div(n_residues, 8)), 256, 0, st>>>(ca, chain_ptr, n_chains, n_residues, angstroms, cnt, nullptr, nullptr, nullptr);
  GCS_CHECK_LAUNCH("contact_kernel<count>");
  return exclusive_scan_i64(cnt, n_residues, rowptr, st);
}

extern "C" int gcs_contact_map_fill(const float* ca, const int32_t* chain_ptr, int32_t n_chains, int64_t n_residues,
                                    float angstroms, const int64_t* rowptr, int32_t* col, float* dist, gcs_stream stream) {
  GCS_CHECK_ARG(chain_ptr && rowptr && n_chains >= 0 && n_residues >= 0, "gcs_contact_map_fill: bad argument");
  GCS_CHECK_ARG(n_residues < INT32_MAX, "gcs_contact_map_fill: residue offsets are int32");
  if (n_residues == 0) return GCS_OK;
  GCS_CHECK_ARG(ca && col && n_chains > 0, "gcs_contact_map_fill: null pointer");
  contact_kernel<true><<<static_cast<unsigned>(ceil_div(n_residues, 8)), 256, 0, as_stream(stream)>>>(ca, chain_ptr, n_chains, n_residues, angstroms, nullptr, rowptr, col, dist);
  GCS_CHECK_LAUNCH("contact_kernel<fill>");
  return GCS_OK;
}

extern "C" int gcs_link_pairs_offsets(const int32_t* chain_ptr, int32_t n_chains, const int32_t* pair_a, const int32_t* pair_b,
                                      int32_t n_pairs, int64_t* node_off, int32_t* status_dev, void* workspace,
                                      int64_t workspace_bytes, gcs_stream stream) {
  GCS_CHECK_ARG(chain_ptr && node_off && status_dev && n_chains >= 0 && n_pairs >= 0, "gcs_link_pairs_offsets: bad argument");
  cudaStream_t st = as_stream(stream);
  GCS_CUDA(cudaMemsetAsync(status_dev, 0, sizeof(int32_t), st));
  if (n_pairs == 0) {
    GCS_CUDA(cudaMemsetAsync(node_off, 0, sizeof(int64_t), st));
    return GCS_OK;
  }
  GCS_CHECK_ARG(pair_a && pair_b && workspace, "gcs_link_pairs_offsets: null pointer");
  if (workspace_bytes < gcs_contact_workspace_bytes(n_pairs))
    return fail(GCS_ERR_WORKSPACE, "gcs_link_pairs_offsets: workspace too small");
  int32_t* cnt = static_cast<int32_t*>(workspace);
  pair_nodes_kernel<<<static_cast<unsigned>(ceil_div(n_pairs, 256)), 256, 0, st>>>(chain_ptr, n_chains, pair_a, pair_b, n_pairs, cnt, status_dev);
  GCS_CHECK_LAUNCH("pair_nodes_kernel");
  return exclusive_scan_i64(cnt, n_pairs, node_off, st);
}

extern "C" int gcs_link_pairs(const int64_t* chain_rowptr, const int32_t* chain_col, const int32_t* chain_ptr,
                              const int32_t* pair_a, const int32_t* pair_b, int32_t n_pairs, const int32_t* bridge_ptr,
                              const int32_t* bridge_a, const int32_t* bridge_b, const int64_t* node_off, int64_t n_rows,
                              int64_t* rowptr, int32_t* col, int32_t* status_dev, void* workspace, int64_t workspace_bytes,
                              gcs_stream stream) {
  GCS_CHECK_ARG(n_pairs >= 0 && n_rows >= 0 && rowptr && status_dev, "gcs_link_pairs: bad argument");
  cudaStream_t st = as_stream(stream);
  if (n_rows == 0) {
    if (!col) GCS_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int64_t), st));
    return GCS_OK;
  }
  GCS_CHECK_ARG(chain_rowptr && chain_col && chain_ptr && pair_a && pair_b && bridge_ptr && node_off,
                "gcs_link_pairs: null pointer");
  const unsigned grid = static_cast<unsigned>(ceil_div(n_rows, 256));
  if (!col) {                                            // phase 1: row pointers
    GCS_CHECK_ARG(workspace, "gcs_link_pairs: null workspace");
    if (workspace_bytes < gcs_contact_workspace_bytes(n_rows))
      return fail(GCS_ERR_WORKSPACE, "gcs_link_pairs: workspace too small");
    int32_t* cnt = static_cast<int32_t*>(workspace);
    link_kernel<false><<<grid, 256, 0, st>>>(chain_rowptr, chain_col, chain_ptr, pair_a, pair_b, n_pairs, bridge_ptr, bridge_a, bridge_b, node_off, n_rows, cnt, nullptr, nullptr, status_dev);
    GCS_CHECK_LAUNCH("link_kernel<count>");
    return exclusive_scan_i64(cnt, n_rows, rowptr, st);
  }
  link_kernel<true><<<grid, 256, 0, st>>>(chain_rowptr, chain_col, chain_ptr, pair_a, pair_b, n_pairs, bridge_ptr, bridge_a, bridge_b, node_off, n_rows, nullptr, rowptr, col, status_dev);
  GCS_CHECK_LAUNCH("link_kernel<fill>");
  return GCS_OK;
}
