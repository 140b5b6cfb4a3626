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
//  * spmm_tile_kernel — one CTA per (row tile of <= 512 rows, 32-column slab), 512 threads.
//    Row tiles follow graph boundaries when the caller has them (gcs_spmm_build_tiles packs
//    whole graphs into tiles and splits only graphs larger than a tile), so nearly every
//    neighbour of a tile row is itself a tile row.  Per CTA:
//      1. stage f(X[tile rows, slab]) in shared memory: each X element is read from HBM once
//         (128-bit loads, one full 128 B line per row) and transformed once;
//      2. stage the tile's neighbour lists, pre-decoded to 16-bit slot numbers; a neighbour
//         outside the tile (split graph, or no tile structure given) gets a "ghost" slot;
//      3. fetch the ghost rows in one cooperative, fully parallel pass (global -> transform ->
//         shared): no long-latency access is left on the gather path;
//      4. gather: a quarter-warp per output row, 8 lanes x float4 = 128 B per neighbour from
//         shared memory (conflict-free), 4 neighbours per iteration, ascending column order;
//         Y rows are written with 128 B coalesced stores straight into the concat slice.
//    ~98 KB of shared memory per CTA -> 2 CTAs (32 warps) per SM, so the HBM-bound staging of
//    one CTA overlaps the shared-memory-bound gathering of the other.
//  * spmm_rows_kernel — row-parallel fallback (odd widths / unaligned views): H/4 lanes per
//    row gather 128-bit pieces straight from global memory, 4 neighbours in flight per lane.
#include "common.cuh"

namespace gcs {

constexpr int kTileRows = 512;      // max rows staged per CTA (64 KB of 128 B rows)
constexpr int kTileThreads = 512;   // 64 row slots x 8 lanes
constexpr int kGhostRows = 128;     // out-of-tile neighbour rows staged per pass (16 KB)
constexpr int kIdxCap = 8192;       // staged neighbour entries per pass (16 KB of uint16)
constexpr int kOverflow = 0xFFFF;   // entry that found no ghost slot: fetched from global memory
constexpr int kTileSmemBytes = (kTileRows + kGhostRows) * 128 + kIdxCap * 2 + (kTileRows + 1) * 4 + kGhostRows * 4 + 16;

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return static_cast<int>(v);
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ void sts32(uint32_t addr, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v)); }
__device__ __forceinline__ void sts16(uint32_t addr, int v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<unsigned short>(v)));
}
__device__ __forceinline__ int atoms_add(uint32_t addr, int v) {
  int old;
  asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v));
  return old;
}

// tile_ptr == nullptr: uniform tiles of kTileRows rows.  Otherwise tile t covers rows
// [tile_ptr[t], tile_ptr[t+1]) (<= kTileRows each), t < *n_tiles_dev.
template <bool kTransform>
__global__ void __launch_bounds__(kTileThreads, 2) spmm_tile_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int n_rows,
    const int32_t* __restrict__ tile_ptr, const int32_t* __restrict__ n_tiles_dev,
    const float* __restrict__ X, int64_t ldx, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ alpha, float* __restrict__ Y, int64_t ldy) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t sx = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));   // [kTileRows + kGhostRows][8] float4
  const uint32_t sidx = sx + (kTileRows + kGhostRows) * 128;                       // [kIdxCap] uint16 slot numbers
  const uint32_t srp = sidx + kIdxCap * 2;                                         // [kTileRows + 1] int32
  const uint32_t sgrow = srp + (kTileRows + 1) * 4;                                // [kGhostRows] int32 global rows
  const uint32_t sgcnt = sgrow + kGhostRows * 4;                                   // ghost counter
  const int c0 = blockIdx.x * 32;
  const int lane8 = threadIdx.x & 7;
  const int sub = threadIdx.x >> 3;
  constexpr int SLOTS = kTileThreads / 8;
  const uint32_t sx_lane = sx + lane8 * 16;
  float4 sc, sh, al;
  if (kTransform) {
    sc = __ldg(reinterpret_cast<const float4*>(scale + c0) + lane8);
    sh = __ldg(reinterpret_cast<const float4*>(shift + c0) + lane8);
    al = __ldg(reinterpret_cast<const float4*>(alpha + c0) + lane8);
  }
  auto f = [&](float4 v) {
    if (kTransform) {
      v.x = bn_prelu(v.x, sc.x, sh.x, al.x);
      v.y = bn_prelu(v.y, sc.y, sh.y, al.y);
      v.z = bn_prelu(v.z, sc.z, sh.z, al.z);
      v.w = bn_prelu(v.w, sc.w, sh.w, al.w);
    }
    return v;
  };
  const float* xg = X + c0;
  const int n_tiles = tile_ptr ? __ldg(n_tiles_dev) : (n_rows + kTileRows - 1) / kTileRows;
  for (int tile = blockIdx.y; tile < n_tiles; tile += gridDim.y) {
    const int r0 = tile_ptr ? __ldg(tile_ptr + tile) : tile * kTileRows;
    const int r1 = tile_ptr ? __ldg(tile_ptr + tile + 1) : min(n_rows, r0 + kTileRows);
    const int n = r1 - r0;
    if (n <= 0) continue;
    for (int k = threadIdx.x; k <= n; k += kTileThreads) sts32(srp + 4 * k, __ldg(rowptr + r0 + k));
    // 1. stage f(X[r0 : r1, slab]): 4 independent 16 B loads in flight per thread, two rounds
    {
      const float* xb = xg + static_cast<int64_t>(r0) * ldx;
#pragma unroll 1
      for (int base = 0; base < kTileRows; base += 4 * SLOTS) {
        if (base >= n) break;
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = base + sub + u * SLOTS;
          if (r < n) v[u] = __ldg(reinterpret_cast<const float4*>(xb + static_cast<int64_t>(r) * ldx) + lane8);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = base + sub + u * SLOTS;
          if (r < n) sts128(sx_lane + r * 128, f(v[u]));
        }
      }
    }
    __syncthreads();
    float* yb = Y + static_cast<int64_t>(r0) * ldy + c0;
    // Rows are processed in passes whose neighbour lists fit the index buffer (one pass for
    // every realistic tile: 512 rows x degree 16).
    int ra = 0;
    while (ra < n) {
      const int e_base = lds32(srp + 4 * ra);
      int rb = n;
      if (lds32(srp + 4 * n) - e_base > kIdxCap) {     // largest rb with nnz[ra, rb) <= cap
        int lo = ra, hi = n;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (lds32(srp + 4 * mid) - e_base <= kIdxCap) lo = mid; else hi = mid - 1;
        }
        rb = lo;
      }
      const bool oversize = rb == ra;                  // one row longer than the buffer
      if (oversize) rb = ra + 1;
      if (threadIdx.x == 0) sts32(sgcnt, 0);
      __syncthreads();
      // 2. neighbour lists -> slot numbers (tile row, ghost slot or overflow)
      if (!oversize) {
        const int cnt = lds32(srp + 4 * rb) - e_base;
        constexpr int PER = kIdxCap / kTileThreads;     // 16 entries per thread, all loads in flight at once
        int jv[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
          const int k = threadIdx.x + u * kTileThreads;
          jv[u] = k < cnt ? __ldg(colidx + e_base + k) : 0;
        }
#pragma unroll
        for (int u = 0; u < PER; ++u) {
          const int k = threadIdx.x + u * kTileThreads;
          if (k < cnt) {
            const int j = jv[u];
            const unsigned jl = static_cast<unsigned>(j - r0);
            int slot;
            if (jl < static_cast<unsigned>(n)) {
              slot = static_cast<int>(jl);
            } else {
              const int g = atoms_add(sgcnt, 1);        // placement only; values do not depend on it
              if (g < kGhostRows) {
                sts32(sgrow + 4 * g, j);
                slot = kTileRows + g;
              } else {
                slot = kOverflow;
              }
            }
            sts16(sidx + 2 * k, slot);
          }
        }
      }
      __syncthreads();
      // 3. ghost rows: global -> f -> shared, all in flight at once
      {
        const int ng = min(lds32(sgcnt), kGhostRows);
        for (int g = sub; g < ng; g += SLOTS) {
          const int j = lds32(sgrow + 4 * g);
          sts128(sx_lane + (kTileRows + g) * 128,
                 f(__ldg(reinterpret_cast<const float4*>(xg + static_cast<int64_t>(j) * ldx) + lane8)));
        }
      }
      __syncthreads();
      // 4. gather
      for (int r = ra + sub; r < rb; r += SLOTS) {
        const int eb = lds32(srp + 4 * r), ee = lds32(srp + 4 * r + 4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        auto add = [&](float4 v) { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; };
        auto add_global = [&](int e) {
          const int j = __ldg(colidx + e);
          const unsigned jl = static_cast<unsigned>(j - r0);
          if (jl < static_cast<unsigned>(n)) add(lds128(sx_lane + jl * 128));
          else add(f(__ldg(reinterpret_cast<const float4*>(xg + static_cast<int64_t>(j) * ldx) + lane8)));
        };
        if (!oversize) {
          int e = eb - e_base;
          const int e_end = ee - e_base;
          for (; e + 4 <= e_end; e += 4) {
            const int s0 = lds16(sidx + 2 * e), s1 = lds16(sidx + 2 * e + 2);
            const int s2 = lds16(sidx + 2 * e + 4), s3 = lds16(sidx + 2 * e + 6);
            if (max(max(s0, s1), max(s2, s3)) == kOverflow) break;     // rare: no ghost slot left
            const float4 v0 = lds128(sx_lane + s0 * 128), v1 = lds128(sx_lane + s1 * 128);
            const float4 v2 = lds128(sx_lane + s2 * 128), v3 = lds128(sx_lane + s3 * 128);
            add(v0); add(v1); add(v2); add(v3);
          }
          for (; e < e_end; ++e) {
            const int sl = lds16(sidx + 2 * e);
            if (sl != kOverflow) add(lds128(sx_lane + sl * 128)); else add_global(e_base + e);
          }
        } else {
          for (int e = eb; e < ee; ++e) add_global(e);
        }
        reinterpret_cast<float4*>(yb + static_cast<int64_t>(r) * ldy)[lane8] = acc;
      }
      ra = rb;
      __syncthreads();                                 // buffers are rewritten by the next pass / tile
    }
  }
}

// Greedy packing of whole graphs into row tiles of at most kTileRows rows; a graph larger than a
// tile is split into equal parts.  Serial over graphs (B is a few thousand), one thread.
__global__ void build_tiles_kernel(const int32_t* __restrict__ graph_ptr, int n_graphs, int n_rows,
                                   int32_t* __restrict__ tile_ptr, int tile_cap, int32_t* __restrict__ n_tiles) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int t = 0, cur = 0;
  tile_ptr[0] = 0;
  for (int g = 0; g < n_graphs; ++g) {
    const int gs = graph_ptr[g], ge = graph_ptr[g + 1], n = ge - gs;
    if (n > kTileRows) {
      if (gs > cur && t < tile_cap) tile_ptr[++t] = gs;
      const int k = (n + kTileRows - 1) / kTileRows;
      for (int i = 1; i <= k && t < tile_cap; ++i) tile_ptr[++t] = gs + static_cast<int>(static_cast<int64_t>(n) * i / k);
      cur = ge;
    } else if (ge - cur > kTileRows) {
      if (t < tile_cap) tile_ptr[++t] = gs;
      cur = gs;
    }
  }
  if (n_rows > cur && t < tile_cap) tile_ptr[++t] = n_rows;
  // if tile_cap was too small (cannot happen with gcs_spmm_tile_capacity) the tail is one tile
  if (tile_ptr[t] != n_rows) tile_ptr[t] = n_rows;
  *n_tiles = t;
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
    // a CTA walks a CONTIGUOUS chunk of rows: consecutive rows of a banded matrix share most of
    // their neighbours, which then hit in L1 instead of going back to L2
    const int64_t chunk = static_cast<int64_t>(rows_per_block) * min_rows;   // min_rows = iterations per CTA here
    const int64_t rbeg = static_cast<int64_t>(blockIdx.x) * chunk;
    const int64_t rend = rbeg + chunk < n_rows ? rbeg + chunk : n_rows;
    for (int64_t r = rbeg + slot; r < rend; r += rows_per_block) do_row(r);
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
  const int32_t* rowptr; const int32_t* colidx; const int32_t* tile_ptr; const int32_t* n_tiles_dev; int64_t n_rows;
  const float* X; int64_t ldx; const float* scale; const float* shift; const float* alpha;
  float* Y; int64_t ldy; int H; cudaStream_t st;
};

int g_rows_iters = 16;   // tuning knob (gcs_debug_set_param 1)

template <int VEC>
int launch_rows(const SpmmArgs& a, const int32_t* graph_ptr, int min_rows) {
  const int lanes = (a.H + VEC - 1) / VEC;
  if (lanes > 256) return fail(GCS_ERR_UNSUPPORTED, "gcs_spmm_sum: H=%d too wide for the row kernel", a.H);
  const int rows_per_block = 256 / lanes;
  dim3 grid;
  int iters = g_rows_iters;                         // rows_per_block * iters contiguous rows per CTA
  if (graph_ptr == nullptr) {
    grid = dim3(static_cast<unsigned>(ceil_div(a.n_rows, static_cast<int64_t>(rows_per_block) * iters)));
    min_rows = iters;
  } else {
    grid = dim3(8, 1);
  }
  if (a.scale)
    spmm_rows_kernel<VEC, true><<<grid, 256, 0, a.st>>>(a.rowptr, a.colidx, a.n_rows, a.X, a.ldx, a.scale, a.shift, a.alpha, a.Y, a.ldy, a.H, lanes, rows_per_block, graph_ptr, 0, min_rows);
  else
    spmm_rows_kernel<VEC, false><<<grid, 256, 0, a.st>>>(a.rowptr, a.colidx, a.n_rows, a.X, a.ldx, a.scale, a.shift, a.alpha, a.Y, a.ldy, a.H, lanes, rows_per_block, graph_ptr, 0, min_rows);
  GCS_CHECK_LAUNCH("spmm_rows_kernel");
  return GCS_OK;
}

int launch_tile(const SpmmArgs& a) {
  auto kt = spmm_tile_kernel<true>;
  auto kf = spmm_tile_kernel<false>;
  static bool attr_set = false;
  if (!attr_set) {
    GCS_CUDA(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
    GCS_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
    attr_set = true;
  }
  // graph-aligned tiles are at least half full on average: ceil(N/T) * 2 CTAs cover them,
  // any remainder is picked up by the tile loop inside the kernel
  int64_t tiles = ceil_div(a.n_rows, kTileRows) * (a.tile_ptr ? 2 : 1);
  if (tiles > 65535) tiles = 65535;
  dim3 grid(a.H / 32, static_cast<unsigned>(tiles));
  if (a.scale)
    kt<<<grid, kTileThreads, kTileSmemBytes, a.st>>>(a.rowptr, a.colidx, static_cast<int>(a.n_rows), a.tile_ptr, a.n_tiles_dev, a.X, a.ldx, a.scale, a.shift, a.alpha, a.Y, a.ldy);
  else
    kf<<<grid, kTileThreads, kTileSmemBytes, a.st>>>(a.rowptr, a.colidx, static_cast<int>(a.n_rows), a.tile_ptr, a.n_tiles_dev, a.X, a.ldx, a.scale, a.shift, a.alpha, a.Y, a.ldy);
  GCS_CHECK_LAUNCH("spmm_tile_kernel");
  return GCS_OK;
}

int g_spmm_mode = 0;   // 0 = auto, 1 = row kernel, 2 = tile kernel

}  // namespace

// Test / sweep hook (not part of the drop-in surface).
extern "C" void gcs_debug_set_spmm_mode(int mode) { g_spmm_mode = mode; }
namespace gcs { namespace tc { void set_wgrad_chain(int c); } }
extern "C" void gcs_debug_set_param(int id, int value) {
  if (id == 1 && value > 0) g_rows_iters = value;
  if (id == 3) gcs::tc::set_wgrad_chain(value);
}

extern "C" int32_t gcs_spmm_tile_capacity(int64_t n_rows, int32_t n_graphs) {
  return static_cast<int32_t>(2LL * n_graphs + ceil_div(n_rows, kTileRows) + 2);
}

extern "C" int gcs_spmm_build_tiles(const int32_t* graph_ptr, int32_t n_graphs, int64_t n_rows, int32_t* tile_ptr,
                                    int32_t tile_capacity, int32_t* n_tiles_dev, gcs_stream stream) {
  GCS_CHECK_ARG(graph_ptr && tile_ptr && n_tiles_dev && n_graphs >= 0 && n_rows >= 0 && n_rows < INT32_MAX,
                "gcs_spmm_build_tiles: bad argument");
  GCS_CHECK_ARG(tile_capacity >= gcs_spmm_tile_capacity(n_rows, n_graphs), "gcs_spmm_build_tiles: tile_ptr too small (need %d + 1 entries)",
                gcs_spmm_tile_capacity(n_rows, n_graphs));
  build_tiles_kernel<<<1, 32, 0, as_stream(stream)>>>(graph_ptr, n_graphs, static_cast<int>(n_rows), tile_ptr, tile_capacity, n_tiles_dev);
  GCS_CHECK_LAUNCH("build_tiles_kernel");
  return GCS_OK;
}

extern "C" int gcs_spmm_sum(const int32_t* rowptr, const int32_t* colidx, const int32_t* tile_ptr,
                            const int32_t* n_tiles_dev, int64_t n_rows, const float* X, int64_t ldx,
                            const float* scale, const float* shift, const float* alpha, float* Y,
                            int64_t ldy, int32_t H, gcs_stream stream) {
  GCS_CHECK_ARG(n_rows >= 0 && H > 0, "gcs_spmm_sum: bad size (n_rows=%lld, H=%d)", (long long)n_rows, H);
  if (n_rows == 0) return GCS_OK;
  GCS_CHECK_ARG(rowptr && colidx && X && Y, "gcs_spmm_sum: null pointer");
  GCS_CHECK_ARG(ldx >= H && ldy >= H, "gcs_spmm_sum: leading dimension smaller than H");
  GCS_CHECK_ARG((scale != nullptr) == (shift != nullptr) && (scale != nullptr) == (alpha != nullptr),
                "gcs_spmm_sum: scale/shift/alpha must be all NULL or all set");
  GCS_CHECK_ARG((tile_ptr != nullptr) == (n_tiles_dev != nullptr), "gcs_spmm_sum: tile_ptr and n_tiles_dev go together");
  GCS_CHECK_ARG(X != Y, "gcs_spmm_sum: in-place aggregation is not defined");
  GCS_CHECK_ARG(n_rows < INT32_MAX, "gcs_spmm_sum: n_rows exceeds int32 CSR range");
  SpmmArgs a{rowptr, colidx, tile_ptr, n_tiles_dev, n_rows, X, ldx, scale, shift, alpha, Y, ldy, H, as_stream(stream)};
  const bool vec_ok = (H % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(X) && aligned16(Y) &&
                      (!scale || (aligned16(scale) && aligned16(shift) && aligned16(alpha)));
  // auto: the L1-served row kernel currently beats the tile kernel on B200 (profiles/); the
  // tile kernel stays selectable (mode 2) and tested
  if (g_spmm_mode == 2 && vec_ok && H % 32 == 0) return launch_tile(a);
  if (vec_ok) return launch_rows<4>(a, nullptr, 0);
  return launch_rows<1>(a, nullptr, 0);
}
