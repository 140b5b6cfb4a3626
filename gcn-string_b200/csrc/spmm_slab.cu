// K3 / K7 for disjoint batches — Y = pattern(A) . f(X) with every graph staged in shared memory.
//
// A disjoint batch (spektral.data.DisjointLoader, src/scripts/gcn.py:316-317) is block-diagonal: every neighbour of a
// node lies in the node's own graph.  One work item = (graph, 32 feature columns): that slice of
// the graph's rows of X fits a shared-memory slab, so
//   * every element of X leaves HBM exactly once, in 128-byte runs, by the TMA unit (2-D tensor copies of 32 rows);
//   * the BatchNorm + PReLU prologue of GeneralConv (transform -> BN -> PReLU -> aggregate) is applied ONCE per
//     element, in place, instead of once per gathered neighbour;
//   * all gathers of the graph are 128-bit shared-memory loads: no L2 / DRAM latency inside the gather chain.
// The kernel is persistent (one CTA per SM, 1024 threads) and warp-specialised, with a ring of 2-3 slabs:
//   warp 0         producer: walks this CTA's work items, waits for a free slab, posts the item's descriptor and issues
//                  the bulk copies (X slice, the graph's row-block entries) that complete on the slab's `landed` mbarrier;
//   warps 1-4      prologue: wait `landed`, rewrite the entry words into ready-made shared-memory addresses, apply
//                  f(x) = prelu(x*scale + shift, alpha) in place, arrive on `ready`;
//   warps 5-31     gather: wait `ready`, take row blocks from a shared counter (4 blocks per warp at a time), sum each
//                  block's neighbours out of the slab, store the rows of Y, arrive on `empty`.
// so the HBM-bound staging of item i+1/i+2 overlaps the shared-memory-bound gather of item i and no warp ever waits
// at a CTA-wide barrier.  Row blocks use the RB format of spmm.cu (height 2 or 4, blocks padded to 4 entries): a
// neighbour row is read once per block, each lane owns 4 columns and adds with packed fp32x2 instructions; per output
// row the summation order is ascending column, bit-identical to the CSR row kernel.  Row blocks are aligned to batch
// rows, so the first / last block of a graph may straddle its neighbour: entries outside the graph's column range
// become no-ops when the prologue warps rewrite them.  Graphs too long for a 32-column slab are walked in passes of
// 16 / 8 / 4 columns; a graph whose entries do not fit even then is gathered straight from global memory by the same
// warps (slow, correct).
#include <cuda.h>

#include <atomic>

#include "common.cuh"

#ifndef GCS_SLAB_ASM
#define GCS_SLAB_ASM volatile
#endif
#ifndef GCS_SLAB_HELPERS
#define GCS_SLAB_HELPERS 0
#endif
#ifndef GCS_SLAB_ADD_MODE
#define GCS_SLAB_ADD_MODE 2   // 1 = FADD2 + 2 FADD per float4 (307 us at cfg2), 2 = 2 FADD2 (300 us)
#endif

namespace gcs {
int spmm_mode();   // spmm.cu: gcs_debug_set_spmm_mode

namespace slab {

constexpr int kCols = 32;            // feature columns per work item
constexpr int kThreads = 1024;
#ifndef GCS_SLAB_XFORM_WARPS
#define GCS_SLAB_XFORM_WARPS 8   // measured at cfg2 (prologue / plain, us): 2 warps 318 / 268, 4: 289 / 260, 6: 289 / 256, 8: 284 / 252
#endif
constexpr int kXformWarps = GCS_SLAB_XFORM_WARPS;
constexpr int kGatherWarps = 31 - kXformWarps;
constexpr int kMaxStages = 3;
constexpr int kHeaderBytes = 1024;
constexpr int kBoxRows = 32;         // rows per TMA box: a graph's slab holds its row count rounded up to this
constexpr uint32_t kAddrMask = 0x00FFFFF0u;

struct Meta {
  int mode;        // 0 = slab, 1 = direct (global-memory gather), -1 = stop
  int off, n;      // the graph's first row and row count
  int n_up;        // n rounded up to kBoxRows: rows the slab holds (the tail belongs to the next graph / is zero fill)
  int lq;          // log2(lanes per row block) of this pass: cw = 4 << lq columns
  int col0;        // first feature column of the pass
  int b_first, nb; // the graph's row blocks
  int e0;          // index of the graph's first entry in the global entry array
  int ent_words;
  int blk_skip;    // the staged block pointers start blk_skip words before the graph's first block (16-byte aligned copy)
  int blk_words;   // staged block-pointer words (a multiple of 4)
};
constexpr int kBigBoxRows = 128;     // the bulk of a graph is fetched in boxes of 128 rows, the tail in boxes of kBoxRows
struct Maps { CUtensorMap m[2][4]; };   // X as a 2-D tensor; boxes of kBoxRows ([0]) / kBigBoxRows ([1]) rows x (4 << k) columns

struct Header {
  unsigned long long landed[kMaxStages], ready[kMaxStages], empty[kMaxStages];
  Meta meta[kMaxStages];
  int counter[kMaxStages];
};
static_assert(sizeof(Header) <= kHeaderBytes, "header too large");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Waiting warps share their scheduler with the warps that gather: the suspend-time hint parks the thread in hardware
// until the phase completes instead of re-issuing the test (the spin loop was 19 % of all issued instructions).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// TMA bulk copy global -> shared (16-byte aligned, size a multiple of 16), completing on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;   // GCS_SLAB_ASM: volatile or empty; an address always depends on words read after the stage's barrier
  asm GCS_SLAB_ASM("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm GCS_SLAB_ASM("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// acc += v as packed fp32x2 adds (round-to-nearest per half: the bits of four scalar adds in half the issue slots).
// FADD2 runs on the fma-heavy pipe only (2 cycles per warp); splitting the float4 add into one FADD2 plus two scalar
// FADDs (mode 1) was measured marginally slower.
__device__ __forceinline__ void add4(float4& acc, const float4& v) {
  asm("{\n\t.reg .b64 a, b;\n\tmov.b64 a, {%0, %1};\n\tmov.b64 b, {%2, %3};\n\tadd.rn.f32x2 a, a, b;\n\tmov.b64 {%0, %1}, a;\n\t}"
      : "+f"(acc.x), "+f"(acc.y) : "f"(v.x), "f"(v.y));
#if GCS_SLAB_ADD_MODE == 2
  asm("{\n\t.reg .b64 a, b;\n\tmov.b64 a, {%0, %1};\n\tmov.b64 b, {%2, %3};\n\tadd.rn.f32x2 a, a, b;\n\tmov.b64 {%0, %1}, a;\n\t}"
      : "+f"(acc.z), "+f"(acc.w) : "f"(v.z), "f"(v.w));
#else
  acc.z += v.z;
  acc.w += v.w;
#endif
}

// Slab entry word: bits [23:4] = shared-memory address of the neighbour's row of the slab in 16-byte units (the lane ORs
// its own column offset in; slab rows start on 128-byte boundaries), bit 31 - r = row r of the block has this neighbour.
// No bit set = padding, or an entry of the neighbouring graph in a straddling block: a no-op that reads slab row 0.
__device__ __forceinline__ uint32_t slab_word(uint32_t raw, int off, int n, int shift, uint32_t base) {
  const int loc = static_cast<int>(raw >> 8) - off;
  const uint32_t m = __brev(raw) & 0xF0000000u;
  return static_cast<unsigned>(loc) < static_cast<unsigned>(n) ? (base + (static_cast<uint32_t>(loc) << shift)) | m : base;
}

template <int RB>
__device__ __forceinline__ void scatter(float4 (&acc)[RB], const float4& v, uint32_t m) {
#pragma unroll
  for (int r = 0; r < RB; ++r)
    if (r == 0 ? static_cast<int>(m) < 0 : (m & (0x80000000u >> r)) != 0u) add4(acc[r], v);
}

struct Out {
  const float* R; int64_t ldr; float* Y; int64_t ldy; bool want_amax;
};

// Rows b*RB .. b*RB+RB-1 of one column quad: + residual (Add()([z, out]) of connectivity='sum'), store, |max|.
template <int RB, bool kResidual>
__device__ __forceinline__ float store_block(const Out& o, float4 (&acc)[RB], int b, int off, int end, int col, float mx) {
  const int row0 = b * RB;
  float* yrow = o.Y + static_cast<int64_t>(row0) * o.ldy + col;
  const float* rrow = kResidual ? o.R + static_cast<int64_t>(row0) * o.ldr + col : nullptr;
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    if (row0 + r >= off && row0 + r < end) {
      if (kResidual) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(rrow));
        acc[r].x += t.x; acc[r].y += t.y; acc[r].z += t.z; acc[r].w += t.w;
      }
      *reinterpret_cast<float4*>(yrow) = acc[r];
      if (o.want_amax) mx = amax4(mx, acc[r]);
    }
    yrow += o.ldy;
    if (kResidual) rrow += o.ldr;
  }
  return mx;
}

// One gather pass of a warp over a staged graph: 2^LQ lanes per row block, each lane owns 4 columns; row blocks are
// handed out through the stage's shared counter, (32 >> LQ) consecutive blocks per warp at a time.
// One round = (32 >> LQ) consecutive row blocks from the stage's counter.  A prologue warp that helps gathering passes the
// `landed` barrier of the next stage (ybar != 0): once that stage has landed its prologue comes first.
__device__ __forceinline__ int claim_round(int* counter, int gpw, uint32_t ybar, uint32_t ypar) {
  int base = 0x7fffffff;
  if ((threadIdx.x & 31) == 0 && !(ybar && mbar_test(ybar, ypar))) base = atomicAdd(counter, gpw);
  return __shfl_sync(0xffffffffu, base, 0);
}

template <int RB, int LQ, bool kResidual>
__device__ __forceinline__ float gather_slab(const Meta& m, uint32_t slab, uint32_t sblk, uint32_t sent, int* counter,
                                             const Out& o, float mx, uint32_t ybar, uint32_t ypar) {
  constexpr int q = 1 << LQ, gpw = 32 >> LQ;
  const int lane = threadIdx.x & 31;
  const int quad = lane & (q - 1), gidx = lane >> LQ;
  const uint32_t qoff = static_cast<uint32_t>(quad) << 4;
  const int col = m.col0 + quad * 4;
  const int end = m.off + m.n;
  for (;;) {
    // (Claiming the next round ahead of time to hide the atomic's latency was measured slower: a warp then sits on
    // blocks that an idle warp could have taken.)
    const int base = claim_round(counter, gpw, ybar, ypar);
    if (base >= m.nb) break;
    // (Handing the blocks out longest first, sorted by the prologue warps so that the blocks of a round run the same
    // number of iterations, was measured: no gain - 281 / 254 us against 279 / 249 us.)
    const int bl = base + gidx;
    if (bl < m.nb) {
      float4 acc[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      int e0, e1;
      asm volatile("ld.shared.s32 %0, [%2];\n\tld.shared.s32 %1, [%2 + 4];" : "=&r"(e0), "=&r"(e1) : "r"(sblk + 4u * (bl + m.blk_skip)));
      e0 -= m.e0;                                     // the staged pointers are the global ones
      e1 -= m.e0;
      uint32_t ea = sent + 4u * e0;
      const uint32_t eb = sent + 4u * e1;
      // (The RB list leads with groups of neighbours that all rows of the block share, flagged for a shared accumulator:
      // spmm_rb4_kernel uses it - it is bound by issue slots.  Here it was measured and brings nothing, 386 / 354 us
      // against 380 / 349 us at degree 32 with 40 % fewer adds: this kernel's time follows its shared-memory wavefronts,
      // 46 M at degree 12, 67 M at degree 32.  The summation order is that of the list either way, so both kernels agree
      // bit for bit.)
      for (; ea < eb; ea += 16) {
        const uint4 w = lds128u(ea);
        // (no-op words - padding, the neighbouring graph's entries of a straddling block - still gather slab row 0:
        // predicating those loads off saves ~6 % of the gather wavefronts and was measured 10 % SLOWER, 290 / 261 us)
        const float4 v0 = lds128((w.x & kAddrMask) | qoff), v1 = lds128((w.y & kAddrMask) | qoff);
        const float4 v2 = lds128((w.z & kAddrMask) | qoff), v3 = lds128((w.w & kAddrMask) | qoff);
        scatter<RB>(acc, v0, w.x); scatter<RB>(acc, v1, w.y); scatter<RB>(acc, v2, w.z); scatter<RB>(acc, v3, w.w);
      }
      mx = store_block<RB, kResidual>(o, acc, m.b_first + bl, m.off, end, col, mx);
    }
  }
  return mx;
}

// The same pass without a slab: entries, block pointers and X rows straight from global memory (a graph whose
// row-block entries do not fit a stage).  Each gather applies the prologue itself.
template <int RB, int LQ, bool kTransform>
__device__ __forceinline__ float gather_direct(const Meta& m, const int32_t* __restrict__ blk_ptr, const uint32_t* __restrict__ ent,
                                               const float* __restrict__ X, int64_t ldx, const float* __restrict__ scale,
                                               const float* __restrict__ shift, const float* __restrict__ alpha, int* counter,
                                               const Out& o, float mx, uint32_t ybar, uint32_t ypar) {
  constexpr int q = 1 << LQ, gpw = 32 >> LQ;
  const int lane = threadIdx.x & 31;
  const int quad = lane & (q - 1), gidx = lane >> LQ;
  const int col = m.col0 + quad * 4;
  const int end = m.off + m.n;
  float4 sc, sh, al;
  if (kTransform) {
    sc = __ldg(reinterpret_cast<const float4*>(scale + col));
    sh = __ldg(reinterpret_cast<const float4*>(shift + col));
    al = __ldg(reinterpret_cast<const float4*>(alpha + col));
  }
  for (;;) {
    const int base = claim_round(counter, gpw, ybar, ypar);
    if (base >= m.nb) break;
    const int bl = base + gidx;
    if (bl < m.nb) {
      const int b = m.b_first + bl;
      float4 acc[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int e0 = __ldg(blk_ptr + b), e1 = __ldg(blk_ptr + b + 1);
      for (int e = e0; e < e1; ++e) {
        const uint32_t raw = __ldg(ent + e);
        const int c = static_cast<int>(raw >> 8);
        if (c < m.off || c >= end) continue;          // the neighbouring graph's part of a straddling block
        float4 v = __ldg(reinterpret_cast<const float4*>(X + static_cast<int64_t>(c) * ldx + col));
        if (kTransform) {
          v.x = bn_prelu(v.x, sc.x, sh.x, al.x);
          v.y = bn_prelu(v.y, sc.y, sh.y, al.y);
          v.z = bn_prelu(v.z, sc.z, sh.z, al.z);
          v.w = bn_prelu(v.w, sc.w, sh.w, al.w);
        }
        scatter<RB>(acc, v, __brev(raw) & 0xF0000000u);
      }
      mx = o.R ? store_block<RB, true>(o, acc, b, m.off, end, col, mx) : store_block<RB, false>(o, acc, b, m.off, end, col, mx);
    }
  }
  return mx;
}

template <int RB, bool kTransform>
__global__ void __launch_bounds__(kThreads, 1) spmm_slab_kernel(
    const __grid_constant__ Maps maps, const int32_t* __restrict__ graph_ptr, int n_graphs,
    const int32_t* __restrict__ blk_ptr, const uint32_t* __restrict__ ent, const float* __restrict__ X, int64_t ldx, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ alpha, const float* __restrict__ R, int64_t ldr, float* __restrict__ Y, int64_t ldy, int H,
    int stage_bytes, int n_stages, float* __restrict__ amax, unsigned int* __restrict__ ticket,
    unsigned long long* __restrict__ dbg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // dbg (gcs_debug_slab_timing): per role {cycles waiting on its barrier, cycles working}, summed over warps and CTAs
#ifdef GCS_SLAB_TIMING                               // development build only (scripts/slab_timing.py): costs registers
  unsigned long long t_wait = 0, t_work = 0, t0 = 0;
#define GCS_TIC() do { if (dbg) t0 = clock64(); } while (0)
#define GCS_TOC(acc) do { if (dbg) { const unsigned long long t1 = clock64(); acc += t1 - t0; t0 = t1; } } while (0)
#define GCS_TREPORT(i, j) do { if (dbg && lane == 0) { atomicAdd(dbg + i, t_wait); atomicAdd(dbg + j, t_work); } } while (0)
#else
#define GCS_TIC() do { } while (0)
#define GCS_TOC(acc) do { } while (0)
#define GCS_TREPORT(i, j) do { } while (0)
#endif
  // 128-byte aligned base inside the shared window: a lane ORs its column offset into the entry words
  const uint32_t raw_base = smem_u32(smem_raw);
  unsigned char* const smem = smem_raw + ((128u - (raw_base & 127u)) & 127u);
  Header* const hd = reinterpret_cast<Header*>(smem);
  unsigned char* const stage0 = smem + kHeaderBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = n_stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&hd->landed[s]), 1);
      mbar_init(smem_u32(&hd->ready[s]), kXformWarps);
      mbar_init(smem_u32(&hd->empty[s]), (GCS_SLAB_HELPERS ? kXformWarps : 0) + kGatherWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int ncg = (H + kCols - 1) / kCols;           // the column groups of one graph are neighbouring work items:
  const int64_t n_items = static_cast<int64_t>(n_graphs) * ncg;   // CTAs read whole rows of X together

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    // The scalars of an item hang off two dependent global loads (graph_ptr -> blk_ptr); they are requested one and
    // two items ahead so that posting an item never waits for them.
    // Work items are handed out through a ticket counter in global memory (ticket[0]; ticket[1] counts the CTAs that
    // have left the queue, the last one re-arms both for the next launch).  Graphs differ in length (log-normal protein
    // lengths: the row count of an item varies by ~26 %), so with a fixed round-robin share the slowest of the 148
    // CTAs carried ~9 % more rows than the average and every other SM idled at the end of the launch.  Which CTA
    // runs an item changes nothing in the result.  A ticket is drawn three items ahead of its use and read one item
    // after it was drawn, so the atomic's round trip is never waited for.
    // Tickets run from the LAST graph to the first: the producer of X (the dense transform in front of the
    // aggregation) wrote its rows in ascending order, so the tail of X is what the 126 MB L2 still holds.
    auto graph_of = [&](int64_t w, int& off, int& n) {
      off = 0; n = 0;
      if (w < n_items) {
        const int g = static_cast<int>((n_items - 1 - w) / ncg);
        off = __ldg(graph_ptr + g);
        n = __ldg(graph_ptr + g + 1) - off;
      }
    };
    auto entries_of = [&](int off, int n, int& E0, int& E1) {
      E0 = 0; E1 = 0;
      if (n > 0) {
        const int b_first = off / RB, nb = (off + n - 1) / RB - b_first + 1;
        E0 = __ldg(blk_ptr + b_first);
        E1 = __ldg(blk_ptr + b_first + nb);
      }
    };
    int off, n, E0, E1, off1, n1, E0n, E1n, off2, n2;
    unsigned int tk = 0;                             // lane 0: the ticket in flight
    if (lane == 0) tk = atomicAdd(ticket, 3u);       // the first three items of a CTA are neighbours
    int64_t w = static_cast<int64_t>(__shfl_sync(0xffffffffu, tk, 0)), w1 = w + 1, w2 = w + 2;
    if (lane == 0) tk = atomicAdd(ticket, 1u);
    graph_of(w, off, n);
    entries_of(off, n, E0, E1);
    graph_of(w1, off1, n1);
    int it = 0;
    for (; w < n_items;) {
      graph_of(w2, off2, n2);                        // consumed two iterations from now
      entries_of(off1, n1, E0n, E1n);                // consumed next iteration
      if (n > 0) {
        const int64_t wr = n_items - 1 - w;
        const int g = static_cast<int>(wr / ncg);
        const int cg0 = static_cast<int>(wr - static_cast<int64_t>(g) * ncg) * kCols;
        const int cgw = min(kCols, H - cg0);
        const int b_first = off / RB, nb = (off + n - 1) / RB - b_first + 1;
        const int ent_words = E1 - E0;                // a multiple of 4 (padded blocks)
        const int blk_skip = b_first & 3;             // the block pointers are copied from a 16-byte boundary
        const int blk_words = (blk_skip + nb + 1 + 3) & ~3;
        const int fixed = 4 * (blk_words + ent_words);
        const int nbox = (n + kBoxRows - 1) / kBoxRows, n_up = nbox * kBoxRows;
        int cfit = kCols;
        while (cfit >= 4 && static_cast<int64_t>(n_up) * cfit * 4 + fixed > stage_bytes) cfit >>= 1;
        const bool direct = cfit < 4;
        for (int c0 = 0; c0 < cgw;) {
          int cw = direct ? kCols : cfit;
          while (cw > cgw - c0) cw >>= 1;             // power of two >= 4 (H % 4 == 0)
          const int lq = 31 - __clz(cw >> 2);
          const int s = it % S;
          unsigned char* const st = stage0 + static_cast<size_t>(s) * stage_bytes;
          const int slab_bytes = n_up * cw * 4;
          const uint32_t bar = smem_u32(&hd->landed[s]);
          // boxes: lanes [0, nbig) fetch 128 rows each, the next nsmall lanes 32 rows each (rounds of 32 lanes)
          const int nbig = n_up / kBigBoxRows, nsmall = (n_up - nbig * kBigBoxRows) / kBoxRows;
          const uint32_t row_bytes = cw * 4;
          GCS_TIC();
          if (it >= S) mbar_wait(smem_u32(&hd->empty[s]), ((it / S) - 1) & 1);
          GCS_TOC(t_wait);
          if (lane == 0) {
            Meta& m = hd->meta[s];
            m.mode = direct ? 1 : 0; m.off = off; m.n = n; m.n_up = n_up; m.lq = lq; m.col0 = cg0 + c0;
            m.b_first = b_first; m.nb = nb; m.e0 = E0; m.ent_words = ent_words; m.blk_skip = blk_skip; m.blk_words = blk_words;
            hd->counter[s] = 0;
            fence_proxy_async();                      // the slab was read / written through the generic proxy before
            if (direct) {
              mbar_arrive(bar);
            } else {
              mbar_arrive_expect_tx(bar, static_cast<uint32_t>(slab_bytes + fixed));
              bulk_g2s(smem_u32(st + slab_bytes), blk_ptr + (b_first - blk_skip), 4u * blk_words, bar);
              if (ent_words > 0) bulk_g2s(smem_u32(st + slab_bytes + 4 * blk_words), ent + E0, 4u * ent_words, bar);
            }
          }
          __syncwarp();
          if (!direct) {
            for (int bx = lane; bx < nbig + nsmall; bx += 32) {
              const bool big = bx < nbig;
              const int row = big ? bx * kBigBoxRows : nbig * kBigBoxRows + (bx - nbig) * kBoxRows;
              tma_load_2d(smem_u32(st) + row * row_bytes, &maps.m[big ? 1 : 0][lq], cg0 + c0, off + row, bar);
            }
          }
          GCS_TOC(t_work);
          ++it;
          c0 += cw;
        }
      }
      off = off1; n = n1; E0 = E0n; E1 = E1n; off1 = off2; n1 = n2;
      w = w1; w1 = w2;
      w2 = static_cast<int64_t>(__shfl_sync(0xffffffffu, tk, 0));   // drawn one item ago: long since returned
      if (lane == 0) tk = atomicAdd(ticket, 1u);
    }
    w2 = static_cast<int64_t>(__shfl_sync(0xffffffffu, tk, 0));     // the last draw has returned before the CTA leaves the queue
    GCS_TREPORT(0, 1);
    if (dbg && lane == 0) atomicAdd(dbg + 6, 1ull * it);
    const int s = it % S;                             // stop marker
    if (it >= S) mbar_wait(smem_u32(&hd->empty[s]), ((it / S) - 1) & 1);
    if (lane == 0) {
      hd->meta[s].mode = -1;
      mbar_arrive(smem_u32(&hd->landed[s]));
      // the last CTA to leave re-arms the counters for the next launch that is given this slot
      __threadfence();
      if (w2 >= 0 && atomicAdd(ticket + 1, 1u) == gridDim.x - 1) {
        ticket[0] = 0u;
        ticket[1] = 0u;
        __threadfence();
      }
    }
  } else {
    // ------------------------------------------------------------------ consumers
    // Warps 1..kXformWarps run the prologue of every stage (entry words -> shared addresses, f(x) in place); the rest
    // gather.  (GCS_SLAB_HELPERS=1 lets the prologue warps gather as well until the next stage lands - measured slower,
    // 306 / 273 us against 279 / 249 us: they only look for the next stage between rounds and the prologue starts late.)
    const bool helper = warp <= kXformWarps;
    const int tt = (warp - 1) * 32 + lane;
    constexpr int NT = kXformWarps * 32;
    const Out o{R, ldr, Y, ldy, amax != nullptr};
    float mx = 0.f;
    for (int it = 0;; ++it) {
      const int s = it % S;
      const uint32_t par = (it / S) & 1;
      unsigned char* const st = stage0 + static_cast<size_t>(s) * stage_bytes;
      if (helper) {
        GCS_TIC();
        mbar_wait(smem_u32(&hd->landed[s]), par);
        GCS_TOC(t_wait);
        const Meta m = hd->meta[s];
        if (m.mode == 0) {
          const int slab_bytes = (m.n_up << m.lq) * 16;
          const uint32_t slab = smem_u32(st);
          uint4* const sent4 = reinterpret_cast<uint4*>(st + slab_bytes + 4 * m.blk_words);
          const int word_shift = m.lq + 4;
#pragma unroll 2
          for (int c = tt; c < (m.ent_words >> 2); c += NT) {
            uint4 w = sent4[c];
            w.x = slab_word(w.x, m.off, m.n, word_shift, slab);
            w.y = slab_word(w.y, m.off, m.n, word_shift, slab);
            w.z = slab_word(w.z, m.off, m.n, word_shift, slab);
            w.w = slab_word(w.w, m.off, m.n, word_shift, slab);
            sent4[c] = w;
          }
          if (kTransform) {
            const int col = m.col0 + (tt & ((1 << m.lq) - 1)) * 4;    // constant per thread: NT is a multiple of q
            const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + col));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + col));
            const float4 al = __ldg(reinterpret_cast<const float4*>(alpha + col));
            float4* d = reinterpret_cast<float4*>(st) + tt;
            float4* const dend = reinterpret_cast<float4*>(st) + (m.n << m.lq);
#pragma unroll 4
            for (; d < dend; d += NT) {
              float4 v = *d;
              v.x = bn_prelu(v.x, sc.x, sh.x, al.x);
              v.y = bn_prelu(v.y, sc.y, sh.y, al.y);
              v.z = bn_prelu(v.z, sc.z, sh.z, al.z);
              v.w = bn_prelu(v.w, sc.w, sh.w, al.w);
              *d = v;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&hd->ready[s]));
        GCS_TOC(t_work);
        if (m.mode < 0) break;
        if (!GCS_SLAB_HELPERS) continue;
      }
      GCS_TIC();
      mbar_wait(smem_u32(&hd->ready[s]), par);
      GCS_TOC(t_wait);
      const Meta m = hd->meta[s];
      if (m.mode < 0) break;
      const uint32_t slab = smem_u32(st);
      const uint32_t sblk = slab + static_cast<uint32_t>(m.n_up << m.lq) * 16u;
      const uint32_t sent = sblk + 4u * m.blk_words;
      int* const ctr = &hd->counter[s];
      const uint32_t ybar = helper ? smem_u32(&hd->landed[(it + 1) % S]) : 0u;
      const uint32_t ypar = ((it + 1) / S) & 1;
#define GCS_GATHER_SLAB(LQ)                                                                          \
  mx = R ? gather_slab<RB, LQ, true>(m, slab, sblk, sent, ctr, o, mx, ybar, ypar)                    \
         : gather_slab<RB, LQ, false>(m, slab, sblk, sent, ctr, o, mx, ybar, ypar)
#define GCS_GATHER_DIRECT(LQ) \
  mx = gather_direct<RB, LQ, kTransform>(m, blk_ptr, ent, X, ldx, scale, shift, alpha, ctr, o, mx, ybar, ypar)
      if (m.mode == 0) {
        switch (m.lq) {
          case 3: GCS_GATHER_SLAB(3); break;
          case 2: GCS_GATHER_SLAB(2); break;
          case 1: GCS_GATHER_SLAB(1); break;
          default: GCS_GATHER_SLAB(0); break;
        }
      } else {
        switch (m.lq) {
          case 3: GCS_GATHER_DIRECT(3); break;
          case 2: GCS_GATHER_DIRECT(2); break;
          case 1: GCS_GATHER_DIRECT(1); break;
          default: GCS_GATHER_DIRECT(0); break;
        }
      }
#undef GCS_GATHER_SLAB
#undef GCS_GATHER_DIRECT
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&hd->empty[s]));
      GCS_TOC(t_work);
    }
    GCS_TREPORT(4, 5);
    amax_commit(mx, amax);
  }
}

int g_stages = 2;          // gcs_debug_set_param 10: 2 stages of 113 KB (a 500-node graph runs 32 columns wide) beat 3 of 75 KB at cfg2
int g_stage_bytes = 0;     // gcs_debug_set_param 11 (0 = as large as the stages allow)
int g_grid = 0;            // gcs_debug_set_param 12 (0 = one CTA per SM)
unsigned long long* g_dbg = nullptr;   // gcs_debug_slab_timing: 8 device counters, see the kernel

constexpr int kSmemMax = 232448;   // opt-in maximum per CTA on sm_100

// Work-queue counters of the persistent kernel: {next ticket, CTAs that left} per slot, zero at module load and re-armed
// by the last CTA of every launch.  Launches take the slots round-robin, so two launches only share a slot when more
// than kTicketSlots aggregation kernels of one device are in flight at once (each occupies every SM).
constexpr int kTicketSlots = 64;
__device__ unsigned int g_tickets[2 * kTicketSlots];

int stage_bytes() {
  const int most = ((kSmemMax - kHeaderBytes - 128) / g_stages) & ~127;
  return g_stage_bytes > 0 && g_stage_bytes < most ? (g_stage_bytes & ~127) : most;
}

}  // namespace slab

void slab_set_param(int id, int value) {
  if (id == 10 && value >= 2 && value <= slab::kMaxStages) slab::g_stages = value;
  if (id == 11 && value >= 0) slab::g_stage_bytes = value;
  if (id == 12 && value >= 0) slab::g_grid = value;
}

}  // namespace gcs

using namespace gcs;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

// X [n_rows, H] (row pitch ldx) as 2-D tensors with boxes of kBoxRows / kBigBoxRows rows x 4 / 8 / 16 / 32 columns, rows
// dense in shared memory (no swizzle); rows past n_rows read as zeros.
int make_maps(slab::Maps* maps, const float* X, int64_t n_rows, int H, int64_t ldx) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(GCS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  for (int k = 0; k < 8; ++k) {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(n_rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldx) * sizeof(float)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(4 << (k & 3)), static_cast<cuuint32_t>(k < 4 ? slab::kBoxRows : slab::kBigBoxRows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&maps->m[k >> 2][k & 3], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(GCS_ERR_CUDA, "cuTensorMapEncodeTiled (slab) failed with CUresult %d", static_cast<int>(r));
  }
  return GCS_OK;
}

// The eight tensor maps of an operand only depend on (X, n_rows, H, ldx); a training loop presents the same handful of
// operands every step (the workspace is fixed), so they are encoded once per operand and host thread (~20 us saved per
// launch, which otherwise sits between two kernels whenever the stream has run dry).
struct MapKey {
  const float* X; int64_t n_rows, ldx; int H;
  bool operator==(const MapKey& o) const { return X == o.X && n_rows == o.n_rows && ldx == o.ldx && H == o.H; }
};
struct MapEntry { MapKey key; slab::Maps maps; };
const slab::Maps* cached_maps(const float* X, int64_t n_rows, int H, int64_t ldx, int* status) {
  thread_local MapEntry* cache = nullptr;             // 32 entries, round-robin replacement
  thread_local int filled = 0, next = 0;
  const MapKey key{X, n_rows, ldx, H};
  *status = GCS_OK;
  for (int i = 0; i < filled; ++i)
    if (cache[i].key == key) return &cache[i].maps;
  if (!cache) {
    void* mem = nullptr;
    if (posix_memalign(&mem, 64, 32 * sizeof(MapEntry)) != 0) { *status = fail(GCS_ERR_CUDA, "out of host memory"); return nullptr; }
    cache = static_cast<MapEntry*>(mem);
  }
  MapEntry& e = cache[next];
  *status = make_maps(&e.maps, X, n_rows, H, ldx);
  if (*status != GCS_OK) { e.key = MapKey{nullptr, 0, 0, 0}; return nullptr; }
  e.key = key;
  next = (next + 1) % 32;
  if (filled < 32) ++filled;
  return &e.maps;
}

template <int RB>
int launch_slab(int64_t n_rows, const int32_t* graph_ptr, int n_graphs, const int32_t* blk_ptr, const uint32_t* ent, const float* X,
                int64_t ldx, const float* scale, const float* shift, const float* alpha, const float* R, int64_t ldr,
                float* Y, int64_t ldy, int H, cudaStream_t st) {
  int map_status = GCS_OK;
  const slab::Maps* maps_p = cached_maps(X, n_rows, H, ldx, &map_status);
  if (!maps_p) return map_status;
  const slab::Maps& maps = *maps_p;
  const int sb = slab::stage_bytes();
  const int smem = slab::kHeaderBytes + 128 + slab::g_stages * sb;
  const int64_t items = static_cast<int64_t>(n_graphs) * ((H + slab::kCols - 1) / slab::kCols);
  int dev = 0;
  GCS_CUDA(cudaGetDevice(&dev));
  int grid = slab::g_grid > 0 ? slab::g_grid : sm_count();
  if (grid > items) grid = static_cast<int>(items);
  static unsigned int* tickets_dev[64] = {};           // per device
  if (!tickets_dev[dev & 63]) GCS_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&tickets_dev[dev & 63]), slab::g_tickets));
  static std::atomic<unsigned int> launch_seq{0};
  unsigned int* const ticket = tickets_dev[dev & 63] + 2 * (launch_seq.fetch_add(1u) % slab::kTicketSlots);
#define GCS_SLAB_LAUNCH(T)                                                                                             \
  do {                                                                                                                 \
    static int attr_smem[64] = {};   /* per instantiation and device; the attribute only ever grows */               \
    if (smem > attr_smem[dev & 63]) {                                                                                  \
      GCS_CUDA(cudaFuncSetAttribute(slab::spmm_slab_kernel<RB, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      attr_smem[dev & 63] = smem;                                                                                      \
    }                                                                                                                  \
    slab::spmm_slab_kernel<RB, T><<<grid, slab::kThreads, smem, st>>>(maps, graph_ptr, n_graphs, blk_ptr, ent, X, ldx, scale, \
                                                                      shift, alpha, R, ldr, Y, ldy, H, sb, slab::g_stages, \
                                                                      amax_sink().produce, ticket, slab::g_dbg);       \
  } while (0)
  if (scale) GCS_SLAB_LAUNCH(true); else GCS_SLAB_LAUNCH(false);
#undef GCS_SLAB_LAUNCH
  GCS_CHECK_LAUNCH("spmm_slab_kernel");
  return GCS_OK;
}

// Whether the slab kernel takes the batch: the longest graph must fit a stage at the narrowest pass (4 columns), and the
// typical graph should still get >= 64-byte rows (16 columns), otherwise the global-memory kernels are the better fit.
bool slab_fits(int64_t n_rows, int32_t n_graphs, int32_t max_graph_nodes) {
  if (n_graphs <= 0 || max_graph_nodes <= 0) return false;
  const int64_t sb = slab::stage_bytes();
  if ((static_cast<int64_t>(max_graph_nodes) + slab::kBoxRows) * 16 > sb) return false;
  return (n_rows / n_graphs + slab::kBoxRows) * 64 <= sb;
}

}  // namespace

// Profiling hook (not part of the drop-in surface): 8 uint64 device counters the slab kernel adds its per-role clock
// cycles to - {producer wait, work, prologue wait, work, gather wait, work, items, -}; NULL switches it off.
extern "C" void gcs_debug_slab_timing(unsigned long long* counters_dev) { slab::g_dbg = counters_dev; }

extern "C" int64_t gcs_spmm_slab_stage_bytes(void) { return slab::stage_bytes(); }

extern "C" int gcs_spmm_sum_graphs(const int32_t* graph_ptr, int32_t n_graphs, int32_t max_graph_nodes,
                                   const int32_t* rowptr, const int32_t* colidx, const int32_t* rb_blk_ptr,
                                   const uint32_t* rb_ent, int32_t rb_height, int64_t n_rows, const float* X, int64_t ldx,
                                   const float* scale, const float* shift, const float* alpha, const float* residual,
                                   int64_t ldr, float* Y, int64_t ldy, int32_t H, gcs_stream stream) {
  GCS_CHECK_ARG(n_rows >= 0 && H > 0, "gcs_spmm_sum_graphs: bad size (n_rows=%lld, H=%d)", (long long)n_rows, H);
  if (n_rows == 0) return GCS_OK;
  GCS_CHECK_ARG(rowptr && colidx && X && Y, "gcs_spmm_sum_graphs: null pointer");
  GCS_CHECK_ARG(ldx >= H && ldy >= H, "gcs_spmm_sum_graphs: leading dimension smaller than H");
  GCS_CHECK_ARG(!residual || ldr >= H, "gcs_spmm_sum_graphs: residual leading dimension smaller than H");
  GCS_CHECK_ARG((scale != nullptr) == (shift != nullptr) && (scale != nullptr) == (alpha != nullptr),
                "gcs_spmm_sum_graphs: scale/shift/alpha must be all NULL or all set");
  GCS_CHECK_ARG((rb_blk_ptr != nullptr) == (rb_ent != nullptr), "gcs_spmm_sum_graphs: rb_blk_ptr and rb_ent go together");
  GCS_CHECK_ARG(!rb_blk_ptr || rb_height == 2 || rb_height == 4, "gcs_spmm_sum_graphs: row-block height must be 2 or 4");
  GCS_CHECK_ARG(X != Y, "gcs_spmm_sum_graphs: in-place aggregation is not defined");
  GCS_CHECK_ARG(n_rows < INT32_MAX, "gcs_spmm_sum_graphs: n_rows exceeds int32 CSR range");
  const bool vec_ok = (H % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(X) && aligned16(Y) &&
                      (!scale || (aligned16(scale) && aligned16(shift) && aligned16(alpha))) &&
                      (!residual || ((ldr % 4 == 0) && aligned16(residual)));
  const int mode = spmm_mode();
  const bool slab = graph_ptr && rb_blk_ptr && aligned16(rb_ent) && aligned16(rb_blk_ptr) && vec_ok && mode != 1 && mode != 2 &&
                    slab_fits(n_rows, n_graphs, max_graph_nodes);
  if (!slab) {
    const bool rb4 = rb_blk_ptr && rb_height == 4;
    return gcs_spmm_aggregate(rowptr, colidx, nullptr, rb4 ? rb_blk_ptr : nullptr, rb4 ? rb_ent : nullptr, n_rows, X, ldx,
                              scale, shift, alpha, residual, ldr, Y, ldy, H, 0, stream);
  }
  cudaStream_t st = as_stream(stream);
  if (rb_height == 2)
    return launch_slab<2>(n_rows, graph_ptr, n_graphs, rb_blk_ptr, rb_ent, X, ldx, scale, shift, alpha, residual, ldr, Y, ldy, H, st);
  return launch_slab<4>(n_rows, graph_ptr, n_graphs, rb_blk_ptr, rb_ent, X, ldx, scale, shift, alpha, residual, ldr, Y, ldy, H, st);
}
