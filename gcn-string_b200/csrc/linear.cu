// K1 / K9 — dense transforms on the CUDA cores (exact fp32 FFMA; the correctness anchor and
// the path for widths below 128; the tcgen05 path in linear_tc.cu takes over above that).
//
// Replaces Keras Dense / K.dot + K.bias_add inside MLP and GeneralConv.call (SURVEY.md §8
// a3/a4; reached from src/scripts/gcn.py:320,334) and their reverse-mode gradients
// (gcn.py:337):   fwd  C = A.W + b        dW = A^T.dH, db = colsum(dH)       dA (+)= dH.W^T
//
// One tiled kernel computes  C[i,j] = sum_r P(i,r) Q(r,j)  for the three layouts:
//   fwd        P = A  [I=M, R=K]  r-contiguous     Q = W  [R=K, J=N]  j-contiguous
//   bwd_input  P = dH [I=M, R=N]  r-contiguous     Q = W  [J=K, R=N]  r-contiguous
//   bwd_weight P = A  [R=M, I=K]  i-contiguous     Q = dH [R=M, J=N]  j-contiguous (split over R)
// 128x128x16 tiles, 256 threads, 8x8 accumulators per thread (2x2 blocks of 4x4 so that
// every shared-memory read is a conflict-free 128-bit load), register-prefetched double
// buffering: one __syncthreads per 16-deep slice.
#include "common.cuh"
#include "prep.cuh"

namespace gcs {

constexpr int BT = 128;      // tile edge in i and j
constexpr int BR = 16;       // reduction slice
constexpr int PAD = 4;       // smem row padding (keeps 16 B alignment, halves store conflicts)
constexpr int kGemmThreads = 256;

// Operand tile loader.  RC = true: operand stored [X, R] row-major (reduction contiguous);
// RC = false: stored [R, X] row-major (output dimension contiguous).  Each thread moves two
// 4-element pieces per tile.
template <bool RC, bool VEC>
struct TileLoader {
  float4 v[2];
  __device__ __forceinline__ void load(const float* __restrict__ p, int64_t ld, int64_t x0, int64_t X,
                                       int64_t r0, int64_t r_end) {
    const int t = threadIdx.x;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int unit = t + u * kGemmThreads;
      int64_t x, r;
      if (RC) { x = x0 + unit / 4; r = r0 + (unit % 4) * 4; }
      else    { r = r0 + unit / 32; x = x0 + (unit % 32) * 4; }
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (VEC) {
        if (x < X && r < r_end)
          val = __ldg(reinterpret_cast<const float4*>(RC ? p + x * ld + r : p + r * ld + x));
      } else {
        float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int64_t xx = RC ? x : x + k, rr = RC ? r + k : r;
          if (xx < X && rr < r_end) e[k] = __ldg(RC ? p + xx * ld + rr : p + rr * ld + xx);
        }
        val = make_float4(e[0], e[1], e[2], e[3]);
      }
      v[u] = val;
    }
  }
  // smem tile layout: s[r][x], row stride BT + PAD
  __device__ __forceinline__ void store(float* __restrict__ s) const {
    const int t = threadIdx.x;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int unit = t + u * kGemmThreads;
      if (RC) {
        const int x = unit / 4, r = (unit % 4) * 4;
        s[(r + 0) * (BT + PAD) + x] = v[u].x;
        s[(r + 1) * (BT + PAD) + x] = v[u].y;
        s[(r + 2) * (BT + PAD) + x] = v[u].z;
        s[(r + 3) * (BT + PAD) + x] = v[u].w;
      } else {
        const int r = unit / 32, x = (unit % 32) * 4;
        *reinterpret_cast<float4*>(s + r * (BT + PAD) + x) = v[u];
      }
    }
  }
};

// C (+)= P.Q (+ bias).  grid: (i tiles, j tiles, R splits); split z handles reduction range
// [z*r_per_split, min(R, (z+1)*r_per_split)) and writes to C + z*c_split_stride.
template <bool P_RC, bool Q_JC, bool VEC>
__global__ void __launch_bounds__(kGemmThreads, 2) sgemm_kernel(
    const float* __restrict__ P, int64_t ldp, const float* __restrict__ Q, int64_t ldq,
    float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int64_t I, int J, int64_t R,
    int64_t r_per_split, int accumulate, int64_t c_split_stride) {
  __shared__ __align__(16) float Ps[2][BR][BT + PAD];
  __shared__ __align__(16) float Qs[2][BR][BT + PAD];
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * BT;
  const int j0 = blockIdx.y * BT;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.z) * r_per_split;
  int64_t r_end = r_begin + r_per_split;
  if (r_end > R) r_end = R;
  C += static_cast<int64_t>(blockIdx.z) * c_split_stride;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;

  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

  TileLoader<P_RC, VEC> lp;
  TileLoader<!Q_JC, VEC> lq;
  const int64_t n_slices = r_end > r_begin ? (r_end - r_begin + BR - 1) / BR : 0;
  if (n_slices > 0) {
    lp.load(P, ldp, i0, I, r_begin, r_end);
    lq.load(Q, ldq, j0, J, r_begin, r_end);
    lp.store(&Ps[0][0][0]);
    lq.store(&Qs[0][0][0]);
  }
  __syncthreads();
  for (int64_t s = 0; s < n_slices; ++s) {
    const int cur = static_cast<int>(s & 1);
    const bool more = s + 1 < n_slices;
    if (more) {
      lp.load(P, ldp, i0, I, r_begin + (s + 1) * BR, r_end);
      lq.load(Q, ldq, j0, J, r_begin + (s + 1) * BR, r_end);
    }
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      const float4 a0 = *reinterpret_cast<const float4*>(&Ps[cur][r][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&Ps[cur][r][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Qs[cur][r][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Qs[cur][r][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int x = 0; x < 8; ++x)
#pragma unroll
        for (int y = 0; y < 8; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    if (more) {
      lp.store(&Ps[cur ^ 1][0][0]);
      lq.store(&Qs[cur ^ 1][0][0]);
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int x = 0; x < 8; ++x) {
    const int64_t i = i0 + (x < 4 ? ty * 4 + x : 64 + ty * 4 + (x - 4));
    if (i >= I) continue;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int j = j0 + half * 64 + tx * 4;
      float* cp = C + i * ldc + j;
      if (VEC) {
        if (j < J) {
          float4 o = make_float4(acc[x][half * 4], acc[x][half * 4 + 1], acc[x][half * 4 + 2], acc[x][half * 4 + 3]);
          if (bias) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + j));
            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
          }
          if (accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(cp);
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *reinterpret_cast<float4*>(cp) = o;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (j + k < J) {
            float o = acc[x][half * 4 + k];
            if (bias) o += __ldg(bias + j + k);
            if (accumulate) o += cp[k];
            cp[k] = o;
          }
        }
      }
    }
  }
}

// dW = sum over splits (fixed order), optional column sums for db handled separately.
__global__ void __launch_bounds__(256) split_reduce_kernel(const float* __restrict__ ws, int splits,
                                                           int64_t n, float* __restrict__ out) {
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float s = ws[k];
    for (int z = 1; z < splits; ++z) s += ws[static_cast<int64_t>(z) * n + k];
    out[k] = s;
  }
}

// db[j] = sum_m dH[m, j]: per-split column partials in fp64 (lane = 4 columns when aligned, 8 row
// warps), combined in split order.
template <int VEC>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ dH, int64_t ldh,
                                                             int64_t M, int N, int64_t rows_per_split,
                                                             double* __restrict__ ws) {
  __shared__ double sm[8][32][VEC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + lane) * VEC;
  const int64_t rb = blockIdx.y * rows_per_split;
  int64_t re = rb + rows_per_split;
  if (re > M) re = M;
  double acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.0;
  if (c < N) {
    for (int64_t r = rb + warp; r < re; r += 8) {
      if (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(dH + r * ldh + c));
        acc[0] += static_cast<double>(t.x); acc[1] += static_cast<double>(t.y);
        acc[2] += static_cast<double>(t.z); acc[3] += static_cast<double>(t.w);
      } else {
        acc[0] += static_cast<double>(__ldg(dH + r * ldh + c));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) sm[warp][lane][k] = acc[k];
  __syncthreads();
  if (warp == 0 && c < N) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      double t = sm[0][lane][k];
      for (int w = 1; w < 8; ++w) t += sm[w][lane][k];
      ws[static_cast<int64_t>(blockIdx.y) * N + c + k] = t;
    }
  }
}

__global__ void colsum_final_kernel(const double* __restrict__ ws, int splits, int N, float* __restrict__ db) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  double s = 0.0;
  for (int z = 0; z < splits; ++z) s += ws[static_cast<int64_t>(z) * N + c];
  db[c] = static_cast<float>(s);
}

// Weight gradient of a THIN layer (K <= 32 input features: the first Dense of the pre-processing MLP, K = F = 16 / 32):
// dW[K, N] = A[M, K]^T . dH[M, N].  The tensor-core kernels need K % 128 == 0 and the tiled FFMA kernel leaves 3/4 of
// its 128-wide tile empty (676 us at cfg2).  Here a CTA of K/4 warps owns 128 columns x all K rows of dW for one slab
// of the M rows: warp = 4 consecutive k, lane = 4 consecutive columns (16 accumulators); per row one 128-bit load of
// dH (coalesced, 4 rows in flight), one broadcast LDS.128 of the staged A row and 16 FFMA.  Partials per row slab,
// summed in slab order by split_reduce_kernel: deterministic.  8.4 GFLOP at cfg2 -> FFMA-bound at ~115 us.
template <int K4>
__global__ void __launch_bounds__(32 * K4) wgrad_thin_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ dH,
                                                             int64_t ldh, float* __restrict__ part, int64_t M, int N,
                                                             int64_t rows_per_split) {
  constexpr int K = 4 * K4, CH = 64, NT = 32 * K4;
  __shared__ __align__(16) float sA[CH][K];
  const int t = threadIdx.x;
  const int lane = t & 31, kg = t >> 5;               // columns 4*lane .. +3 of the CTA's 128, rows 4*kg .. +3 of dW
  const int c = blockIdx.x * 128 + 4 * lane;
  const int64_t r0 = blockIdx.y * rows_per_split;
  const int64_t r1 = r0 + rows_per_split < M ? r0 + rows_per_split : M;
  const bool live = c < N;                             // N % 4 == 0: a quad is inside or outside
  float4 acc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto fma4 = [&](const float4& a, const float4& d) {
    acc[0].x = fmaf(a.x, d.x, acc[0].x); acc[0].y = fmaf(a.x, d.y, acc[0].y); acc[0].z = fmaf(a.x, d.z, acc[0].z); acc[0].w = fmaf(a.x, d.w, acc[0].w);
    acc[1].x = fmaf(a.y, d.x, acc[1].x); acc[1].y = fmaf(a.y, d.y, acc[1].y); acc[1].z = fmaf(a.y, d.z, acc[1].z); acc[1].w = fmaf(a.y, d.w, acc[1].w);
    acc[2].x = fmaf(a.z, d.x, acc[2].x); acc[2].y = fmaf(a.z, d.y, acc[2].y); acc[2].z = fmaf(a.z, d.z, acc[2].z); acc[2].w = fmaf(a.z, d.w, acc[2].w);
    acc[3].x = fmaf(a.w, d.x, acc[3].x); acc[3].y = fmaf(a.w, d.y, acc[3].y); acc[3].z = fmaf(a.w, d.z, acc[3].z); acc[3].w = fmaf(a.w, d.w, acc[3].w);
  };
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t rc = r0; rc < r1; rc += CH) {
    const int rows = static_cast<int>(r1 - rc < CH ? r1 - rc : CH);
    __syncthreads();                                   // the previous chunk has been consumed
    for (int i = t; i < CH * K4; i += NT) {
      const int rr = i / K4, q = i - rr * K4;
      *reinterpret_cast<float4*>(&sA[rr][4 * q]) = rr < rows ? __ldg(reinterpret_cast<const float4*>(A + (rc + rr) * lda + 4 * q)) : zero;
    }
    __syncthreads();
    const float* dcol = dH + rc * ldh + c;
    if (!live) continue;                               // (uniform per warp: a warp's 32 quads are 128 consecutive columns)
    if (rows == CH) {
      // full chunk: groups of 4 rows, two register sets in ping-pong so that 8 dH loads are in flight behind the FMAs
      float4 d0[4], d1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) d0[u] = __ldg(reinterpret_cast<const float4*>(dcol + u * ldh));
#pragma unroll
      for (int rr = 0; rr < CH; rr += 8) {
#pragma unroll
        for (int u = 0; u < 4; ++u) d1[u] = __ldg(reinterpret_cast<const float4*>(dcol + (rr + 4 + u) * ldh));
#pragma unroll
        for (int u = 0; u < 4; ++u) fma4(*reinterpret_cast<const float4*>(&sA[rr + u][4 * kg]), d0[u]);
        if (rr + 8 < CH) {
#pragma unroll
          for (int u = 0; u < 4; ++u) d0[u] = __ldg(reinterpret_cast<const float4*>(dcol + (rr + 8 + u) * ldh));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) fma4(*reinterpret_cast<const float4*>(&sA[rr + 4 + u][4 * kg]), d1[u]);
      }
    } else {
      for (int rr = 0; rr < rows; ++rr)
        fma4(*reinterpret_cast<const float4*>(&sA[rr][4 * kg]), __ldg(reinterpret_cast<const float4*>(dcol + rr * ldh)));
    }
  }
  if (live) {
    float* out = part + static_cast<int64_t>(blockIdx.y) * K * N + static_cast<int64_t>(4 * kg) * N + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(out + static_cast<int64_t>(k) * N) = acc[k];
  }
}

static int g_thin = 1;     // gcs_debug_set_param 13: 0 sends thin layers to the tiled FFMA kernel again
void set_thin_wgrad(int v) { g_thin = v; }
static int thin_splits(int64_t M, int N) {
  const int64_t colblocks = ceil_div(N, 128);
  int64_t s = ceil_div(4LL * sm_count(), colblocks);
  const int64_t max_s = ceil_div(M > 0 ? M : 1, 256);
  if (s > max_s) s = max_s;
  return static_cast<int>(s < 1 ? 1 : s);
}
static bool thin_ok(int64_t M, int K, int N, const float* A, int64_t lda, const float* dH) {
  return g_thin && K <= 32 && K % 4 == 0 && N % 4 == 0 && lda % 4 == 0 && aligned16(A) && dH && M >= 4096 && N >= 64;
}

struct WeightSplit {
  int splits;
  int64_t r_per_split;
  int col_splits;
  int64_t col_rows_per_split;
};

static WeightSplit weight_split(int64_t M, int K, int N) {
  WeightSplit w;
  const int64_t tiles = ceil_div(K, BT) * ceil_div(N, BT);
  int64_t s = ceil_div(2LL * sm_count(), tiles);
  const int64_t max_s = ceil_div(M > 0 ? M : 1, 4 * BR);
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  w.r_per_split = round_up(ceil_div(M > 0 ? M : 1, s), BR);
  w.splits = static_cast<int>(ceil_div(M > 0 ? M : 1, w.r_per_split));
  int64_t cs = ceil_div(4LL * sm_count(), ceil_div(N, 128));
  const int64_t max_cs = ceil_div(M > 0 ? M : 1, 64);
  if (cs > max_cs) cs = max_cs;
  if (cs < 1) cs = 1;
  w.col_rows_per_split = ceil_div(M > 0 ? M : 1, cs);
  w.col_splits = static_cast<int>(ceil_div(M > 0 ? M : 1, w.col_rows_per_split));
  return w;
}

}  // namespace gcs

namespace gcs {
namespace tc {   // linear_tc.cu
bool shape_ok(int64_t M, int K, int N, const float* A, int64_t lda, const float* C, int64_t ldc, const float* bias);
int64_t split_workspace_bytes(int K, int N);
int launch(const float* A, int64_t lda, const float* Bt_hi, const float* Bt_lo, const float* bias, float* C, int64_t ldc,
           int64_t M, int K, int N, int accumulate, cudaStream_t st, const float* rowbias = nullptr, int64_t ld_rowbias = 0,
           const int64_t* seg = nullptr);
int split(const float* W, int rows, int cols, bool transpose, float* hi, float* lo, cudaStream_t st);
int split_strided(const float* W, int rows, int cols, int64_t ldw, bool transpose, int64_t ldo, float* hi, float* lo,
                  cudaStream_t st);
int f16_mode();
bool f16_shape_ok(int K);
float* f16_cells(void* workspace, int K, int N);
int f16_begin(float* cells, cudaStream_t st);
int f16_max_chain();
int absmax(const float* A, int64_t lda, int64_t M, int K, float* cell, cudaStream_t st, const float* col_scale = nullptr);
int fold_bias(const float* bias, const float* scale, const float* shift, int N, float* out, cudaStream_t st);
int amax_merge(float* cell, const float* other, cudaStream_t st);
int split_f16_strided(const float* W, int rows, int cols, int64_t ldw, bool transpose, int64_t ldo, void* hi, void* lo,
                      float* cells, cudaStream_t st, const float* col_scale = nullptr);
int launch_f16(const float* A, int64_t lda, const void* Bt_hi, const void* Bt_lo, const float* cells, const float* a_amax,
               const float* bias, float* C, int64_t ldc, int64_t M, int K, int N, int accumulate, cudaStream_t st,
               const float* rowbias = nullptr, int64_t ld_rowbias = 0, const int64_t* seg = nullptr,
               const float* alpha = nullptr, float* amax_out = nullptr, float* stats_part = nullptr);
bool wgrad_shape_ok(int64_t M, int K, int N, const float* A, int64_t lda, const float* dH, int64_t ldh);
void wgrad_split(int64_t M, int K, int N, int* splits, int* kb_per_split);
int64_t wgrad_workspace_bytes(int64_t M, int K, int N);
int wgrad_launch(const float* A, int64_t lda, const float* dH, int64_t ldh, float* out, int64_t M, int K, int N,
                 int splits, int kb_per_split, cudaStream_t st, const float* a_amax = nullptr, const float* b_amax = nullptr);
bool wgrad_f16_ok(int K);
}  // namespace tc
}  // namespace gcs

namespace gcs {
PreparedTable*& prepared_table() {
  thread_local PreparedTable* t = nullptr;
  return t;
}
void* find_prepared(int kind, const float* w, int kred, int nout) {
  const PreparedTable* t = prepared_table();
  if (!t) return nullptr;
  for (int i = 0; i < t->n; ++i)
    if (t->e[i].kind == kind && t->e[i].w == w && t->e[i].kred == kred && t->e[i].nout == nout) return t->e[i].ws;
  return nullptr;
}
}  // namespace gcs

using namespace gcs;

static int g_gemm_mode = 0;   // 0 = auto (tensor cores when the shape allows), 1 = CUDA cores only, 2 = tensor cores or error
extern "C" void gcs_debug_set_gemm_mode(int mode) { g_gemm_mode = mode; }

// Tensor-core path if the shape, alignment and workspace allow it.  *used = 1 when it ran.
static int try_tensor_cores(const float* A, int64_t lda, const float* W, int w_rows, int w_cols, bool transpose_w,
                            const float* bias, float* C, int64_t ldc, int64_t M, int Kred, int Nout, int accumulate,
                            void* workspace, int64_t workspace_bytes, cudaStream_t st, int* used) {
  *used = 0;
  if (g_gemm_mode == 1) return GCS_OK;
  const bool ok = tc::shape_ok(M, Kred, Nout, A, lda, C, ldc, bias) && workspace && aligned16(workspace) &&
                  workspace_bytes >= tc::split_workspace_bytes(Kred, Nout);
  if (!ok) {
    if (g_gemm_mode == 2) return fail(GCS_ERR_UNSUPPORTED, "tensor-core GEMM forced but shape/workspace do not allow it");
    return GCS_OK;
  }
  // fp16 split (half the tensor time) when the |max| of A is known: from the producer kernels of the fused model
  // (amax_sink().consume), or - debug mode 2 - from an extra pass over A.
  const float* a_amax = amax_sink().consume;
  if (tc::f16_mode() && tc::f16_shape_ok(Kred) && (a_amax || tc::f16_mode() == 2)) {
    void* const ready = (a_amax && !transpose_w) ? find_prepared(2, W, Kred, Nout) : nullptr;   // split at the start of the step
    void* const wsp = ready ? ready : workspace;
    char* hi = static_cast<char*>(wsp);
    char* lo = hi + 2LL * Kred * Nout;
    float* cells = tc::f16_cells(wsp, Kred, Nout);
    if (!ready) {
      GCS_TRY(tc::f16_begin(cells, st));
      GCS_TRY(tc::absmax(W, w_cols, w_rows, w_cols, cells, st));
      if (!a_amax) {
        GCS_TRY(tc::absmax(A, lda, M, Kred, cells + 2, st));
        a_amax = cells + 2;
      }
      GCS_TRY(tc::split_f16_strided(W, w_rows, w_cols, w_cols, transpose_w, transpose_w ? w_rows : w_cols, hi, lo, cells, st));
    }
    GCS_TRY(tc::launch_f16(A, lda, hi, lo, cells, a_amax, bias, C, ldc, M, Kred, Nout, accumulate, st));
    *used = 1;
    return GCS_OK;
  }
  float* hi = static_cast<float*>(workspace);
  float* lo = hi + static_cast<int64_t>(Kred) * Nout;
  GCS_TRY(tc::split(W, w_rows, w_cols, transpose_w, hi, lo, st));
  GCS_TRY(tc::launch(A, lda, hi, lo, bias, C, ldc, M, Kred, Nout, accumulate, st));
  *used = 1;
  return GCS_OK;
}

template <bool P_RC, bool Q_JC>
static int launch_sgemm(const float* P, int64_t ldp, const float* Q, int64_t ldq, float* C, int64_t ldc,
                        const float* bias, int64_t I, int J, int64_t R, int splits, int64_t r_per_split,
                        int accumulate, int64_t c_split_stride, bool vec, cudaStream_t st, const char* who) {
  const int64_t ti = ceil_div(I, BT);
  if (ti > 0x7fffffffLL) return fail(GCS_ERR_INVALID_ARGUMENT, "%s: too many row tiles", who);
  dim3 grid(static_cast<unsigned>(ti), static_cast<unsigned>(ceil_div(J, BT)), splits);
  if (vec)
    sgemm_kernel<P_RC, Q_JC, true><<<grid, kGemmThreads, 0, st>>>(P, ldp, Q, ldq, C, ldc, bias, I, J, R, r_per_split, accumulate, c_split_stride);
  else
    sgemm_kernel<P_RC, Q_JC, false><<<grid, kGemmThreads, 0, st>>>(P, ldp, Q, ldq, C, ldc, bias, I, J, R, r_per_split, accumulate, c_split_stride);
  GCS_CHECK_LAUNCH(who);
  return GCS_OK;
}

namespace gcs {

// One dense block as ONE fp16 tensor-core GEMM with a fused epilogue.
//  * inference (scale != NULL): C = prelu((A . W + b) * scale + shift, alpha) - the moving-statistics
//    BatchNorm is folded into the split weights (columns times scale) and the bias (b * scale + shift), PReLU runs in the
//    epilogue (alpha == NULL: none); amax_out (optional) receives the |max| of C (zeroed here).
//  * training (scale == NULL, stats_part != NULL): C = A . W + b, and the epilogue leaves {sum, sum of squares} per
//    32-row group and column in stats_part [ceil(M/32)][N][2] for the BatchNorm that follows (bn_stats_from_partials).
// Needs the fp16 path (|max| of A known through amax_sink().consume, reduction in one chain); *used = 0 otherwise and
// nothing was launched.
int linear_fwd_fused(const float* A, int64_t lda, const float* W, const float* bias, const float* scale,
                     const float* shift, const float* alpha, float* C, int64_t ldc, int64_t M, int K, int N,
                     void* workspace, int64_t workspace_bytes, float* amax_out, float* stats_part, cudaStream_t st,
                     int* used) {
  *used = 0;
  const float* a_amax = amax_sink().consume;
  const int64_t need = round_up(4LL * K * N, 256) + 256 + round_up(4LL * N, 256);
  if (g_gemm_mode == 1 || !tc::f16_mode() || !a_amax || !tc::f16_shape_ok(K) || K > tc::f16_max_chain() ||
      !tc::shape_ok(M, K, N, A, lda, C, ldc, bias) || !workspace || !aligned16(workspace) || workspace_bytes < need ||
      (alpha && !aligned16(alpha)))
    return GCS_OK;
  void* const ready = scale ? nullptr : find_prepared(0, W, K, N);      // split at the start of the step (training)
  void* const wsp = ready ? ready : workspace;
  char* hi = static_cast<char*>(wsp);
  char* lo = hi + 2LL * K * N;
  float* cells = tc::f16_cells(wsp, K, N);
  const float* b = bias;
  if (!ready) {
    GCS_TRY(tc::f16_begin(cells, st));
    GCS_TRY(tc::absmax(W, N, K, N, cells, st, scale));
    if (scale) {
      float* bias2 = cells + 64;
      GCS_TRY(tc::fold_bias(bias, scale, shift, N, bias2, st));
      b = bias2;
    }
    GCS_TRY(tc::split_f16_strided(W, K, N, N, true, K, hi, lo, cells, st, scale));
  }
  if (amax_out) GCS_CUDA(cudaMemsetAsync(amax_out, 0, sizeof(float), st));
  GCS_TRY(tc::launch_f16(A, lda, hi, lo, cells, a_amax, b, C, ldc, M, K, N, 0, st, nullptr, 0, nullptr, alpha, amax_out, stats_part));
  *used = 1;
  return GCS_OK;
}

// Gradient w.r.t. one H-wide block of the concatenated node embedding, all consumers at once:
//   C[M, Nout] = rowbias[seg[m]] + sum_b  DH_b[M, Hred] . W_b[row_off_b : row_off_b + Nout, 0 : Hred]^T
// where the DH_b are adjacent column blocks of one buffer (dh, leading dimension ld).  One long-K
// tensor-core GEMM (no read-modify-write of C) when the shape allows; otherwise the segment
// broadcast is materialised and the blocks are accumulated one by one on the CUDA cores.
int dense_dx_concat(const float* dh, int64_t ld, const float* const* W, const int* row_off, int n_blocks, int Hred,
                    int Nout, const float* rowbias, int64_t ld_rowbias, const int64_t* seg, const int32_t* graph_ptr,
                    int n_graphs, float* C, int64_t ldc, int64_t M, int accumulate, void* workspace,
                    int64_t workspace_bytes, cudaStream_t st) {
  const int Kred = n_blocks * Hred;
  const bool tc_ok = g_gemm_mode != 1 && n_blocks > 0 && tc::shape_ok(M, Kred, Nout, dh, ld, C, ldc, nullptr) && workspace &&
                     aligned16(workspace) && workspace_bytes >= tc::split_workspace_bytes(Kred, Nout) &&
                     (!rowbias || (ld_rowbias % 4 == 0 && aligned16(rowbias)));
  const float* a_amax = amax_sink().consume;
  if (tc_ok && tc::f16_mode() && tc::f16_shape_ok(Kred) && (a_amax || tc::f16_mode() == 2)) {
    void* const ready = a_amax ? find_prepared(1, W[0] + static_cast<int64_t>(row_off[0]) * Hred, Kred, Nout) : nullptr;
    void* const wsp = ready ? ready : workspace;
    char* hi = static_cast<char*>(wsp);
    char* lo = hi + 2LL * Kred * Nout;
    float* cells = tc::f16_cells(wsp, Kred, Nout);
    if (!ready) {
      GCS_TRY(tc::f16_begin(cells, st));
      for (int b = 0; b < n_blocks; ++b)
        GCS_TRY(tc::absmax(W[b] + static_cast<int64_t>(row_off[b]) * Hred, Hred, Nout, Hred, cells, st));
      if (!a_amax) {
        GCS_TRY(tc::absmax(dh, ld, M, Kred, cells + 2, st));
        a_amax = cells + 2;
      }
      for (int b = 0; b < n_blocks; ++b)   // Bt[c][b*Hred + n] = W_b[row_off_b + c][n]
        GCS_TRY(tc::split_f16_strided(W[b] + static_cast<int64_t>(row_off[b]) * Hred, Nout, Hred, Hred, false, Kred,
                                      hi + 2LL * b * Hred, lo + 2LL * b * Hred, cells, st));
    }
    return tc::launch_f16(dh, ld, hi, lo, cells, a_amax, nullptr, C, ldc, M, Kred, Nout, accumulate, st, rowbias, ld_rowbias, seg);
  }
  if (tc_ok) {
    float* hi = static_cast<float*>(workspace);
    float* lo = hi + static_cast<int64_t>(Kred) * Nout;
    for (int b = 0; b < n_blocks; ++b)   // Bt[c][b*Hred + n] = W_b[row_off_b + c][n]
      GCS_TRY(tc::split_strided(W[b] + static_cast<int64_t>(row_off[b]) * Hred, Nout, Hred, Hred, false, Kred,
                                hi + static_cast<int64_t>(b) * Hred, lo + static_cast<int64_t>(b) * Hred, st));
    return tc::launch(dh, ld, hi, lo, nullptr, C, ldc, M, Kred, Nout, accumulate, st, rowbias, ld_rowbias, seg);
  }
  // CUDA-core path
  if (rowbias) {
    if (accumulate) return fail(GCS_ERR_INVALID_ARGUMENT, "dense_dx_concat: rowbias and accumulate are exclusive");
    GCS_TRY(gcs_segment_sum_bwd(rowbias, ld_rowbias, graph_ptr, n_graphs, Nout, C, ldc, st));
  } else if (n_blocks == 0) {
    return fail(GCS_ERR_INVALID_ARGUMENT, "dense_dx_concat: nothing to compute");
  }
  for (int b = 0; b < n_blocks; ++b) {
    const float* P = dh + static_cast<int64_t>(b) * Hred;
    const float* Q = W[b] + static_cast<int64_t>(row_off[b]) * Hred;       // [Nout rows][Hred], reduction contiguous
    const bool vec = Hred % 4 == 0 && Nout % 4 == 0 && ld % 4 == 0 && ldc % 4 == 0 && aligned16(P) && aligned16(Q) && aligned16(C);
    GCS_TRY((launch_sgemm<true, false>(P, ld, Q, Hred, C, ldc, nullptr, M, Nout, Hred, 1, Hred, (rowbias || accumulate || b > 0) ? 1 : 0, 0,
                                       vec, st, "dense_dx_concat")));
  }
  return GCS_OK;
}

}  // namespace gcs

extern "C" int64_t gcs_linear_workspace_bytes(int64_t M, int32_t K, int32_t N) {
  (void)M;
  if (K <= 0 || N <= 0) return 0;
  return tc::split_workspace_bytes(K, N);
}

extern "C" int gcs_linear_fwd(const float* A, int64_t lda, const float* W, const float* bias, float* C,
                              int64_t ldc, int64_t M, int32_t K, int32_t N, void* workspace,
                              int64_t workspace_bytes, gcs_stream stream) {
  GCS_CHECK_ARG(M >= 0 && K > 0 && N > 0, "gcs_linear_fwd: bad size M=%lld K=%d N=%d", (long long)M, K, N);
  if (M == 0) return GCS_OK;
  GCS_CHECK_ARG(A && W && C && lda >= K && ldc >= N, "gcs_linear_fwd: bad pointer / leading dimension");
  {
    // W [K, N] -> split + transposed to [N][K] (reduction contiguous) for the B operand
    int used = 0;
    GCS_TRY(try_tensor_cores(A, lda, W, K, N, true, bias, C, ldc, M, K, N, 0, workspace, workspace_bytes, as_stream(stream), &used));
    if (used) return GCS_OK;
  }
  const bool vec = K % 4 == 0 && N % 4 == 0 && lda % 4 == 0 && ldc % 4 == 0 && aligned16(A) && aligned16(W) &&
                   aligned16(C) && (!bias || aligned16(bias));
  return launch_sgemm<true, true>(A, lda, W, N, C, ldc, bias, M, N, K, 1, K, 0, 0, vec, as_stream(stream), "gcs_linear_fwd");
}

extern "C" int gcs_linear_bwd_input(const float* dH, int64_t ldh, const float* W, float* dA, int64_t lda,
                                    int64_t M, int32_t K, int32_t N, int32_t accumulate, void* workspace,
                                    int64_t workspace_bytes, gcs_stream stream) {
  GCS_CHECK_ARG(M >= 0 && K > 0 && N > 0, "gcs_linear_bwd_input: bad size");
  if (M == 0) return GCS_OK;
  GCS_CHECK_ARG(dH && W && dA && ldh >= N && lda >= K, "gcs_linear_bwd_input: bad pointer / leading dimension");
  {
    // dA[m, k] = sum_n dH[m, n] W[k, n]: W as stored IS the [out][reduction] layout, no transpose
    int used = 0;
    GCS_TRY(try_tensor_cores(dH, ldh, W, K, N, false, nullptr, dA, lda, M, N, K, accumulate ? 1 : 0, workspace,
                             workspace_bytes, as_stream(stream), &used));
    if (used) return GCS_OK;
  }
  const bool vec = K % 4 == 0 && N % 4 == 0 && lda % 4 == 0 && ldh % 4 == 0 && aligned16(dH) && aligned16(W) && aligned16(dA);
  // C[i=m, j=k] = sum_{r=n} dH[m, n] * W[k, n]
  return launch_sgemm<true, false>(dH, ldh, W, N, dA, lda, nullptr, M, K, N, 1, N, accumulate, 0, vec, as_stream(stream), "gcs_linear_bwd_input");
}

extern "C" int64_t gcs_linear_bwd_weight_workspace_bytes(int64_t M, int32_t K, int32_t N) {
  if (M < 0 || K <= 0 || N <= 0) return 0;
  const WeightSplit w = weight_split(M, K, N);
  int64_t part = round_up(static_cast<int64_t>(w.splits) * K * N * sizeof(float), 256);
  if (K <= 32 && K % 4 == 0) {
    const int64_t thin = round_up(static_cast<int64_t>(thin_splits(M, N)) * K * N * sizeof(float), 256);
    if (thin > part) part = thin;
  }
  if (K % 128 == 0 && N % 128 == 0) {
    const int64_t tcb = tc::wgrad_workspace_bytes(M, K, N);
    if (tcb > part) part = tcb;
  }
  return part + round_up(static_cast<int64_t>(w.col_splits) * N * sizeof(double), 256);
}

extern "C" int gcs_linear_bwd_weight(const float* A, int64_t lda, const float* dH, int64_t ldh, float* dW,
                                     float* db, int64_t M, int32_t K, int32_t N, void* workspace,
                                     int64_t workspace_bytes, gcs_stream stream) {
  GCS_CHECK_ARG(M > 0 && K > 0 && N > 0, "gcs_linear_bwd_weight: bad size");
  GCS_CHECK_ARG(A && dH && dW && workspace && lda >= K && ldh >= N, "gcs_linear_bwd_weight: bad pointer / leading dimension");
  if (workspace_bytes < gcs_linear_bwd_weight_workspace_bytes(M, K, N))
    return fail(GCS_ERR_WORKSPACE, "gcs_linear_bwd_weight: workspace %lld < %lld bytes", (long long)workspace_bytes,
                (long long)gcs_linear_bwd_weight_workspace_bytes(M, K, N));
  GCS_CHECK_ARG(aligned16(workspace), "gcs_linear_bwd_weight: workspace must be 16-byte aligned");
  const WeightSplit w = weight_split(M, K, N);
  cudaStream_t st = as_stream(stream);
  float* part = static_cast<float*>(workspace);
  const bool vec = K % 4 == 0 && N % 4 == 0 && lda % 4 == 0 && ldh % 4 == 0 && aligned16(A) && aligned16(dH) && aligned16(dW);
  const int64_t kn = static_cast<int64_t>(K) * N;
  const int64_t part_bytes = gcs_linear_bwd_weight_workspace_bytes(M, K, N) -
                             round_up(static_cast<int64_t>(w.col_splits) * N * sizeof(double), 256);
  const bool use_tc = g_gemm_mode != 1 && tc::wgrad_shape_ok(M, K, N, A, lda, dH, ldh) && aligned16(dW);
  if (g_gemm_mode == 2 && !use_tc)
    return fail(GCS_ERR_UNSUPPORTED, "tensor-core weight gradient forced but the shape does not allow it");
  if (use_tc) {
    int splits, per;
    tc::wgrad_split(M, K, N, &splits, &per);
    // fp16 split when the |max| of both operands is known (fused model) or, in debug mode 2, measured by an extra pass
    const float* a_amax = amax_sink().consume_act;
    const float* b_amax = amax_sink().consume;
    if (!(a_amax && b_amax)) a_amax = b_amax = nullptr;
    if (!a_amax && tc::f16_mode() == 2 && tc::wgrad_f16_ok(K)) {
      static float* scratch = nullptr;                       // debug mode only: two cells, allocated once
      if (!scratch) GCS_CUDA(cudaMalloc(&scratch, 256));
      GCS_CUDA(cudaMemsetAsync(scratch, 0, 2 * sizeof(float), st));
      GCS_TRY(tc::absmax(A, lda, M, K, scratch, st));
      GCS_TRY(tc::absmax(dH, ldh, M, N, scratch + 1, st));
      a_amax = scratch;
      b_amax = scratch + 1;
    }
    if (splits == 1) {
      GCS_TRY(tc::wgrad_launch(A, lda, dH, ldh, dW, M, K, N, 1, per, st, a_amax, b_amax));
    } else {
      GCS_TRY(tc::wgrad_launch(A, lda, dH, ldh, part, M, K, N, splits, per, st, a_amax, b_amax));
      int64_t blocks = ceil_div(kn, 256);
      if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
      split_reduce_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(part, splits, kn, dW);
      GCS_CHECK_LAUNCH("split_reduce_kernel");
    }
  } else if (thin_ok(M, K, N, A, lda, dH) && ldh % 4 == 0 && aligned16(dH) && aligned16(dW) && aligned16(part)) {
    const int splits = thin_splits(M, N);
    const int64_t per = ceil_div(M, splits);
    dim3 grid(static_cast<unsigned>(ceil_div(N, 128)), static_cast<unsigned>(splits));
    float* dst = splits == 1 ? dW : part;
    switch (K / 4) {
      case 1: wgrad_thin_kernel<1><<<grid, 32, 0, st>>>(A, lda, dH, ldh, dst, M, N, per); break;
      case 2: wgrad_thin_kernel<2><<<grid, 64, 0, st>>>(A, lda, dH, ldh, dst, M, N, per); break;
      case 3: wgrad_thin_kernel<3><<<grid, 96, 0, st>>>(A, lda, dH, ldh, dst, M, N, per); break;
      case 4: wgrad_thin_kernel<4><<<grid, 128, 0, st>>>(A, lda, dH, ldh, dst, M, N, per); break;
      case 5: wgrad_thin_kernel<5><<<grid, 160, 0, st>>>(A, lda, dH, ldh, dst, M, N, per); break;
      case 6: wgrad_thin_kernel<6><<<grid, 192, 0, st>>>(A, lda, dH, ldh, dst, M, N, per); break;
      case 7: wgrad_thin_kernel<7><<<grid, 224, 0, st>>>(A, lda, dH, ldh, dst, M, N, per); break;
      default: wgrad_thin_kernel<8><<<grid, 256, 0, st>>>(A, lda, dH, ldh, dst, M, N, per); break;
    }
    GCS_CHECK_LAUNCH("wgrad_thin_kernel");
    if (splits > 1) {
      int64_t blocks = ceil_div(kn, 256);
      if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
      split_reduce_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(part, splits, kn, dW);
      GCS_CHECK_LAUNCH("split_reduce_kernel");
    }
  } else if (w.splits == 1) {
    // C[i=k, j=n] = sum_{r=m} A[m, k] * dH[m, n]
    GCS_TRY((launch_sgemm<false, true>(A, lda, dH, ldh, dW, N, nullptr, K, N, M, 1, w.r_per_split, 0, 0, vec, st, "gcs_linear_bwd_weight")));
  } else {
    GCS_TRY((launch_sgemm<false, true>(A, lda, dH, ldh, part, N, nullptr, K, N, M, w.splits, w.r_per_split, 0, kn, vec, st, "gcs_linear_bwd_weight")));
    int64_t blocks = ceil_div(kn, 256);
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    split_reduce_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(part, w.splits, kn, dW);
    GCS_CHECK_LAUNCH("split_reduce_kernel");
  }
  if (db) {
    double* cws = reinterpret_cast<double*>(static_cast<char*>(workspace) + part_bytes);
    if (N % 4 == 0 && ldh % 4 == 0 && aligned16(dH)) {
      dim3 grid(static_cast<unsigned>(ceil_div(N, 128)), w.col_splits);
      colsum_partial_kernel<4><<<grid, 256, 0, st>>>(dH, ldh, M, N, w.col_rows_per_split, cws);
    } else {
      dim3 grid(static_cast<unsigned>(ceil_div(N, 32)), w.col_splits);
      colsum_partial_kernel<1><<<grid, 256, 0, st>>>(dH, ldh, M, N, w.col_rows_per_split, cws);
    }
    GCS_CHECK_LAUNCH("colsum_partial_kernel");
    colsum_final_kernel<<<static_cast<unsigned>(ceil_div(N, 128)), 128, 0, st>>>(cws, w.col_splits, N, db);
    GCS_CHECK_LAUNCH("colsum_final_kernel");
  }
  return GCS_OK;
}
