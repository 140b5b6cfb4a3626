// K10 — fused multi-tensor optimizer steps over the flat parameter buffer.
//   SGD  (what the reference runs: tf.keras.optimizers.SGD, src/scripts/gcn.py:325,338):
//        w <- w - lr * g
//   Adam (Keras semantics, what the upstream example the script derives from uses and what
//        BASELINE.json's north_star names):
//        m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; lr_t = lr sqrt(1-b2^t)/(1-b1^t)
//        w <- w - lr_t m / (sqrt(v) + eps)            (eps OUTSIDE the corrected root)
// grad_scale folds the 1/world_size (or count weighting) after the NCCL all-reduce in.
#include "common.cuh"

namespace gcs {

__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                  int64_t n, float lr, float grad_scale) {
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x)
    w[k] = w[k] - lr * (g[k] * grad_scale);
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                   float lr_t, float beta1, float beta2, float omb1,
                                                   float omb2, float eps, float grad_scale) {
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float gk = g[k] * grad_scale;
    const float mk = beta1 * m[k] + omb1 * gk;
    const float vk = beta2 * v[k] + omb2 * gk * gk;
    m[k] = mk;
    v[k] = vk;
    w[k] = w[k] - lr_t * mk / (sqrtf(vk) + eps);
  }
}

}  // namespace gcs

using namespace gcs;

static inline unsigned opt_blocks(int64_t n) {
  int64_t b = ceil_div(n, 256);
  const int64_t cap = 8LL * sm_count();
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int gcs_sgd_step(float* w, const float* g, int64_t n, float lr, float grad_scale, gcs_stream stream) {
  GCS_CHECK_ARG(n >= 0 && (n == 0 || (w && g)), "gcs_sgd_step: bad argument");
  if (n == 0) return GCS_OK;
  sgd_kernel<<<opt_blocks(n), 256, 0, as_stream(stream)>>>(w, g, n, lr, grad_scale);
  GCS_CHECK_LAUNCH("sgd_kernel");
  return GCS_OK;
}

extern "C" int gcs_adam_step(float* w, const float* g, float* m, float* v, int64_t n, double lr, double beta1,
                             double beta2, double eps, int64_t step, float grad_scale, gcs_stream stream) {
  GCS_CHECK_ARG(n >= 0 && step >= 1 && (n == 0 || (w && g && m && v)), "gcs_adam_step: bad argument (step is 1-based)");
  if (n == 0) return GCS_OK;
  const double t = static_cast<double>(step);
  // hyper-parameters arrive as doubles so that 1 - beta is rounded once, not computed in fp32
  const float lr_t = static_cast<float>(lr * sqrt(1.0 - pow(beta2, t)) / (1.0 - pow(beta1, t)));
  adam_kernel<<<opt_blocks(n), 256, 0, as_stream(stream)>>>(w, g, m, v, n, lr_t, static_cast<float>(beta1),
                                                            static_cast<float>(beta2), static_cast<float>(1.0 - beta1),
                                                            static_cast<float>(1.0 - beta2), static_cast<float>(eps),
                                                            grad_scale);
  GCS_CHECK_LAUNCH("adam_kernel");
  return GCS_OK;
}
