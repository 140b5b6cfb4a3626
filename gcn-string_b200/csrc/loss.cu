// K5 — softmax + CategoricalCrossentropy (mean) + categorical_accuracy, forward and the
// gradient w.r.t. the logits (reference: src/scripts/gcn.py:326 loss_fn, :335 loss,
// :339 accuracy).  Keras recovers the logits of the softmax activation
// and evaluates softmax_cross_entropy_with_logits:
//   loss_b = logsumexp(z_b) * sum(y_b) - y_b . z_b ;  loss = mean_b loss_b
//          (evaluated as log(sum exp(z - m)) * sum(y) - y . (z - m), m = max z)
//   dz_b   = (softmax(z_b) * sum(y_b) - y_b) * grad_scale
// One CTA, fixed-order tree reduction: deterministic.  B x C is tiny ([1024, 2]).
#include "common.cuh"

namespace gcs {

constexpr int kMaxClasses = 64;

__global__ void __launch_bounds__(1024) softmax_xent_kernel(
    const float* __restrict__ logits, const float* __restrict__ y, int B, int C,
    float* __restrict__ probs, float* __restrict__ loss_acc, float* __restrict__ dlogits,
    float grad_scale) {
  __shared__ double s_loss[1024];
  __shared__ int s_hit[1024];
  double loss = 0.0;
  int hit = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* z = logits + static_cast<int64_t>(b) * C;
    float m = z[0];
    int arg_p = 0;
    for (int c = 1; c < C; ++c)
      if (z[c] > m) { m = z[c]; arg_p = c; }       // first maximum, like tf.argmax
    float e[kMaxClasses];
    float rest = 0.f;                                // sum over c != argmax; e[argmax] = 1 exactly
    for (int c = 0; c < C; ++c) {
      e[c] = expf(z[c] - m);
      if (c != arg_p) rest += e[c];
    }
    const float sum = 1.0f + rest;
    const float log_sum = log1pf(rest);              // lse = m + log_sum, accurate for confident rows
    const float inv = 1.0f / sum;
    float ysum = 0.f, yz = 0.f, ymax = 0.f;
    int arg_y = 0;
    if (y) {
      const float* yy = y + static_cast<int64_t>(b) * C;
      ymax = yy[0];
      for (int c = 0; c < C; ++c) {
        ysum += yy[c];
        yz = fmaf(yy[c], z[c] - m, yz);            // y . (z - m): no cancellation against lse
        if (c > 0 && yy[c] > ymax) { ymax = yy[c]; arg_y = c; }
      }
      loss += static_cast<double>(log_sum * ysum - yz);
      hit += (arg_y == arg_p);
    }
    for (int c = 0; c < C; ++c) {
      const float p = e[c] * inv;
      if (probs) probs[static_cast<int64_t>(b) * C + c] = p;
      if (dlogits) dlogits[static_cast<int64_t>(b) * C + c] = (p * ysum - y[static_cast<int64_t>(b) * C + c]) * grad_scale;
    }
  }
  s_loss[threadIdx.x] = loss;
  s_hit[threadIdx.x] = hit;
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      s_loss[threadIdx.x] += s_loss[threadIdx.x + off];
      s_hit[threadIdx.x] += s_hit[threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && loss_acc) {
    loss_acc[0] = B > 0 ? static_cast<float>(s_loss[0] / B) : 0.f;
    loss_acc[1] = B > 0 ? static_cast<float>(s_hit[0]) / static_cast<float>(B) : 0.f;
  }
}

}  // namespace gcs

using namespace gcs;

extern "C" int gcs_softmax_xent(const float* logits, const float* y, int32_t B, int32_t C, float* probs,
                                float* loss_acc, float* dlogits, float grad_scale, gcs_stream stream) {
  GCS_CHECK_ARG(B >= 0 && C > 0 && C <= kMaxClasses, "gcs_softmax_xent: C=%d outside [1, %d]", C, kMaxClasses);
  GCS_CHECK_ARG(logits || B == 0, "gcs_softmax_xent: null logits");
  GCS_CHECK_ARG(y || (!dlogits && !loss_acc), "gcs_softmax_xent: loss / gradient requested without labels");
  softmax_xent_kernel<<<1, 1024, 0, as_stream(stream)>>>(logits, y, B, C, probs, loss_acc, dlogits, grad_scale);
  GCS_CHECK_LAUNCH("softmax_xent_kernel");
  return GCS_OK;
}
