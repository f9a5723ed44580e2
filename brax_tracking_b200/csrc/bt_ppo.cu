// Learner-side helper of the PPO loop that drives the step: the tanh-Normal policy terms of
// brax's compute_ppo_loss -- log-probability of the stored raw action and the sampled-entropy estimate -- as ONE pass over
// the logits forward and ONE pass backward, instead of ~60 elementwise launches over [rows, action] tensors.
// Reference: custom_brax/custom_ppo.py:250-284 -> brax.training.agents.ppo.losses.compute_ppo_loss,
// brax.training.distribution.NormalTanhDistribution (min_std = 0.001).
#include <cuda_runtime.h>
#include <stdint.h>

#include "bt_api.h"

namespace {
constexpr float kMinStd = 0.001f;
constexpr float kHalfLog2Pi = 0.91893853320467274178f;
constexpr float kLog2 = 0.69314718055994530942f;

__device__ __forceinline__ float softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // torch's threshold
__device__ __forceinline__ float sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
// log |d tanh(x) / dx| = 2 (log 2 - x - softplus(-2x))
__device__ __forceinline__ float log_det_jac(float x) { return 2.f * (kLog2 - x - softplus(-2.f * x)); }
__device__ __forceinline__ float log_det_jac_grad(float x) { return 2.f * (2.f * sigmoid(-2.f * x) - 1.f); }

struct Row { int b, t; };
__device__ __forceinline__ Row row_of(int64_t r, int T) { Row o; o.b = (int)(r / T); o.t = (int)(r - (int64_t)o.b * T); return o; }

// one warp per (b, t) row; logits [B, T, 2A] contiguous, raw / noise addressed by (b, t) strides, outputs by (b, t) strides
__global__ void __launch_bounds__(256) k_tanh_normal_fwd(int B, int T, int A, const float* __restrict__ logits,
                                                         const float* __restrict__ raw, int64_t raw_sb, int64_t raw_st,
                                                         const float* __restrict__ noise, int64_t noise_sb, int64_t noise_st,
                                                         float* __restrict__ lp, float* __restrict__ ent, int64_t out_sb, int64_t out_st) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= (int64_t)B * T) return;
  const Row rt = row_of(r, T);
  const float* lg = logits + r * 2 * A;
  const float* ra = raw + rt.b * raw_sb + rt.t * raw_st;
  const float* nz = noise + rt.b * noise_sb + rt.t * noise_st;
  float slp = 0.f, sent = 0.f;
  for (int a = lane; a < A; a += 32) {
    const float loc = lg[a], scale = softplus(lg[A + a]) + kMinStd;
    const float x = ra[a], z = (x - loc) / scale, ls = logf(scale);
    slp += -0.5f * z * z - ls - kHalfLog2Pi - log_det_jac(x);
    sent += 0.5f + kHalfLog2Pi + ls + log_det_jac(loc + scale * nz[a]);
  }
  for (int o = 16; o > 0; o >>= 1) { slp += __shfl_xor_sync(0xffffffffu, slp, o); sent += __shfl_xor_sync(0xffffffffu, sent, o); }
  if (lane == 0) { lp[rt.b * out_sb + rt.t * out_st] = slp; ent[rt.b * out_sb + rt.t * out_st] = sent; }
}

__global__ void __launch_bounds__(256) k_tanh_normal_bwd(int B, int T, int A, const float* __restrict__ logits,
                                                         const float* __restrict__ raw, int64_t raw_sb, int64_t raw_st,
                                                         const float* __restrict__ noise, int64_t noise_sb, int64_t noise_st,
                                                         const float* __restrict__ glp, const float* __restrict__ gent, int64_t out_sb,
                                                         int64_t out_st, float* __restrict__ glogits) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= (int64_t)B * T) return;
  const Row rt = row_of(r, T);
  const float* lg = logits + r * 2 * A;
  const float* ra = raw + rt.b * raw_sb + rt.t * raw_st;
  const float* nz = noise + rt.b * noise_sb + rt.t * noise_st;
  const float gl = glp[rt.b * out_sb + rt.t * out_st], ge = gent[rt.b * out_sb + rt.t * out_st];
  float* go = glogits + r * 2 * A;
  for (int a = lane; a < A; a += 32) {
    const float loc = lg[a], s = lg[A + a], scale = softplus(s) + kMinStd, inv = 1.f / scale;
    const float z = (ra[a] - loc) * inv, n = nz[a], dj = log_det_jac_grad(loc + scale * n);
    go[a] = gl * z * inv + ge * dj;
    const float dscale = gl * (z * z - 1.f) * inv + ge * (inv + dj * n);
    go[A + a] = dscale * (s > 20.f ? 1.f : sigmoid(s));
  }
}
// optax.adam over ONE flat parameter buffer, fused with the 1 / world scale that completes lax.pmean of the all-reduced (summed)
// gradient (custom_brax/custom_ppo.py:246-257): p, m, v, g are [n]; `step` is a device-resident float holding the number of
// updates already applied (read by every thread, advanced by the caller on the stream after the launch), so that the launch
// is capturable in a CUDA graph.  optax semantics: m_hat = m / (1 - b1^t), v_hat = v / (1 - b2^t), p -= lr m_hat / (sqrt(v_hat) + eps).
__global__ void __launch_bounds__(256) k_flat_adam(int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, const float* __restrict__ step, float lr, float b1, float b2,
                                                   float eps, float gscale) {
  const float t = *step + 1.f;
  const float c1 = 1.f / (1.f - powf(b1, t)), c2 = 1.f / (1.f - powf(b2, t));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr * (mi * c1) / (sqrtf(vi * c2) + eps);
  }
}
}  // namespace

extern "C" {
int bt_ppo_flat_adam(int64_t n, float* p, const float* g, float* m, float* v, const float* step, float lr, float b1, float b2,
                     float eps, float gscale, void* stream) {
  if (n < 0 || !p || !g || !m || !v || !step) return BT_E_ARG;
  if (n == 0) return BT_OK;
  const int64_t blocks = (n + 255) / 256;
  k_flat_adam<<<(unsigned)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, (cudaStream_t)stream>>>(n, p, g, m, v, step, lr, b1, b2, eps, gscale);
  return cudaGetLastError() == cudaSuccess ? BT_OK : BT_E_CUDA;
}

int bt_ppo_tanh_normal_fwd(int B, int T, int A, const float* logits, const float* raw, int64_t raw_sb, int64_t raw_st,
                           const float* noise, int64_t noise_sb, int64_t noise_st, float* lp, float* ent, int64_t out_sb,
                           int64_t out_st, void* stream) {
  if (B < 0 || T < 0 || A < 1 || !logits || !raw || !noise || !lp || !ent) return BT_E_ARG;
  const int64_t rows = (int64_t)B * T;
  if (rows == 0) return BT_OK;
  k_tanh_normal_fwd<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(B, T, A, logits, raw, raw_sb, raw_st, noise, noise_sb,
                                                                                 noise_st, lp, ent, out_sb, out_st);
  return cudaGetLastError() == cudaSuccess ? BT_OK : BT_E_CUDA;
}

int bt_ppo_tanh_normal_bwd(int B, int T, int A, const float* logits, const float* raw, int64_t raw_sb, int64_t raw_st,
                           const float* noise, int64_t noise_sb, int64_t noise_st, const float* glp, const float* gent,
                           int64_t out_sb, int64_t out_st, float* glogits, void* stream) {
  if (B < 0 || T < 0 || A < 1 || !logits || !raw || !noise || !glp || !gent || !glogits) return BT_E_ARG;
  const int64_t rows = (int64_t)B * T;
  if (rows == 0) return BT_OK;
  k_tanh_normal_bwd<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(B, T, A, logits, raw, raw_sb, raw_st, noise, noise_sb,
                                                                                 noise_st, glp, gent, out_sb, out_st, glogits);
  return cudaGetLastError() == cudaSuccess ? BT_OK : BT_E_CUDA;
}
}
