// Fused Adam over one flat parameter buffer + the device-side NaN guard.
//
// Replaces optimizer/__init__.py:5-17 (torch.optim.Adam created at
// depth_fine_tuning.py:231-236) .step() (:283) and the host-side
// `if torch.isnan(loss): continue` (:278-280), which in the reference forces a
// D2H sync every iteration.  Update rule follows torch.optim.Adam (no
// weight decay, no amsgrad, eps=1e-8 added AFTER the bias-corrected sqrt):
//   m += (1-b1)(g-m) ; v = b2 v + (1-b2) g g
//   p -= (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// HBM-bound: 16 B read + 12 B written per parameter.
#include "cvd_common.cuh"

namespace {

struct AdamState { int step; int skip; float step_size; float bc2_sqrt; };

__global__ void adam_tick(AdamState* st, const float* loss_flag, float lr, float beta1, float beta2)
{
  const bool bad = loss_flag && isnan(*loss_flag);
  if (bad) { st->skip = 1; return; }
  const int t = st->step + 1;
  st->step = t; st->skip = 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)t);
  const double bc2 = 1.0 - pow((double)beta2, (double)t);
  st->step_size = (float)((double)lr / bc1);
  st->bc2_sqrt = (float)sqrt(bc2);
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v,
                                         float w1, float beta2, float w2, float eps,
                                         float step_size, float bc2s)
{
  m = m + w1 * (g - m);
  v = v * beta2 + (w2 * g) * g;
  const float denom = sqrtf(v) / bc2s + eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, long long n, float beta1, float beta2, float eps,
            float grad_scale, const AdamState* __restrict__ st)
{
  if (st->skip) return;
  const float step_size = st->step_size, bc2s = st->bc2_sqrt;
  const float w1 = 1.0f - beta1, w2 = 1.0f - beta2;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 P = reinterpret_cast<float4*>(p)[i];
    float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    adam_one(P.x, G.x * grad_scale, M.x, V.x, w1, beta2, w2, eps, step_size, bc2s);
    adam_one(P.y, G.y * grad_scale, M.y, V.y, w1, beta2, w2, eps, step_size, bc2s);
    adam_one(P.z, G.z * grad_scale, M.z, V.z, w1, beta2, w2, eps, step_size, bc2s);
    adam_one(P.w, G.w * grad_scale, M.w, V.w, w1, beta2, w2, eps, step_size, bc2s);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
  }
  // tail
  const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    float P = p[t], M = m[t], V = v[t];
    adam_one(P, g[t] * grad_scale, M, V, w1, beta2, w2, eps, step_size, bc2s);
    p[t] = P; m[t] = M; v[t] = V;
  }
}

// loss/parameter_loss.py:13-19 — lambda * sum |p - p0| (L1) and its gradient
__global__ void __launch_bounds__(256)
param_l1_kernel(const float* __restrict__ p, const float* __restrict__ p0, long long n, float lambda,
                float* __restrict__ gacc, float* __restrict__ loss_acc)
{
  __shared__ float red[8];
  float s = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = p[i] - p0[i];
    s += fabsf(d);
    if (gacc) gacc[i] += lambda * ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f));
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(loss_acc, lambda * t);
  }
}

}  // namespace

extern "C" int cvd_adam_flat(float* p, const float* g, float* m, float* v, long long n,
                             float lr, float beta1, float beta2, float eps, float grad_scale,
                             int* step_state, const float* loss_flag, void* stream)
{
  CVD_CHECK_ARG(p && g && m && v && step_state, "cvd_adam_flat: null pointer");
  CVD_CHECK_ARG(n > 0, "cvd_adam_flat: n=%lld", n);
  CVD_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0,
                "cvd_adam_flat: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  adam_tick<<<1, 1, 0, st>>>(reinterpret_cast<AdamState*>(step_state), loss_flag, lr, beta1, beta2);
  CVD_LAUNCH_OK("adam_tick");
  long long blocks = ((n >> 2) + 255) / 256;
  const long long cap = (long long)cvd_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, beta1, beta2, eps, grad_scale,
                                                reinterpret_cast<const AdamState*>(step_state));
  CVD_LAUNCH_OK("adam_kernel");
  return 0;
}

extern "C" int cvd_param_l1(const float* p, const float* p0, long long n, float lambda,
                            float* g_accum, float* out_loss_accum, void* stream)
{
  CVD_CHECK_ARG(p && p0 && out_loss_accum, "cvd_param_l1: null pointer");
  CVD_CHECK_ARG(n > 0, "cvd_param_l1: n=%lld", n);
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)cvd_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  param_l1_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, p0, n, lambda, g_accum, out_loss_accum);
  CVD_LAUNCH_OK("param_l1_kernel");
  return 0;
}
