// FlowNet2 custom ops for sm_100a: correlation, resample2d, channelnorm -- forward only.
// The reference builds them for sm_50..sm_70 only (third_party/flownet2/networks/*_package/setup.py).  One thread per
// output element over flownet_ops_core.h; these maps are 1/8-resolution feature maps (correlation: 441 x C MACs per
// pixel, < 1 GFLOP per frame pair), so no tiling is attempted.
// Run on a B200 against the oracle (tests/test_flownet_ops_gpu.py); the arithmetic is also checked on the host
// (tests/test_flownet_ops_core_cpu.py).
#include <cuda_runtime.h>

#include "../../include/cvd.h"
#include "cvd_common.cuh"
#include "flownet_ops_core.h"

namespace {

__global__ void correlation_kernel(const float* __restrict__ in1, const float* __restrict__ in2, float* __restrict__ out, long long n,
                                   int C, int H, int W, int Ho, int Wo, int pad, int K, int md, int s1, int s2)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = cvd_correlation_element(in1, in2, i, C, H, W, Ho, Wo, pad, K, md, s1, s2);
}

__global__ void resample2d_kernel(const float* __restrict__ in1, const float* __restrict__ flow, float* __restrict__ out, long long n,
                                  int C, int H, int W)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = cvd_resample2d_element(in1, flow, i, C, H, W);
}

__global__ void channelnorm_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int C, int H, int W)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = cvd_channelnorm_element(in, i, C, H, W);
}

inline unsigned nblocks(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

extern "C" int cvd_correlation_out_size(int H, int W, int pad, int K, int md, int s1, int s2, int* channels, int* Ho, int* Wo)
{
  CVD_CHECK_ARG(channels && Ho && Wo && K >= 1 && (K & 1) && s1 >= 1 && s2 >= 1 && md >= 0 && pad >= 0, "cvd_correlation_out_size: bad arguments");
  const int border = (K - 1) / 2 + md, D = 2 * (md / s2) + 1;
  const int ph = H + 2 * pad - 2 * border, pw = W + 2 * pad - 2 * border;      // correlation_cuda.cc:25-36
  CVD_CHECK_ARG(ph > 0 && pw > 0, "cvd_correlation_out_size: input smaller than the correlation border");
  *channels = D * D; *Ho = (ph + s1 - 1) / s1; *Wo = (pw + s1 - 1) / s1;
  return 0;
}

extern "C" int cvd_correlation_fwd(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                                   int pad, int K, int md, int s1, int s2, void* stream)
{
  CVD_CHECK_ARG(in1 && in2 && out && B > 0 && C > 0 && H > 0 && W > 0, "cvd_correlation_fwd: bad arguments");
  int ch, Ho, Wo;
  if (cvd_correlation_out_size(H, W, pad, K, md, s1, s2, &ch, &Ho, &Wo)) return 1;
  const long long n = (long long)B * ch * Ho * Wo;
  correlation_kernel<<<nblocks(n), 256, 0, (cudaStream_t)stream>>>(in1, in2, out, n, C, H, W, Ho, Wo, pad, K, md, s1, s2);
  CVD_LAUNCH_OK("correlation_kernel");
  return 0;
}

extern "C" int cvd_resample2d_fwd(const float* in1, const float* flow, float* out, int B, int C, int H, int W, void* stream)
{
  CVD_CHECK_ARG(in1 && flow && out && B > 0 && C > 0 && H > 0 && W > 0, "cvd_resample2d_fwd: bad arguments");
  const long long n = (long long)B * C * H * W;
  resample2d_kernel<<<nblocks(n), 256, 0, (cudaStream_t)stream>>>(in1, flow, out, n, C, H, W);
  CVD_LAUNCH_OK("resample2d_kernel");
  return 0;
}

extern "C" int cvd_channelnorm_fwd(const float* in, float* out, int B, int C, int H, int W, void* stream)
{
  CVD_CHECK_ARG(in && out && B > 0 && C > 0 && H > 0 && W > 0, "cvd_channelnorm_fwd: bad arguments");
  const long long n = (long long)B * H * W;
  channelnorm_kernel<<<nblocks(n), 256, 0, (cudaStream_t)stream>>>(in, out, n, C, H, W);
  CVD_LAUNCH_OK("channelnorm_kernel");
  return 0;
}
