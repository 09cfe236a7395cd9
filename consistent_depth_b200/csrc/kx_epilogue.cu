// Shifted sum that completes the kx-fused forward convolution (conv_col.cu):
//     out[n, y, x, co] = bias[co] + sum_kx D[n, y, x + kx, kx*Cout + co]
// D (N, H, W + k - 1, k*Cout) is the output of the k x 1 column conv with N = k*Cout GEMM columns.
// EXPERIMENTAL (CVD_KXFWD=1).  HBM-bound: reads D once (k*Cout*4 B per window slot), writes Cout*4 B per pixel.
#include <cuda_runtime.h>

#include "../../include/cvd.h"
#include "cvd_common.cuh"

namespace {

__global__ void shift_sum_kernel(const float* __restrict__ D, int Dct, const float* __restrict__ bias, float* __restrict__ out,
                                 int o_ct, int o_c0, int N, int H, int W, int k, int cout)
{
  const int C4 = cout >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * H * W * C4) return;
  const int c = (int)(i % C4); long long p = i / C4;
  const int x = (int)(p % W); const long long row = p / W;              // row = n * H + y
  const int Wd = W + k - 1;
  const float* base = D + ((size_t)row * Wd + x) * Dct + c * 4;
  float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias + c * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int kx = 0; kx < k; ++kx) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)kx * Dct + kx * cout));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(out + (size_t)p * o_ct + o_c0 + c * 4) = acc;
}

}  // namespace

extern "C" int cvd_shift_sum(const float* D, int d_ctotal, const float* bias, float* out, int o_ctotal, int o_coff,
                             int N, int H, int W, int k, int cout, void* stream)
{
  CVD_CHECK_ARG(D && out && N > 0 && H > 0 && W > 0 && k >= 1, "cvd_shift_sum: bad arguments");
  CVD_CHECK_ARG((cout & 3) == 0 && cout > 0 && d_ctotal >= k * cout && (d_ctotal & 3) == 0 && (o_ctotal & 3) == 0 && (o_coff & 3) == 0,
                "cvd_shift_sum: channel counts / strides must be multiples of 4");
  CVD_CHECK_ARG(!bias || ((size_t)bias & 15) == 0, "cvd_shift_sum: bias must be 16-byte aligned");
  const long long n = (long long)N * H * W * (cout / 4);
  shift_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(D, d_ctotal, bias, out, o_ctotal, o_coff, N, H, W, k, cout);
  CVD_LAUNCH_OK("shift_sum_kernel");
  return 0;
}
