/* TEST INFRASTRUCTURE: the per-element arithmetic of flownet_ops.cu (csrc/flownet_ops_core.h) compiled for the host by
 * tests/test_flownet_ops_core_cpu.py.  Never linked into the product library. */
#include "flownet_ops_core.h"

void correlation_host(const float* in1, const float* in2, float* out, long long n, int C, int H, int W, int Ho, int Wo,
                      int pad, int K, int md, int s1, int s2)
{
  for (long long i = 0; i < n; ++i) out[i] = cvd_correlation_element(in1, in2, i, C, H, W, Ho, Wo, pad, K, md, s1, s2);
}

void resample2d_host(const float* in1, const float* flow, float* out, long long n, int C, int H, int W)
{
  for (long long i = 0; i < n; ++i) out[i] = cvd_resample2d_element(in1, flow, i, C, H, W);
}

void channelnorm_host(const float* in, float* out, long long n, int C, int H, int W)
{
  for (long long i = 0; i < n; ++i) out[i] = cvd_channelnorm_element(in, i, C, H, W);
}
