// Element-wise kernels of the MiDaS-v2 network around the tcgen05 conv engine; NHWC fp32.
//
//   (x - mean) / std per channel            midas_v2_model.py:58-59 (ImageNet statistics on the BGR input as is)
//   relu(x) + other                          blocks.py:111-117 ResidualConvUnit skip (nn.ReLU(inplace=True) rewrites x),
//                                            blocks.py:146-147 FeatureFusionBlock sum
//   bilinear x2, align_corners True / False  blocks.py:151-153 / midas_net.py:40 (Interpolate)
//   depth = 1 / relu(out)                    midas_net.py:43, midas_v2_model.py:67
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cvd.h"
#include "cvd_common.cuh"

namespace {

constexpr int kT = 256;
inline unsigned blocks_for(long long n) { long long b = (n + kT - 1) / kT; return (unsigned)(b < 1 ? 1 : b); }

__global__ void image_normalize_kernel(const float* __restrict__ img, int N, int H, int W, float m0, float m1, float m2,
                                       float s0, float s1, float s2, float4* __restrict__ out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long hw = (long long)H * W;
  if (i >= N * hw) return;
  const long long n = i / hw, p = i - n * hw;
  const float* b = img + n * 3 * hw + p;
  out[i] = make_float4((__ldg(b) - m0) * s0, (__ldg(b + hw) - m1) * s1, (__ldg(b + 2 * hw) - m2) * s2, 0.f);
}

__global__ void relu_add_kernel(const float4* __restrict__ x, const float4* __restrict__ other, float4* __restrict__ out, long long n4)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = __ldg(x + i);
  v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  if (other) { const float4 o = __ldg(other + i); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
  out[i] = v;
}

// source coordinate of destination index d for a x2 bilinear resize (torch upsample_bilinear2d)
__device__ __forceinline__ void bil_taps(int d, int n_in, int n_out, int align, int& i0, int& i1, float& t)
{
  float src;
  if (align) src = n_out > 1 ? (float)d * ((float)(n_in - 1) / (float)(n_out - 1)) : 0.f;
  else { src = ((float)d + 0.5f) * 0.5f - 0.5f; if (src < 0.f) src = 0.f; }
  i0 = (int)src;
  if (i0 > n_in - 1) i0 = n_in - 1;
  i1 = i0 + 1 < n_in ? i0 + 1 : n_in - 1;
  t = src - (float)i0;
}

__global__ void up2_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ r, int relu_r, int N, int h, int w, int C4,
                               int align, float4* __restrict__ out)
{
  const int H = 2 * h, W = 2 * w;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * H * W * C4) return;
  const int c = (int)(i % C4); long long p = i / C4;
  const int ox = (int)(p % W), oy = (int)((p / W) % H), n = (int)(p / ((long long)W * H));
  int y0, y1, x0, x1; float ty, tx;
  bil_taps(oy, h, H, align, y0, y1, ty);
  bil_taps(ox, w, W, align, x0, x1, tx);
  const float4* b = x + (size_t)n * h * w * C4 + c;
  const float4 v00 = __ldg(b + ((size_t)y0 * w + x0) * C4), v01 = __ldg(b + ((size_t)y0 * w + x1) * C4);
  const float4 v10 = __ldg(b + ((size_t)y1 * w + x0) * C4), v11 = __ldg(b + ((size_t)y1 * w + x1) * C4);
  const float w00 = (1.f - ty) * (1.f - tx), w01 = (1.f - ty) * tx, w10 = ty * (1.f - tx), w11 = ty * tx;
  float4 o = make_float4(w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x, w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y,
                         w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z, w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w);
  if (r) {
    float4 q = __ldg(r + i);
    if (relu_r) { q.x = fmaxf(q.x, 0.f); q.y = fmaxf(q.y, 0.f); q.z = fmaxf(q.z, 0.f); q.w = fmaxf(q.w, 0.f); }
    o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
  }
  out[i] = o;
}

// transpose in gather form: source pixel s collects from the destination pixels whose two taps include it
__device__ __forceinline__ int up2_weights(int s, int n_in, int n_out, int align, int d_first[1], float wts[8])
{
  int lo = 2 * s - 3; if (lo < 0) lo = 0;
  int hi = 2 * s + 4; if (hi > n_out - 1) hi = n_out - 1;
  d_first[0] = lo;
  int n = 0;
  for (int d = lo; d <= hi; ++d, ++n) {
    int i0, i1; float t;
    bil_taps(d, n_in, n_out, align, i0, i1, t);
    wts[n] = (i0 == s ? 1.f - t : 0.f) + (i1 == s ? t : 0.f);
  }
  return n;
}

__global__ void up2_bwd_kernel(const float4* __restrict__ dout, int N, int h, int w, int C4, int align, float4* __restrict__ dx, int accumulate)
{
  const int H = 2 * h, W = 2 * w;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * h * w * C4) return;
  const int c = (int)(i % C4); long long p = i / C4;
  const int sx = (int)(p % w), sy = (int)((p / w) % h), n = (int)(p / ((long long)w * h));
  int fy[1], fx[1]; float wy[8], wx[8];
  const int ny = up2_weights(sy, h, H, align, fy, wy), nx = up2_weights(sx, w, W, align, fx, wx);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* b = dout + (size_t)n * H * W * C4 + c;
  for (int a = 0; a < ny; ++a) {
    if (wy[a] == 0.f) continue;
    for (int e = 0; e < nx; ++e) {
      const float wgt = wy[a] * wx[e];
      if (wgt == 0.f) continue;
      const float4 g = __ldg(b + ((size_t)(fy[0] + a) * W + fx[0] + e) * C4);
      acc.x += wgt * g.x; acc.y += wgt * g.y; acc.z += wgt * g.z; acc.w += wgt * g.w;
    }
  }
  if (accumulate) { const float4 o = dx[i]; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
  dx[i] = acc;
}

__global__ void recip_relu_fwd_kernel(const float4* __restrict__ raw4, float* __restrict__ depth, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  depth[i] = 1.f / fmaxf(__ldg(raw4 + i).x, 0.f);
}

__global__ void recip_relu_bwd_kernel(const float* __restrict__ ddepth, const float* __restrict__ depth, const float4* __restrict__ raw4,
                                      float4* __restrict__ draw4, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = depth[i];
  const float g = __ldg(raw4 + i).x > 0.f ? -ddepth[i] * d * d : 0.f;     // threshold_backward: exactly 0 where clamped
  draw4[i] = make_float4(g, 0.f, 0.f, 0.f);
}

}  // namespace

extern "C" int cvd_image_normalize_nhwc4(const float* img_nchw, int N, int H, int W, const float* mean3_host,
                                         const float* std3_host, float* out_nhwc4, void* stream)
{
  CVD_CHECK_ARG(img_nchw && out_nhwc4 && mean3_host && std3_host && N > 0 && H > 0 && W > 0, "cvd_image_normalize_nhwc4: bad arguments");
  image_normalize_kernel<<<blocks_for((long long)N * H * W), kT, 0, (cudaStream_t)stream>>>(
      img_nchw, N, H, W, mean3_host[0], mean3_host[1], mean3_host[2], 1.f / std3_host[0], 1.f / std3_host[1], 1.f / std3_host[2],
      reinterpret_cast<float4*>(out_nhwc4));
  CVD_LAUNCH_OK("image_normalize_kernel");
  return 0;
}

extern "C" int cvd_relu_add(const float* x, const float* other, float* out, long long n, void* stream)
{
  CVD_CHECK_ARG(x && out && n > 0 && (n & 3) == 0, "cvd_relu_add: bad arguments");
  relu_add_kernel<<<blocks_for(n / 4), kT, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(other),
                                                                      reinterpret_cast<float4*>(out), n / 4);
  CVD_LAUNCH_OK("relu_add_kernel");
  return 0;
}

extern "C" int cvd_up2_bilinear_fwd(const float* x, const float* r, int relu_r, int N, int h, int w, int C, int align_corners,
                                    float* out, void* stream)
{
  CVD_CHECK_ARG(x && out && N > 0 && h > 0 && w > 0 && C > 0 && (C & 3) == 0, "cvd_up2_bilinear_fwd: bad arguments");
  up2_fwd_kernel<<<blocks_for((long long)N * 4 * h * w * (C / 4)), kT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(r), relu_r, N, h, w, C / 4, align_corners, reinterpret_cast<float4*>(out));
  CVD_LAUNCH_OK("up2_fwd_kernel");
  return 0;
}

extern "C" int cvd_up2_bilinear_bwd(const float* dout, int N, int h, int w, int C, int align_corners, float* dx, int accumulate, void* stream)
{
  CVD_CHECK_ARG(dout && dx && N > 0 && h > 0 && w > 0 && C > 0 && (C & 3) == 0, "cvd_up2_bilinear_bwd: bad arguments");
  up2_bwd_kernel<<<blocks_for((long long)N * h * w * (C / 4)), kT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(dout), N, h, w, C / 4, align_corners, reinterpret_cast<float4*>(dx), accumulate);
  CVD_LAUNCH_OK("up2_bwd_kernel");
  return 0;
}

extern "C" int cvd_recip_relu_fwd(const float* raw4, float* depth, long long n, void* stream)
{
  CVD_CHECK_ARG(raw4 && depth && n > 0, "cvd_recip_relu_fwd: bad arguments");
  recip_relu_fwd_kernel<<<blocks_for(n), kT, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(raw4), depth, n);
  CVD_LAUNCH_OK("recip_relu_fwd_kernel");
  return 0;
}

extern "C" int cvd_recip_relu_bwd(const float* ddepth, const float* depth, const float* raw4, float* draw4, long long n, void* stream)
{
  CVD_CHECK_ARG(ddepth && depth && raw4 && draw4 && n > 0, "cvd_recip_relu_bwd: bad arguments");
  recip_relu_bwd_kernel<<<blocks_for(n), kT, 0, (cudaStream_t)stream>>>(ddepth, depth, reinterpret_cast<const float4*>(raw4),
                                                                        reinterpret_cast<float4*>(draw4), n);
  CVD_LAUNCH_OK("recip_relu_bwd_kernel");
  return 0;
}
