// Element-wise / gather kernels of the monodepth2 backbone around the tcgen05 conv engine.
// All activations are NHWC fp32.  HBM-bound passes: one thread per (pixel, 4-channel group), 128-bit accesses.
//
//   bicubic resize (align_corners=False, A=-0.75)   monodepth2_model.py:72-74 (image in), :79-80 (disparity out)
//   (x-0.45)/0.225                                  resnet_encoder.py:89 (fused into the image resize)
//   depth = 1/disp                                   monodepth2_model.py:82
//   sigmoid                                          depth_decoder.py:63
//   MaxPool2d(3, 2, 1)                               torchvision resnet18 (resnet_encoder.py:93)
//   stride-2 sub-sampling / zero-stuffing            stride-2 convs = stride-1 conv at the input resolution + pick (2y,2x)
//   relu(bn2(conv2) + identity)                      torchvision BasicBlock.forward
//   ReflectionPad2d(1) (+ ELU) (+ nearest x2) (+ cat) layers.py:106-136,196-199, depth_decoder.py:53-60, as ONE gather into
//                                                    the padded input buffer of the next 3x3 conv
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cvd.h"
#include "cvd_common.cuh"

namespace {

constexpr int kT = 256;
inline unsigned blocks_for(long long n) { long long b = (n + kT - 1) / kT; return (unsigned)(b < 1 ? 1 : b); }

// ---------------------------------------------------------------- bicubic
__device__ __forceinline__ void cubic_taps(int o, float scale, int n_in, int idx[4], float w[4])
{
  // torch upsample_bicubic2d: src = scale * (dst + 0.5) - 0.5 (not clamped), taps floor-1 .. floor+2 clamped to the border
  const float A = -0.75f;
  const float src = scale * ((float)o + 0.5f) - 0.5f;
  const float fl = floorf(src);
  const float t = src - fl;
  const int i0 = (int)fl;
  const float x0 = t + 1.f, x3 = 2.f - t, x2 = 1.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
#pragma unroll
  for (int k = 0; k < 4; ++k) { int i = i0 - 1 + k; idx[k] = i < 0 ? 0 : (i > n_in - 1 ? n_in - 1 : i); }
}

__global__ void bicubic_image_kernel(const float* __restrict__ img /* N,3,H,W */, int N, int H, int W,
                                     float* __restrict__ out /* N,oh,ow,4 */, int oh, int ow, float sub, float mul)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * oh * ow) return;
  const int ox = (int)(i % ow), oy = (int)((i / ow) % oh), n = (int)(i / ((long long)ow * oh));
  int iy[4], ix[4]; float wy[4], wx[4];
  cubic_taps(oy, (float)H / (float)oh, H, iy, wy);
  cubic_taps(ox, (float)W / (float)ow, W, ix, wx);
  float r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* pl = img + ((size_t)n * 3 + c) * H * W;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float* row = pl + (size_t)iy[a] * W;
      acc += wy[a] * (wx[0] * __ldg(row + ix[0]) + wx[1] * __ldg(row + ix[1]) + wx[2] * __ldg(row + ix[2]) + wx[3] * __ldg(row + ix[3]));
    }
    r[c] = (acc - sub) * mul;
  }
  reinterpret_cast<float4*>(out)[i] = make_float4(r[0], r[1], r[2], 0.f);
}

__global__ void disp_resize_fwd_kernel(const float* __restrict__ disp /* N,fh,fw */, int N, int fh, int fw,
                                       float* __restrict__ depth /* N,H,W */, int H, int W)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * H * W) return;
  const int ox = (int)(i % W), oy = (int)((i / W) % H), n = (int)(i / ((long long)W * H));
  int iy[4], ix[4]; float wy[4], wx[4];
  cubic_taps(oy, (float)fh / (float)H, fh, iy, wy);
  cubic_taps(ox, (float)fw / (float)W, fw, ix, wx);
  const float* pl = disp + (size_t)n * fh * fw;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float* row = pl + (size_t)iy[a] * fw;
    acc += wy[a] * (wx[0] * __ldg(row + ix[0]) + wx[1] * __ldg(row + ix[1]) + wx[2] * __ldg(row + ix[2]) + wx[3] * __ldg(row + ix[3]));
  }
  depth[i] = 1.f / acc;
}

__global__ void disp_resize_bwd_kernel(const float* __restrict__ ddepth, const float* __restrict__ depth, int N, int fh, int fw,
                                       int H, int W, float* __restrict__ ddisp /* zeroed */)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * H * W) return;
  const int ox = (int)(i % W), oy = (int)((i / W) % H), n = (int)(i / ((long long)W * H));
  int iy[4], ix[4]; float wy[4], wx[4];
  cubic_taps(oy, (float)fh / (float)H, fh, iy, wy);
  cubic_taps(ox, (float)fw / (float)W, fw, ix, wx);
  const float d = depth[i];
  const float g = -ddepth[i] * d * d;                    // depth = 1/r  =>  d depth / d r = -depth^2
  float* pl = ddisp + (size_t)n * fh * fw;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) atomicAdd(pl + (size_t)iy[a] * fw + ix[b], wy[a] * wx[b] * g);
}

// raw: dispconv output on the (fh+2) x (fw+2) padded grid, c_total channels, channel 0 used
__global__ void sigmoid_fwd_kernel(const float* __restrict__ raw, int ct, int N, int fh, int fw, float* __restrict__ disp)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * fh * fw) return;
  const int x = (int)(i % fw), y = (int)((i / fw) % fh), n = (int)(i / ((long long)fw * fh));
  const float r = raw[(((size_t)n * (fh + 2) + y + 1) * (fw + 2) + x + 1) * ct];
  disp[i] = 1.f / (1.f + expf(-r));
}

__global__ void sigmoid_bwd_kernel(const float* __restrict__ ddisp, const float* __restrict__ disp, int N, int fh, int fw,
                                   float* __restrict__ draw, int ct)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * fh * fw) return;
  const int x = (int)(i % fw), y = (int)((i / fw) % fh), n = (int)(i / ((long long)fw * fh));
  const float s = disp[i];
  draw[(((size_t)n * (fh + 2) + y + 1) * (fw + 2) + x + 1) * ct] = ddisp[i] * s * (1.f - s);
}

// ---------------------------------------------------------------- stride-2 helpers
__global__ void subsample2_kernel(const float4* __restrict__ src, int N, int H, int W, int C4, float4* __restrict__ dst, int oh, int ow)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * oh * ow * C4) return;
  const int c = (int)(i % C4); long long p = i / C4;
  const int x = (int)(p % ow), y = (int)((p / ow) % oh), n = (int)(p / ((long long)ow * oh));
  dst[i] = __ldg(src + (((size_t)n * H + 2 * y) * W + 2 * x) * C4 + c);
}

__global__ void stuff2_kernel(const float4* __restrict__ src, int N, int oh, int ow, int C4, float4* __restrict__ dst, int H, int W, int accumulate)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * oh * ow * C4) return;
  const int c = (int)(i % C4); long long p = i / C4;
  const int x = (int)(p % ow), y = (int)((p / ow) % oh), n = (int)(p / ((long long)ow * oh));
  float4* d = dst + (((size_t)n * H + 2 * y) * W + 2 * x) * C4 + c;
  float4 v = __ldg(src + i);
  if (accumulate) { const float4 o = *d; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
  *d = v;
}

// BatchNorm(+ReLU) backward applied element-wise (the conv kernels do the same on load; a stride-2 conv's gradient
// has to be materialised because the zero-stuffed positions must stay exactly zero):
//   yh = a x + b ; g = dy * [relu ? yh > 0 : 1] ; dx = c0 g - c1 - c2 yh     (bw = float4 c0,c1,c2,-)
__global__ void bnbwd_stuff2_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, const float* __restrict__ a,
                                    const float* __restrict__ b, const float4* __restrict__ bw, int relu,
                                    int N, int oh, int ow, int C4, float4* __restrict__ dst, int H, int W, int stride)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * oh * ow * C4) return;
  const int c = (int)(i % C4); long long p = i / C4;
  const int xx = (int)(p % ow), yy = (int)((p / ow) % oh), n = (int)(p / ((long long)ow * oh));
  const float4 xv = __ldg(x + i), gv = __ldg(dy + i);
  const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
  float r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ch = c * 4 + j;
    const float yh = fmaf(__ldg(a + ch), xs[j], __ldg(b + ch));
    const float g = (!relu || yh > 0.f) ? gs[j] : 0.f;
    const float4 q = __ldg(bw + ch);
    r[j] = q.x * g - q.y - q.z * yh;
  }
  dst[(((size_t)n * H + stride * yy) * W + stride * xx) * C4 + c] = make_float4(r[0], r[1], r[2], r[3]);
}

// ---------------------------------------------------------------- MaxPool2d(3, 2, 1) over relu(a x + b)
__device__ __forceinline__ float4 xf_affine_relu(float4 v, const float* a, const float* b, int ch, int relu)
{
  if (a) {
    v.x = fmaf(__ldg(a + ch), v.x, __ldg(b + ch)); v.y = fmaf(__ldg(a + ch + 1), v.y, __ldg(b + ch + 1));
    v.z = fmaf(__ldg(a + ch + 2), v.z, __ldg(b + ch + 2)); v.w = fmaf(__ldg(a + ch + 3), v.w, __ldg(b + ch + 3));
  }
  if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  return v;
}

__global__ void maxpool_fwd_kernel(const float4* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b, int relu,
                                   int N, int H, int W, int C4, float4* __restrict__ out, uchar4* __restrict__ idx, int oh, int ow)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * oh * ow * C4) return;
  const int c = (int)(i % C4); long long p = i / C4;
  const int ox = (int)(p % ow), oy = (int)((p / ow) % oh), n = (int)(p / ((long long)ow * oh));
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  unsigned char am[4] = {0, 0, 0, 0};
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const float4 v = xf_affine_relu(__ldg(x + (((size_t)n * H + iy) * W + ix) * C4 + c), a, b, c * 4, relu);
      const float vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (vs[j] > m[j] || vs[j] != vs[j]) { m[j] = vs[j]; am[j] = (unsigned char)(ky * 3 + kx); }   // first max wins (torch)
    }
  out[i] = make_float4(m[0], m[1], m[2], m[3]);
  idx[i] = make_uchar4(am[0], am[1], am[2], am[3]);
}

// gather form: input pixel (iy, ix) collects from the <= 4 windows whose recorded arg-max is this pixel
__global__ void maxpool_bwd_kernel(const float4* __restrict__ dout, const uchar4* __restrict__ idx, int N, int H, int W, int C4,
                                   int oh, int ow, float4* __restrict__ dx, int accumulate)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * H * W * C4) return;
  const int c = (int)(i % C4); long long p = i / C4;
  const int ix = (int)(p % W), iy = (int)((p / W) % H), n = (int)(p / ((long long)W * H));
  float r[4] = {0.f, 0.f, 0.f, 0.f};
  for (int oy = iy >> 1; oy <= (iy + 1) >> 1; ++oy) {          // windows with 2 oy - 1 <= iy <= 2 oy + 1
    if (oy >= oh) continue;
    const int ky = iy - (2 * oy - 1);
    for (int ox = ix >> 1; ox <= (ix + 1) >> 1; ++ox) {
      if (ox >= ow) continue;
      const int tap = ky * 3 + ix - (2 * ox - 1);
      const size_t o = (((size_t)n * oh + oy) * ow + ox) * C4 + c;
      const uchar4 am = __ldg(idx + o);
      const float4 g = __ldg(dout + o);
      if (am.x == tap) r[0] += g.x;
      if (am.y == tap) r[1] += g.y;
      if (am.z == tap) r[2] += g.z;
      if (am.w == tap) r[3] += g.w;
    }
  }
  float4 v = make_float4(r[0], r[1], r[2], r[3]);
  if (accumulate) { const float4 o = dx[i]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
  dx[i] = v;
}

// ---------------------------------------------------------------- residual add
// out = relu(a y + b + (ra ? ra r + rb : r))
__global__ void bn_add_relu_kernel(const float4* __restrict__ y, const float* __restrict__ a, const float* __restrict__ b,
                                   const float4* __restrict__ r, const float* __restrict__ ra, const float* __restrict__ rb,
                                   long long n4, int C4, float4* __restrict__ out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int ch = (int)(i % C4) * 4;
  float4 v = xf_affine_relu(__ldg(y + i), a, b, ch, 0);
  const float4 q = xf_affine_relu(__ldg(r + i), ra, rb, ch, 0);
  v.x = fmaxf(v.x + q.x, 0.f); v.y = fmaxf(v.y + q.y, 0.f); v.z = fmaxf(v.z + q.z, 0.f); v.w = fmaxf(v.w + q.w, 0.f);
  out[i] = v;
}

// g = dout * [out > 0] (in place) ; optionally dres (+)= g
__global__ void relu_bwd_add_kernel(float4* __restrict__ dout, const float4* __restrict__ out, float4* __restrict__ dres,
                                    int accumulate, long long n4)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 g = dout[i];
  const float4 o = __ldg(out + i);
  g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
  dout[i] = g;
  if (dres) {
    if (accumulate) { const float4 d = dres[i]; g.x += d.x; g.y += d.y; g.z += d.z; g.w += d.w; }
    dres[i] = g;
  }
}

// ---------------------------------------------------------------- reflect-pad gather
struct GatherArgs {
  const float* src; int s_ct, s_coff, s_pad;      // source tensor: (N, hs + 2 s_pad, ws + 2 s_pad, s_ct), interior at +s_pad
  const float* a; const float* b;                 // mode 2: relu(a x + b), indexed by physical source channel
  float* dst; int d_ct, d_coff;                   // padded destination (N, hu + 2, wu + 2, d_ct)
  int N, hs, ws, C4, shift, mode;                 // hu = hs << shift ; mode 0 identity, 1 ELU, 2 affine + ReLU
};

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void gather_pad_fwd_kernel(GatherArgs g)
{
  const int hu = g.hs << g.shift, wu = g.ws << g.shift;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)g.N * (hu + 2) * (wu + 2) * g.C4) return;
  const int c = (int)(i % g.C4); long long p = i / g.C4;
  const int x = (int)(p % (wu + 2)), y = (int)((p / (wu + 2)) % (hu + 2)), n = (int)(p / ((long long)(wu + 2) * (hu + 2)));
  const int sy = reflect_idx(y - 1, hu) >> g.shift, sx = reflect_idx(x - 1, wu) >> g.shift;
  const int SH = g.hs + 2 * g.s_pad, SW = g.ws + 2 * g.s_pad;
  const int sc = g.s_coff + c * 4;
  float4 v = __ldg(reinterpret_cast<const float4*>(g.src + (((size_t)n * SH + sy + g.s_pad) * SW + sx + g.s_pad) * g.s_ct + sc));
  if (g.mode == 1) {
    v.x = v.x > 0.f ? v.x : expm1f(v.x); v.y = v.y > 0.f ? v.y : expm1f(v.y);
    v.z = v.z > 0.f ? v.z : expm1f(v.z); v.w = v.w > 0.f ? v.w : expm1f(v.w);
  } else if (g.mode == 2) {
    v = xf_affine_relu(v, g.a, g.b, sc, 1);
  }
  *reinterpret_cast<float4*>(g.dst + (size_t)p * g.d_ct + g.d_coff + c * 4) = v;
}

struct GatherBwdArgs {
  const float* dP; int p_ct, p_coff;              // gradient of the padded buffer (N, hu + 2, wu + 2, p_ct)
  const float* src; int s_ct, s_coff, s_pad;      // forward source (mode 1 needs it for ELU')
  float* dsrc; int ds_ct, ds_coff, ds_pad;        // gradient of the source, own geometry
  int N, hs, ws, C4, shift, mode, accumulate;
};

// transposed gather, itself in gather form: source pixel s collects from padded coordinates p with
// reflect(p - 1) >> shift == s, i.e. p = u + 1 for u in [s << shift, (s + 1) << shift), plus the mirrored border
// p = 0 (reflects to u = 1) and p = nu + 1 (reflects to u = nu - 2).
__device__ __forceinline__ int preimages(int s, int nu, int shift, int out[4])
{
  int n = 0;
  const int u0 = s << shift, u1 = ((s + 1) << shift) - 1;
  for (int u = u0; u <= u1; ++u) out[n++] = u + 1;
  if (1 >= u0 && 1 <= u1) out[n++] = 0;
  if (nu - 2 >= u0 && nu - 2 <= u1) out[n++] = nu + 1;
  return n;
}

__global__ void gather_pad_bwd_kernel(GatherBwdArgs g)
{
  const int hu = g.hs << g.shift, wu = g.ws << g.shift;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)g.N * g.hs * g.ws * g.C4) return;
  const int c = (int)(i % g.C4); long long p = i / g.C4;
  const int sx = (int)(p % g.ws), sy = (int)((p / g.ws) % g.hs), n = (int)(p / ((long long)g.ws * g.hs));
  int py[4], px[4];
  const int ny = preimages(sy, hu, g.shift, py), nx = preimages(sx, wu, g.shift, px);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g.dP + (((size_t)n * (hu + 2) + py[a]) * (wu + 2) + px[b]) * g.p_ct + g.p_coff + c * 4));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  if (g.mode == 1) {
    const int SH = g.hs + 2 * g.s_pad, SW = g.ws + 2 * g.s_pad;
    const float4 r = __ldg(reinterpret_cast<const float4*>(g.src + (((size_t)n * SH + sy + g.s_pad) * SW + sx + g.s_pad) * g.s_ct + g.s_coff + c * 4));
    acc.x *= r.x > 0.f ? 1.f : expf(r.x); acc.y *= r.y > 0.f ? 1.f : expf(r.y);
    acc.z *= r.z > 0.f ? 1.f : expf(r.z); acc.w *= r.w > 0.f ? 1.f : expf(r.w);
  }
  const int DH = g.hs + 2 * g.ds_pad, DW = g.ws + 2 * g.ds_pad;
  float4* d = reinterpret_cast<float4*>(g.dsrc + (((size_t)n * DH + sy + g.ds_pad) * DW + sx + g.ds_pad) * g.ds_ct + g.ds_coff + c * 4);
  if (g.accumulate) { const float4 o = *d; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
  *d = acc;
}

// ---------------------------------------------------------------- per-channel sum (conv bias gradient)
__global__ void channel_sum_kernel(const float* __restrict__ x, int ct, int coff, int C, long long npix, float* __restrict__ out)
{
  __shared__ float sh[kT];
  const int rows = kT / C;                        // C <= 256 and a power of two (or 1)
  const int c = threadIdx.x % C, r = threadIdx.x / C;
  float acc = 0.f;
  if (r < rows)
    for (long long p = (long long)blockIdx.x * rows + r; p < npix; p += (long long)gridDim.x * rows) acc += __ldg(x + p * ct + coff + c);
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (r == 0) {
    for (int k = 1; k < rows; ++k) acc += sh[k * C + c];
    atomicAdd(out + c, acc);
  }
}

}  // namespace

#define CVD_M2_CHECK4(C, what) CVD_CHECK_ARG(((C) & 3) == 0 && (C) > 0, what ": channel count must be a positive multiple of 4")

extern "C" int cvd_bicubic_image_fwd(const float* img_nchw, int N, int H, int W, float* out_nhwc4, int oh, int ow,
                                     float mean, float inv_std, void* stream)
{
  CVD_CHECK_ARG(img_nchw && out_nhwc4 && N > 0 && H > 0 && W > 0 && oh > 0 && ow > 0, "cvd_bicubic_image_fwd: bad arguments");
  bicubic_image_kernel<<<blocks_for((long long)N * oh * ow), kT, 0, (cudaStream_t)stream>>>(img_nchw, N, H, W, out_nhwc4, oh, ow, mean, inv_std);
  CVD_LAUNCH_OK("bicubic_image_kernel");
  return 0;
}

extern "C" int cvd_disp_to_depth_fwd(const float* disp, int N, int fh, int fw, float* depth, int H, int W, void* stream)
{
  CVD_CHECK_ARG(disp && depth && N > 0 && fh > 0 && fw > 0 && H > 0 && W > 0, "cvd_disp_to_depth_fwd: bad arguments");
  disp_resize_fwd_kernel<<<blocks_for((long long)N * H * W), kT, 0, (cudaStream_t)stream>>>(disp, N, fh, fw, depth, H, W);
  CVD_LAUNCH_OK("disp_resize_fwd_kernel");
  return 0;
}

extern "C" int cvd_disp_to_depth_bwd(const float* ddepth, const float* depth, int N, int fh, int fw, int H, int W,
                                     float* ddisp, void* stream)
{
  CVD_CHECK_ARG(ddepth && depth && ddisp && N > 0, "cvd_disp_to_depth_bwd: bad arguments");
  cudaError_t e = cudaMemsetAsync(ddisp, 0, (size_t)N * fh * fw * sizeof(float), (cudaStream_t)stream);
  if (e != cudaSuccess) return cvd_fail("cvd_disp_to_depth_bwd: memset: %s", cudaGetErrorString(e));
  disp_resize_bwd_kernel<<<blocks_for((long long)N * H * W), kT, 0, (cudaStream_t)stream>>>(ddepth, depth, N, fh, fw, H, W, ddisp);
  CVD_LAUNCH_OK("disp_resize_bwd_kernel");
  return 0;
}

extern "C" int cvd_sigmoid_fwd(const float* raw_padded, int c_total, int N, int fh, int fw, float* disp, void* stream)
{
  CVD_CHECK_ARG(raw_padded && disp && c_total > 0 && N > 0, "cvd_sigmoid_fwd: bad arguments");
  sigmoid_fwd_kernel<<<blocks_for((long long)N * fh * fw), kT, 0, (cudaStream_t)stream>>>(raw_padded, c_total, N, fh, fw, disp);
  CVD_LAUNCH_OK("sigmoid_fwd_kernel");
  return 0;
}

extern "C" int cvd_sigmoid_bwd(const float* ddisp, const float* disp, int N, int fh, int fw, float* draw_padded, int c_total, void* stream)
{
  CVD_CHECK_ARG(ddisp && disp && draw_padded && c_total > 0 && N > 0, "cvd_sigmoid_bwd: bad arguments");
  sigmoid_bwd_kernel<<<blocks_for((long long)N * fh * fw), kT, 0, (cudaStream_t)stream>>>(ddisp, disp, N, fh, fw, draw_padded, c_total);
  CVD_LAUNCH_OK("sigmoid_bwd_kernel");
  return 0;
}

extern "C" int cvd_subsample2(const float* src, int N, int H, int W, int C, float* dst, void* stream)
{
  CVD_CHECK_ARG(src && dst && N > 0 && H > 0 && W > 0, "cvd_subsample2: bad arguments");
  CVD_M2_CHECK4(C, "cvd_subsample2");
  const int oh = (H + 1) / 2, ow = (W + 1) / 2;
  subsample2_kernel<<<blocks_for((long long)N * oh * ow * (C / 4)), kT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(src), N, H, W, C / 4, reinterpret_cast<float4*>(dst), oh, ow);
  CVD_LAUNCH_OK("subsample2_kernel");
  return 0;
}

extern "C" int cvd_stuff2(const float* src, int N, int H, int W, int C, float* dst, int accumulate, void* stream)
{
  CVD_CHECK_ARG(src && dst && N > 0 && H > 0 && W > 0, "cvd_stuff2: bad arguments");
  CVD_M2_CHECK4(C, "cvd_stuff2");
  const int oh = (H + 1) / 2, ow = (W + 1) / 2;
  stuff2_kernel<<<blocks_for((long long)N * oh * ow * (C / 4)), kT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(src), N, oh, ow, C / 4, reinterpret_cast<float4*>(dst), H, W, accumulate);
  CVD_LAUNCH_OK("stuff2_kernel");
  return 0;
}

extern "C" int cvd_bnbwd_stuff(const float* x, const float* dy, const float* a, const float* b, const float* bw, int relu,
                               int N, int h, int w, int C, float* dst, int H, int W, int stride, void* stream)
{
  CVD_CHECK_ARG(x && dy && a && b && bw && dst && N > 0 && (stride == 1 || stride == 2), "cvd_bnbwd_stuff: bad arguments");
  CVD_CHECK_ARG((h - 1) * stride < H && (w - 1) * stride < W, "cvd_bnbwd_stuff: destination too small");
  CVD_M2_CHECK4(C, "cvd_bnbwd_stuff");
  bnbwd_stuff2_kernel<<<blocks_for((long long)N * h * w * (C / 4)), kT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(dy), a, b, reinterpret_cast<const float4*>(bw), relu,
      N, h, w, C / 4, reinterpret_cast<float4*>(dst), H, W, stride);
  CVD_LAUNCH_OK("bnbwd_stuff2_kernel");
  return 0;
}

extern "C" int cvd_maxpool3s2_fwd(const float* x, const float* a, const float* b, int relu, int N, int H, int W, int C,
                                  float* out, unsigned char* argmax, void* stream)
{
  CVD_CHECK_ARG(x && out && argmax && N > 0 && H > 0 && W > 0, "cvd_maxpool3s2_fwd: bad arguments");
  CVD_M2_CHECK4(C, "cvd_maxpool3s2_fwd");
  const int oh = (H + 1) / 2, ow = (W + 1) / 2;
  maxpool_fwd_kernel<<<blocks_for((long long)N * oh * ow * (C / 4)), kT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), a, b, relu, N, H, W, C / 4, reinterpret_cast<float4*>(out),
      reinterpret_cast<uchar4*>(argmax), oh, ow);
  CVD_LAUNCH_OK("maxpool_fwd_kernel");
  return 0;
}

extern "C" int cvd_maxpool3s2_bwd(const float* dout, const unsigned char* argmax, int N, int H, int W, int C, float* dx,
                                  int accumulate, void* stream)
{
  CVD_CHECK_ARG(dout && argmax && dx && N > 0 && H > 0 && W > 0, "cvd_maxpool3s2_bwd: bad arguments");
  CVD_M2_CHECK4(C, "cvd_maxpool3s2_bwd");
  const int oh = (H + 1) / 2, ow = (W + 1) / 2;
  maxpool_bwd_kernel<<<blocks_for((long long)N * H * W * (C / 4)), kT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(dout), reinterpret_cast<const uchar4*>(argmax), N, H, W, C / 4, oh, ow,
      reinterpret_cast<float4*>(dx), accumulate);
  CVD_LAUNCH_OK("maxpool_bwd_kernel");
  return 0;
}

extern "C" int cvd_bn_add_relu(const float* y, const float* a, const float* b, const float* res, const float* ra, const float* rb,
                               long long npix, int C, float* out, void* stream)
{
  CVD_CHECK_ARG(y && a && b && res && out && npix > 0 && (!ra == !rb), "cvd_bn_add_relu: bad arguments");
  CVD_M2_CHECK4(C, "cvd_bn_add_relu");
  bn_add_relu_kernel<<<blocks_for(npix * (C / 4)), kT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(y), a, b, reinterpret_cast<const float4*>(res), ra, rb, npix * (C / 4), C / 4,
      reinterpret_cast<float4*>(out));
  CVD_LAUNCH_OK("bn_add_relu_kernel");
  return 0;
}

extern "C" int cvd_relu_bwd_add(float* dout, const float* out, float* dres, int accumulate, long long n, void* stream)
{
  CVD_CHECK_ARG(dout && out && n > 0 && (n & 3) == 0, "cvd_relu_bwd_add: bad arguments");
  relu_bwd_add_kernel<<<blocks_for(n / 4), kT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float4*>(dout), reinterpret_cast<const float4*>(out), reinterpret_cast<float4*>(dres), accumulate, n / 4);
  CVD_LAUNCH_OK("relu_bwd_add_kernel");
  return 0;
}

extern "C" int cvd_gather_pad_fwd(const float* src, int s_ctotal, int s_coff, int s_pad, const float* a, const float* b,
                                  float* dst, int d_ctotal, int d_coff, int N, int hs, int ws, int C, int upsample, int mode,
                                  void* stream)
{
  CVD_CHECK_ARG(src && dst && N > 0 && hs > 0 && ws > 0 && (upsample == 0 || upsample == 1), "cvd_gather_pad_fwd: bad arguments");
  CVD_CHECK_ARG(mode == CVD_GATHER_IDENTITY || mode == CVD_GATHER_ELU || (mode == CVD_GATHER_AFFINE_RELU && a && b), "cvd_gather_pad_fwd: bad mode");
  CVD_CHECK_ARG((hs << upsample) >= 2 && (ws << upsample) >= 2, "cvd_gather_pad_fwd: reflection needs at least 2 pixels");
  CVD_CHECK_ARG(((s_ctotal | s_coff | d_ctotal | d_coff) & 3) == 0, "cvd_gather_pad_fwd: channel strides / offsets must be multiples of 4");
  CVD_M2_CHECK4(C, "cvd_gather_pad_fwd");
  GatherArgs g{src, s_ctotal, s_coff, s_pad, a, b, dst, d_ctotal, d_coff, N, hs, ws, C / 4, upsample, mode};
  const long long n = (long long)N * ((hs << upsample) + 2) * ((ws << upsample) + 2) * (C / 4);
  gather_pad_fwd_kernel<<<blocks_for(n), kT, 0, (cudaStream_t)stream>>>(g);
  CVD_LAUNCH_OK("gather_pad_fwd_kernel");
  return 0;
}

extern "C" int cvd_gather_pad_bwd(const float* dpad, int p_ctotal, int p_coff, const float* src, int s_ctotal, int s_coff, int s_pad,
                                  float* dsrc, int ds_ctotal, int ds_coff, int ds_pad, int N, int hs, int ws, int C, int upsample,
                                  int mode, int accumulate, void* stream)
{
  CVD_CHECK_ARG(dpad && dsrc && N > 0 && hs > 0 && ws > 0 && (upsample == 0 || upsample == 1), "cvd_gather_pad_bwd: bad arguments");
  CVD_CHECK_ARG(mode != CVD_GATHER_ELU || src, "cvd_gather_pad_bwd: ELU backward needs the forward source");
  CVD_CHECK_ARG((hs << upsample) >= 2 && (ws << upsample) >= 2, "cvd_gather_pad_bwd: reflection needs at least 2 pixels");
  CVD_CHECK_ARG(((p_ctotal | p_coff | s_ctotal | s_coff | ds_ctotal | ds_coff) & 3) == 0, "cvd_gather_pad_bwd: channel strides / offsets must be multiples of 4");
  CVD_M2_CHECK4(C, "cvd_gather_pad_bwd");
  GatherBwdArgs g{dpad, p_ctotal, p_coff, src, s_ctotal, s_coff, s_pad, dsrc, ds_ctotal, ds_coff, ds_pad, N, hs, ws, C / 4,
                  upsample, mode, accumulate};
  gather_pad_bwd_kernel<<<blocks_for((long long)N * hs * ws * (C / 4)), kT, 0, (cudaStream_t)stream>>>(g);
  CVD_LAUNCH_OK("gather_pad_bwd_kernel");
  return 0;
}

extern "C" int cvd_channel_sum(const float* x, int c_total, int c_off, int C, long long npix, float* out, void* stream)
{
  CVD_CHECK_ARG(x && out && npix > 0 && C >= 1 && C <= 256 && (C & (C - 1)) == 0, "cvd_channel_sum: C must be a power of two <= 256");
  long long b = (npix + (kT / C) - 1) / (kT / C);
  const long long cap = (long long)cvd_num_sms() * 8;
  if (b > cap) b = cap;
  channel_sum_kernel<<<(unsigned)b, kT, 0, (cudaStream_t)stream>>>(x, c_total, c_off, C, npix, out);
  CVD_LAUNCH_OK("channel_sum_kernel");
  return 0;
}
