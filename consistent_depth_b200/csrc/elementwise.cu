// Memory-bound glue of the hourglass network around the tcgen05 convs (all NHWC fp32):
// BatchNorm batch statistics / finalisation / backward reductions, AvgPool2d(2), the
// bilinear x2 upsample + residual add of the hourglass levels, layout conversion.
//
// Reference ops replaced (monodepth/mannequin_challenge/models/hourglass.py):
//   nn.BatchNorm2d train mode (:28,40,43,165)   -> bn_stats (+ finalize), bn_bwd_reduce (+ finalize)
//   nn.AvgPool2d(2) (:70,95,113,138)            -> pool_fwd / pool_bwd
//   nn.UpsamplingBilinear2d(2) + branch sum (:74,81,101,106,119,131,144,156) -> merge_up_fwd / up2x_bwd
// Every kernel streams its tensors exactly once with 128-bit accesses; per-channel
// reductions go block-local in fp32, cross-block in fp64 atomics (few thousand per layer).
#include "cvd_common.cuh"
#include <cstdlib>

namespace {

__device__ __forceinline__ int vphys(int c, int c0, int n0, int gap) { return c0 + c + (c >= n0 ? gap : 0); }

// ---------------------------------------------------------------- BN forward statistics
// x: raw conv output, physical channels [c0, c0+C) of a c_total-wide NHWC buffer.
// Thread layout: C/4 channel-quads across threads, pixels strided; block partials -> f64 atomics.
// The last block to finish (ticket) finalises: mean / biased var -> a = gamma*rstd, b = beta - mean*a,
// running stats (momentum, unbiased var), rstd saved for backward; scratch re-zeroed for the next step.
__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ x, int c_total, int c0, int C, long long npix,
                double* __restrict__ scratch /* [C][2] + ticket */, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, float momentum,
                float* __restrict__ running_mean, float* __restrict__ running_var,
                float* __restrict__ a, float* __restrict__ b, float* __restrict__ rstd_out,
                float* __restrict__ mean_out)
{
  extern __shared__ double sh[];                   // [2][C]
  const int cq = C >> 2;                           // channel quads
  const int lanes = 256 / cq;                      // pixel lanes per block (cq <= 64 guaranteed by host)
  const int q = threadIdx.x % cq, pl = threadIdx.x / cq;
  for (int i = threadIdx.x; i < 2 * C; i += 256) sh[i] = 0.0;
  __syncthreads();
  float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
  if (pl < lanes) {
    const long long stride = (long long)gridDim.x * lanes;
    int cnt = 0;
    double ds[4] = {0, 0, 0, 0}, dss[4] = {0, 0, 0, 0};
    // 4 independent 128-bit loads in flight per thread (the pass is pure streaming: latency must be hidden)
    for (long long p0 = (long long)blockIdx.x * lanes + pl; p0 < npix; p0 += 8 * stride) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long p = p0 + u * stride;
        v[u] = p < npix ? __ldg(reinterpret_cast<const float4*>(x + p * c_total + c0 + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
        ss[0] = fmaf(v[u].x, v[u].x, ss[0]); ss[1] = fmaf(v[u].y, v[u].y, ss[1]);
        ss[2] = fmaf(v[u].z, v[u].z, ss[2]); ss[3] = fmaf(v[u].w, v[u].w, ss[3]);
      }
      if (++cnt == 8) {                            // bound fp32 partial sums to 64 terms
#pragma unroll
        for (int i = 0; i < 4; ++i) { ds[i] += s[i]; dss[i] += ss[i]; s[i] = 0.f; ss[i] = 0.f; }
        cnt = 0;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(&sh[4 * q + i], ds[i] + (double)s[i]);
      atomicAdd(&sh[C + 4 * q + i], dss[i] + (double)ss[i]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    atomicAdd(&scratch[2 * c], sh[c]);
    atomicAdd(&scratch[2 * c + 1], sh[C + c]);
  }
  __threadfence();
  __shared__ int last;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + 2 * C);
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    if (last) *ticket = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += 256) {
    const double sum = __ldcg(&scratch[2 * c]), sq = __ldcg(&scratch[2 * c + 1]);
    scratch[2 * c] = 0.0; scratch[2 * c + 1] = 0.0;
    const double mean = sum / (double)npix;
    double var = sq / (double)npix - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    const float av = g * rs;
    a[c0 + c] = av;
    b[c0 + c] = be - (float)mean * av;
    rstd_out[c0 + c] = rs;
    mean_out[c0 + c] = (float)mean;
    if (running_mean) {
      const double unb = npix > 1 ? var * (double)npix / (double)(npix - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  }
}

// ---------------------------------------------------------------- BN(+ReLU) backward reductions
// y = a x + b ; g = dy * [y > 0 or !relu] ; sums over pixels of g and g*y.  Finalise:
//   mg = mean g ; mgy = mean g y ; mgxh = (mgy - beta mg)/gamma
//   bw[c] = (c0, c1, c2, 0) with dx = c0 g - c1 - c2 y ; c0 = rstd gamma ; c2 = rstd mgxh ; c1 = c0 mg - c2 beta
//   dgamma = npix mgxh ; dbeta = npix mg ; dbias (conv bias feeding this BN) = sum dx = npix (c0 mg - c1 - c2 (a mean_x + b))
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ x, int x_ct, int x_c0,
                     const float* __restrict__ dy, int dy_ct, int dy_c0, int dy_n0, int dy_gap, int dy_lc0,
                     const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ rstd,
                     const float* __restrict__ mean_x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     int relu, long long npix, int C, double* __restrict__ scratch,
                     float4* __restrict__ bw, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ dbias)
{
  extern __shared__ double sh[];                   // [2][C]
  if (blockIdx.y) {
    // chunked launch: blockIdx.y = 256-channel chunk of a wider BatchNorm, one statistics block each
    const int ch = (int)blockIdx.y * 256;
    x_c0 += ch; dy_lc0 += ch; scratch += (size_t)blockIdx.y * (2 * 256 + 1);
    if (gamma) gamma += ch;
    if (beta) beta += ch;
    if (dgamma) dgamma += ch;
    if (dbeta) dbeta += ch;
    if (dbias) dbias += ch;
  }
  const int cq = C >> 2;
  const int lanes = 256 / cq;
  const int q = threadIdx.x % cq, pl = threadIdx.x / cq;
  for (int i = threadIdx.x; i < 2 * C; i += 256) sh[i] = 0.0;
  __syncthreads();
  if (pl < lanes) {
    const float4 av = *reinterpret_cast<const float4*>(a + x_c0 + 4 * q);
    const float4 bv = *reinterpret_cast<const float4*>(b + x_c0 + 4 * q);
    const int dc = vphys(dy_lc0 + 4 * q, dy_c0, dy_n0, dy_gap);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, sy[4] = {0.f, 0.f, 0.f, 0.f};
    double ds[4] = {0, 0, 0, 0}, dsy[4] = {0, 0, 0, 0};
    int cnt = 0;
    const long long stride = (long long)gridDim.x * lanes;
    for (long long p0 = (long long)blockIdx.x * lanes + pl; p0 < npix; p0 += 4 * stride) {
      float4 xq[4], dq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long p = p0 + u * stride;
        const bool ok = p < npix;
        xq[u] = ok ? __ldg(reinterpret_cast<const float4*>(x + p * x_ct + x_c0 + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
        dq[u] = ok ? __ldg(reinterpret_cast<const float4*>(dy + p * dy_ct + dc)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 xv = xq[u], dv = dq[u];
        const float y0 = fmaf(av.x, xv.x, bv.x), y1 = fmaf(av.y, xv.y, bv.y), y2 = fmaf(av.z, xv.z, bv.z), y3 = fmaf(av.w, xv.w, bv.w);
        const float g0 = (!relu || y0 > 0.f) ? dv.x : 0.f, g1 = (!relu || y1 > 0.f) ? dv.y : 0.f;
        const float g2 = (!relu || y2 > 0.f) ? dv.z : 0.f, g3 = (!relu || y3 > 0.f) ? dv.w : 0.f;
        s[0] += g0; s[1] += g1; s[2] += g2; s[3] += g3;
        sy[0] = fmaf(g0, y0, sy[0]); sy[1] = fmaf(g1, y1, sy[1]); sy[2] = fmaf(g2, y2, sy[2]); sy[3] = fmaf(g3, y3, sy[3]);
      }
      if (++cnt == 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { ds[i] += s[i]; dsy[i] += sy[i]; s[i] = 0.f; sy[i] = 0.f; }
        cnt = 0;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(&sh[4 * q + i], ds[i] + (double)s[i]);
      atomicAdd(&sh[C + 4 * q + i], dsy[i] + (double)sy[i]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    atomicAdd(&scratch[2 * c], sh[c]);
    atomicAdd(&scratch[2 * c + 1], sh[C + c]);
  }
  __threadfence();
  __shared__ int last;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + 2 * C);
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    if (last) *ticket = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += 256) {
    const double sg = __ldcg(&scratch[2 * c]), sgy = __ldcg(&scratch[2 * c + 1]);
    scratch[2 * c] = 0.0; scratch[2 * c + 1] = 0.0;
    const double mg = sg / (double)npix, mgy = sgy / (double)npix;
    const double ga = gamma ? (double)gamma[c] : 1.0, be = beta ? (double)beta[c] : 0.0;
    const double mgxh = (mgy - be * mg) / ga;
    const double rs = (double)rstd[x_c0 + c];
    const double c0 = rs * ga, c2 = rs * mgxh, c1 = c0 * mg - c2 * be;
    bw[x_c0 + c] = make_float4((float)c0, (float)c1, (float)c2, 0.f);
    if (dgamma) dgamma[c] = (float)((double)npix * mgxh);
    if (dbeta) dbeta[c] = (float)((double)npix * mg);
    if (dbias) {
      const double my = (double)a[x_c0 + c] * (double)mean_x[x_c0 + c] + (double)b[x_c0 + c];
      dbias[c] = (float)((double)npix * (c0 * mg - c1 - c2 * my));
    }
  }
}

// ---------------------------------------------------------------- pool / upsample+merge
__device__ __forceinline__ float4 xf4(const float* __restrict__ x, long long pix, int ct, int pc,
                                      const float* __restrict__ a, const float* __restrict__ b, int relu)
{
  float4 v = __ldg(reinterpret_cast<const float4*>(x + pix * ct + pc));
  if (a) {
    const float4 av = __ldg(reinterpret_cast<const float4*>(a + pc)), bv = __ldg(reinterpret_cast<const float4*>(b + pc));
    v.x = fmaf(av.x, v.x, bv.x); v.y = fmaf(av.y, v.y, bv.y); v.z = fmaf(av.z, v.z, bv.z); v.w = fmaf(av.w, v.w, bv.w);
  }
  if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  return v;
}

// p[n, y, x, c] = mean of the 2x2 block of relu(a x + b); x is a channel view, p is plain (N,H/2,W/2,C)
__global__ void __launch_bounds__(256)
pool_fwd_kernel(const float* __restrict__ x, int ct, int c0, int n0, int gap, const float* __restrict__ a,
                const float* __restrict__ b, int relu, float* __restrict__ p, int N, int H, int W, int C)
{
  const int cq = C >> 2, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * Ho * Wo * cq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cq); long long r = i / cq;
    const int xo = (int)(r % Wo); r /= Wo; const int yo = (int)(r % Ho); const int n = (int)(r / Ho);
    const int pc = vphys(4 * q, c0, n0, gap);
    const long long base = ((long long)n * H + 2 * yo) * W + 2 * xo;
    const float4 v00 = xf4(x, base, ct, pc, a, b, relu), v01 = xf4(x, base + 1, ct, pc, a, b, relu);
    const float4 v10 = xf4(x, base + W, ct, pc, a, b, relu), v11 = xf4(x, base + W + 1, ct, pc, a, b, relu);
    float4 o;
    o.x = ((v00.x + v01.x) + (v10.x + v11.x)) * 0.25f; o.y = ((v00.y + v01.y) + (v10.y + v11.y)) * 0.25f;
    o.z = ((v00.z + v01.z) + (v10.z + v11.z)) * 0.25f; o.w = ((v00.w + v01.w) + (v10.w + v11.w)) * 0.25f;
    reinterpret_cast<float4*>(p)[i] = o;
  }
}

// dx[n, y, x, view(c)] (+)= 0.25 dp[n, y/2, x/2, c]   (dp plain C channels, dx through a channel view)
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const float* __restrict__ dp, float* __restrict__ dx, int ct, int c0, int n0, int gap,
                int accumulate, int N, int H, int W, int C)
{
  const int cq = C >> 2, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * H * W * cq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cq); long long r = i / cq;
    const int xx = (int)(r % W); r /= W; const int yy = (int)(r % H); const int n = (int)(r / H);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((yy >> 1) < Ho && (xx >> 1) < Wo) {
      g = __ldg(reinterpret_cast<const float4*>(dp) + (((long long)n * Ho + (yy >> 1)) * Wo + (xx >> 1)) * cq + q);
      g.x *= 0.25f; g.y *= 0.25f; g.z *= 0.25f; g.w *= 0.25f;
    }
    float4* o = reinterpret_cast<float4*>(dx + (((long long)n * H + yy) * W + xx) * ct + vphys(4 * q, c0, n0, gap));
    if (accumulate) { const float4 old = *o; g.x += old.x; g.y += old.y; g.z += old.z; g.w += old.w; }
    *o = g;
  }
}

// z = relu(a1 x1 + b1) + up2x(relu(a2 x2 + b2)); bilinear, align_corners=True (nn.UpsamplingBilinear2d)
__global__ void __launch_bounds__(256)
merge_up_fwd_kernel(const float* __restrict__ x1, int ct1, int c01, int n01, int gap1,
                    const float* __restrict__ a1, const float* __restrict__ b1,
                    const float* __restrict__ x2, int ct2, int c02, int n02, int gap2,
                    const float* __restrict__ a2, const float* __restrict__ b2,
                    float* __restrict__ z, int N, int H, int W, int C)
{
  const int cq = C >> 2, Hs = H >> 1, Ws = W >> 1;
  const float sy = Hs > 1 ? (float)(Hs - 1) / (float)(H - 1) : 0.f;
  const float sx = Ws > 1 ? (float)(Ws - 1) / (float)(W - 1) : 0.f;
  const long long total = (long long)N * H * W * cq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cq); long long r = i / cq;
    const int xx = (int)(r % W); r /= W; const int yy = (int)(r % H); const int n = (int)(r / H);
    const int pc1 = vphys(4 * q, c01, n01, gap1), pc2 = vphys(4 * q, c02, n02, gap2);
    const float4 v1 = xf4(x1, ((long long)n * H + yy) * W + xx, ct1, pc1, a1, b1, 1);
    const float fy = sy * (float)yy, fx = sx * (float)xx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = min(y0 + 1, Hs - 1), x1i = min(x0 + 1, Ws - 1);
    const float wy = fy - (float)y0, wx = fx - (float)x0;
    const long long sb = (long long)n * Hs * Ws;
    const float4 u00 = xf4(x2, sb + (long long)y0 * Ws + x0, ct2, pc2, a2, b2, 1);
    const float4 u01 = xf4(x2, sb + (long long)y0 * Ws + x1i, ct2, pc2, a2, b2, 1);
    const float4 u10 = xf4(x2, sb + (long long)y1 * Ws + x0, ct2, pc2, a2, b2, 1);
    const float4 u11 = xf4(x2, sb + (long long)y1 * Ws + x1i, ct2, pc2, a2, b2, 1);
    const float w00 = (1.f - wy) * (1.f - wx), w01 = (1.f - wy) * wx, w10 = wy * (1.f - wx), w11 = wy * wx;
    float4 o;
    o.x = v1.x + (w00 * u00.x + w01 * u01.x + w10 * u10.x + w11 * u11.x);
    o.y = v1.y + (w00 * u00.y + w01 * u01.y + w10 * u10.y + w11 * u11.y);
    o.z = v1.z + (w00 * u00.z + w01 * u01.z + w10 * u10.z + w11 * u11.z);
    o.w = v1.w + (w00 * u00.w + w01 * u01.w + w10 * u10.w + w11 * u11.w);
    reinterpret_cast<float4*>(z)[i] = o;
  }
}

// Backward of z = y1 + up2x(y2) for plain dz (N,H,W,C):
//   dy2[n, ys, xs, view2(c)] = sum over the (<= 4x4) fine pixels whose bilinear footprint touches (ys, xs)
//                              (gather form of the transposed upsample: deterministic, no atomics)
//   dy1[n, y, x, view1(c)]   = dz[n, y, x, c]  (the thread also forwards its 2x2 block; dy1 may be NULL)
__global__ void __launch_bounds__(256)
up2x_bwd_kernel(const float* __restrict__ dz, float* __restrict__ dy2, int ct2, int c02, int n02, int gap2,
                float* __restrict__ dy1, int ct1, int c01, int n01, int gap1, int acc1, int N, int H, int W, int C)
{
  const int cq = C >> 2, Hs = H >> 1, Ws = W >> 1;
  const float sy = Hs > 1 ? (float)(Hs - 1) / (float)(H - 1) : 0.f;
  const float sx = Ws > 1 ? (float)(Ws - 1) / (float)(W - 1) : 0.f;
  const long long total = (long long)N * Hs * Ws * cq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cq); long long r = i / cq;
    const int xs = (int)(r % Ws); r /= Ws; const int ys = (int)(r % Hs); const int n = (int)(r / Hs);
    const int ylo = max(0, 2 * ys - 3), yhi = min(H - 1, 2 * ys + 3);
    const int xlo = max(0, 2 * xs - 3), xhi = min(W - 1, 2 * xs + 3);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int yy = ylo; yy <= yhi; ++yy) {
      const float fy = sy * (float)yy; const int y0 = (int)fy; const float wy = fy - (float)y0;
      const int y1 = min(y0 + 1, Hs - 1);
      float wyv = 0.f;
      if (y0 == ys) wyv += 1.f - wy;
      if (y1 == ys) wyv += wy;
      if (wyv == 0.f) continue;
      for (int xx = xlo; xx <= xhi; ++xx) {
        const float fx = sx * (float)xx; const int x0 = (int)fx; const float wx = fx - (float)x0;
        const int x1 = min(x0 + 1, Ws - 1);
        float wxv = 0.f;
        if (x0 == xs) wxv += 1.f - wx;
        if (x1 == xs) wxv += wx;
        if (wxv == 0.f) continue;
        const float wgt = wyv * wxv;
        const float4 g = __ldg(reinterpret_cast<const float4*>(dz) + (((long long)n * H + yy) * W + xx) * cq + q);
        acc.x = fmaf(wgt, g.x, acc.x); acc.y = fmaf(wgt, g.y, acc.y); acc.z = fmaf(wgt, g.z, acc.z); acc.w = fmaf(wgt, g.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(dy2 + (((long long)n * Hs + ys) * Ws + xs) * ct2 + vphys(4 * q, c02, n02, gap2)) = acc;
    if (dy1) {
      const int pc1 = vphys(4 * q, c01, n01, gap1);
#pragma unroll
      for (int dyy = 0; dyy < 2; ++dyy)
#pragma unroll
        for (int dxx = 0; dxx < 2; ++dxx) {
          const long long pix = ((long long)n * H + 2 * ys + dyy) * W + 2 * xs + dxx;
          float4 g = __ldg(reinterpret_cast<const float4*>(dz) + pix * cq + q);
          float4* o = reinterpret_cast<float4*>(dy1 + pix * ct1 + pc1);
          if (acc1) { const float4 old = *o; g.x += old.x; g.y += old.y; g.z += old.z; g.w += old.w; }
          *o = g;
        }
    }
  }
}

// (N,3,H,W) NCHW BGR image -> (N,H,W,4) with a zero 4th channel
__global__ void __launch_bounds__(256)
image_to_nhwc4_kernel(const float* __restrict__ img, float* __restrict__ out, int N, int H, int W)
{
  const long long HW = (long long)H * W, total = (long long)N * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i - n * HW;
    const float* s = img + n * 3 * HW + p;
    reinterpret_cast<float4*>(out)[i] = make_float4(__ldg(s), __ldg(s + HW), __ldg(s + 2 * HW), 0.f);
  }
}

// d log-depth = dL/d depth * depth (mannequin_challenge_model.py:66 exp backward), written as
// channel 0 of an (N,H,W,4) tensor for the pred layer's dgrad / wgrad; also accumulates the
// pred-layer bias gradient sum(d log-depth).
__global__ void __launch_bounds__(256)
dlogdepth_kernel(const float* __restrict__ gdepth, const float* __restrict__ depth, float* __restrict__ out4,
                 long long n, float* __restrict__ dbias)
{
  __shared__ float red[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(gdepth + i) * __ldg(depth + i);
    reinterpret_cast<float4*>(out4)[i] = make_float4(v, 0.f, 0.f, 0.f);
    s += v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0 && dbias) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(dbias, t);
  }
}

inline unsigned ew_grid(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)cvd_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

// one block of [2 x 256 sums | ticket] per 256 channels: chunked conv launches accumulate their chunks concurrently
extern "C" size_t cvd_bn_scratch_bytes(int C) { return (size_t)((C + 255) / 256 > 0 ? (C + 255) / 256 : 1) * (2 * 256 + 1) * sizeof(double); }

extern "C" int cvd_bn_stats(const float* x, int c_total, int c_off, int C, long long npix, void* scratch,
                            const float* gamma, const float* beta, float eps, float momentum,
                            float* running_mean, float* running_var,
                            float* a, float* b, float* rstd, float* mean, void* stream)
{
  CVD_CHECK_ARG(x && scratch && a && b && rstd && mean, "cvd_bn_stats: null pointer");
  if (C > 256) {            // 256-channel chunks (the block-level reduction holds 256 channels); scratch is reused in stream order
    for (int c0 = 0; c0 < C; c0 += 256) {
      const int rc = cvd_bn_stats(x, c_total, c_off + c0, C - c0 < 256 ? C - c0 : 256, npix, scratch,
                                  gamma ? gamma + c0 : nullptr, beta ? beta + c0 : nullptr, eps, momentum,
                                  running_mean ? running_mean + c0 : nullptr, running_var ? running_var + c0 : nullptr,
                                  a, b, rstd, mean, stream);
      if (rc) return rc;
    }
    return 0;
  }
  CVD_CHECK_ARG(C > 0 && C <= 256 && (C & 3) == 0 && (c_off & 3) == 0 && (c_total & 3) == 0 && npix > 0,
                "cvd_bn_stats: bad channels C=%d c_off=%d c_total=%d", C, c_off, c_total);
  const int lanes = 256 / (C >> 2);
  long long blocks = (npix + lanes * 16 - 1) / ((long long)lanes * 16);   // small tensors: spread over the SMs
  const long long cap = (long long)cvd_num_sms() * 2;       // few blocks: the f64 atomics per block are the tail cost
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  bn_stats_kernel<<<(unsigned)blocks, 256, 2 * C * sizeof(double), (cudaStream_t)stream>>>(
      x, c_total, c_off, C, npix, (double*)scratch, gamma, beta, eps, momentum, running_mean, running_var, a, b, rstd, mean);
  CVD_LAUNCH_OK("bn_stats_kernel");
  return 0;
}

extern "C" int cvd_bn_bwd_reduce(const float* x, int x_ctotal, int x_coff,
                                 const float* dy, int dy_ctotal, int dy_coff, int dy_n0, int dy_gap, int dy_lc0,
                                 const float* a, const float* b, const float* rstd, const float* mean,
                                 const float* gamma, const float* beta, int relu,
                                 long long npix, int C, void* scratch,
                                 float* bw, float* dgamma, float* dbeta, float* dbias, void* stream)
{
  CVD_CHECK_ARG(x && dy && a && b && rstd && mean && scratch && bw, "cvd_bn_bwd_reduce: null pointer");
  int nchunks = 1;
  if (C > 256 && C % 256 == 0 && !getenv("CVD_BN_NO_CHUNKS")) { nchunks = C / 256; C = 256; }   // one launch, blockIdx.y = chunk
  if (C > 256) {
    for (int c0 = 0; c0 < C; c0 += 256) {
      const int rc = cvd_bn_bwd_reduce(x, x_ctotal, x_coff + c0, dy, dy_ctotal, dy_coff, dy_n0, dy_gap, dy_lc0 + c0,
                                       a, b, rstd, mean, gamma ? gamma + c0 : nullptr, beta ? beta + c0 : nullptr, relu,
                                       npix, C - c0 < 256 ? C - c0 : 256, scratch, bw,
                                       dgamma ? dgamma + c0 : nullptr, dbeta ? dbeta + c0 : nullptr,
                                       dbias ? dbias + c0 : nullptr, stream);
      if (rc) return rc;
    }
    return 0;
  }
  CVD_CHECK_ARG(C > 0 && C <= 256 && (C & 3) == 0 && (x_coff & 3) == 0 && (x_ctotal & 3) == 0 && (dy_ctotal & 3) == 0,
                "cvd_bn_bwd_reduce: bad channels");
  const int lanes = 256 / (C >> 2);
  long long blocks = (npix + lanes * 16 - 1) / ((long long)lanes * 16);
  static const int bpsm = getenv("CVD_BNBWD_BLOCKS") ? atoi(getenv("CVD_BNBWD_BLOCKS")) : 3;
  const long long cap = ((long long)cvd_num_sms() * (bpsm > 0 ? bpsm : 3) + nchunks - 1) / nchunks;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  bn_bwd_reduce_kernel<<<dim3((unsigned)blocks, (unsigned)nchunks), 256, 2 * C * sizeof(double), (cudaStream_t)stream>>>(
      x, x_ctotal, x_coff, dy, dy_ctotal, dy_coff, dy_n0 > 0 ? dy_n0 : (1 << 30), dy_gap, dy_lc0, a, b, rstd, mean,
      gamma, beta, relu, npix, C, (double*)scratch, reinterpret_cast<float4*>(bw), dgamma, dbeta, dbias);
  CVD_LAUNCH_OK("bn_bwd_reduce_kernel");
  return 0;
}

extern "C" int cvd_pool_fwd(const float* x, int c_total, int c_off, int n0, int gap, const float* a, const float* b,
                            int relu, float* p, int N, int H, int W, int C, void* stream)
{
  CVD_CHECK_ARG(x && p, "cvd_pool_fwd: null pointer");
  CVD_CHECK_ARG((C & 3) == 0 && (H & 1) == 0 && (W & 1) == 0, "cvd_pool_fwd: C %% 4, even H, W required");
  pool_fwd_kernel<<<ew_grid((long long)N * (H / 2) * (W / 2) * (C / 4)), 256, 0, (cudaStream_t)stream>>>(
      x, c_total, c_off, n0 > 0 ? n0 : (1 << 30), gap, a, b, relu, p, N, H, W, C);
  CVD_LAUNCH_OK("pool_fwd_kernel");
  return 0;
}

extern "C" int cvd_pool_bwd(const float* dp, float* dx, int c_total, int c_off, int n0, int gap, int accumulate,
                            int N, int H, int W, int C, void* stream)
{
  CVD_CHECK_ARG(dp && dx, "cvd_pool_bwd: null pointer");
  CVD_CHECK_ARG((C & 3) == 0 && (c_total & 3) == 0 && (c_off & 3) == 0 && (gap & 3) == 0, "cvd_pool_bwd: 4-channel alignment required");
  pool_bwd_kernel<<<ew_grid((long long)N * H * W * (C / 4)), 256, 0, (cudaStream_t)stream>>>(
      dp, dx, c_total, c_off, n0 > 0 ? n0 : (1 << 30), gap, accumulate, N, H, W, C);
  CVD_LAUNCH_OK("pool_bwd_kernel");
  return 0;
}

extern "C" int cvd_merge_up_fwd(const float* x1, int ct1, int c01, int n01, int gap1, const float* a1, const float* b1,
                                const float* x2, int ct2, int c02, int n02, int gap2, const float* a2, const float* b2,
                                float* z, int N, int H, int W, int C, void* stream)
{
  CVD_CHECK_ARG(x1 && x2 && z, "cvd_merge_up_fwd: null pointer");
  CVD_CHECK_ARG((C & 3) == 0 && (H & 1) == 0 && (W & 1) == 0, "cvd_merge_up_fwd: C %% 4, even H, W required");
  merge_up_fwd_kernel<<<ew_grid((long long)N * H * W * (C / 4)), 256, 0, (cudaStream_t)stream>>>(
      x1, ct1, c01, n01 > 0 ? n01 : (1 << 30), gap1, a1, b1, x2, ct2, c02, n02 > 0 ? n02 : (1 << 30), gap2, a2, b2, z, N, H, W, C);
  CVD_LAUNCH_OK("merge_up_fwd_kernel");
  return 0;
}

extern "C" int cvd_merge_up_bwd(const float* dz, float* dy2, int ct2, int c02, int n02, int gap2,
                                float* dy1, int ct1, int c01, int n01, int gap1, int accumulate1,
                                int N, int H, int W, int C, void* stream)
{
  CVD_CHECK_ARG(dz && dy2, "cvd_merge_up_bwd: null pointer");
  CVD_CHECK_ARG((C & 3) == 0 && (H & 1) == 0 && (W & 1) == 0, "cvd_merge_up_bwd: C %% 4, even H, W required");
  up2x_bwd_kernel<<<ew_grid((long long)N * (H / 2) * (W / 2) * (C / 4)), 256, 0, (cudaStream_t)stream>>>(
      dz, dy2, ct2, c02, n02 > 0 ? n02 : (1 << 30), gap2, dy1, ct1, c01, n01 > 0 ? n01 : (1 << 30), gap1, accumulate1, N, H, W, C);
  CVD_LAUNCH_OK("up2x_bwd_kernel");
  return 0;
}

extern "C" int cvd_image_to_nhwc4(const float* img_nchw, float* out, int N, int H, int W, void* stream)
{
  CVD_CHECK_ARG(img_nchw && out, "cvd_image_to_nhwc4: null pointer");
  image_to_nhwc4_kernel<<<ew_grid((long long)N * H * W), 256, 0, (cudaStream_t)stream>>>(img_nchw, out, N, H, W);
  CVD_LAUNCH_OK("image_to_nhwc4_kernel");
  return 0;
}

extern "C" int cvd_dlogdepth(const float* grad_depth, const float* depth, float* out4, long long n, float* dbias, void* stream)
{
  CVD_CHECK_ARG(grad_depth && depth && out4, "cvd_dlogdepth: null pointer");
  dlogdepth_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(grad_depth, depth, out4, n, dbias);
  CVD_LAUNCH_OK("dlogdepth_kernel");
  return 0;
}
