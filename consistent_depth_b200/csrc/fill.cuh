// Activation staging shared by the conv forward/dgrad and wgrad kernels:
// global NHWC fp32 (through a channel view) -> per-channel transform -> bf16 hi/lo -> shared memory in the
// UMMA SWIZZLE_NONE canonical layout  [8-channel chunk][window row][window col][16 B].
//
// kProducers (128-256) producer threads share the (pixel, 8-channel chunk) work items of a window; per-channel
// constants come from a per-CTA shared-memory cache; loads are issued UNROLL items ahead of the math so several
// 128-bit requests per thread are in flight (the window is read once: latency, not bandwidth, must be hidden).
#pragma once
#include "cvd_common.cuh"
#include "tc_common.cuh"

namespace fillns {


struct SrcView {
  const float* x; const float* dy; const float* a; const float* b; const float4* bw;
  int ct, c0, n0, gap, dy_ct, dy_c0, dy_n0, dy_gap, relu, mode, cvalid;
};

__device__ __forceinline__ int vphys(int c, int c0, int n0, int gap) { return c0 + c + (c >= n0 ? gap : 0); }

// Per-channel constants of a source view, staged ONCE per CTA into shared memory, indexed by LOGICAL channel:
// sp[0*nch + c] = a, [1*nch + c] = b, [2..4] = c0, c1, c2 (BatchNorm-backward constants).  Every fill call
// then reads its 8 channels with a handful of LDS instead of ~40 dependent global loads (which dominated the
// small tiles of the 1x1 convolutions).
__device__ __forceinline__ void stage_params(const SrcView& s, float* sp, int nch, int tid, int nthreads)
{
  for (int c = tid; c < nch; c += nthreads) {
    float a = 1.f, b = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
    if (c < s.cvalid) {
      const int pc = vphys(c, s.c0, s.n0, s.gap);
      if (s.a) { a = __ldg(s.a + pc); b = __ldg(s.b + pc); }
      if (s.mode == CVD_XF_BNBWD) { const float4 q = __ldg(s.bw + pc); c0 = q.x; c1 = q.y; c2 = q.z; }
    }
    sp[c] = a; sp[nch + c] = b; sp[2 * nch + c] = c0; sp[3 * nch + c] = c1; sp[4 * nch + c] = c2;
  }
}
__host__ __device__ inline int param_bytes(int nch) { return 5 * nch * 4; }

// Stage window pixels (r, c), r < rows, c < cols  <->  image pixel (y0 + r, x0 + c) of image n, logical channels
// [cfirst, cfirst + 8*nchunks).  Out-of-image pixels, columns outside [vx0, vx1) and channels >= cvalid are zero.
// Work item = (window pixel, 8-channel chunk), chunk fastest: consecutive threads read consecutive 32 B of one
// pixel (coalesced) and write 16 B to consecutive chunk planes.  UNROLL independent items per thread are loaded
// before any is transformed, so several 128-bit requests per thread are in flight.
template <int MODE, int UNROLL, int kProducers>
__device__ __forceinline__ void fill_window_impl(const SrcView& s, uint8_t* dst, int plane_bytes, int lo_off, int nsplit,
                                                 int n, int H, int W, int y0, int x0, int rows, int cols,
                                                 int cfirst, int nchunks, int tid, int vx0, int vx1,
                                                 const float* __restrict__ sp, int nch)
{
  const int total = rows * cols * nchunks;
  const size_t img_off = (size_t)n * H * W;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // item -> (chunk c8, window pixel hp -> (r, c)) is tracked incrementally: one division pair per call, none per item
  const int dch = kProducers % nchunks, dhp = kProducers / nchunks;
  int it = tid;
  int c8 = it % nchunks, hp = it / nchunks;
  int r = hp / cols, c = hp - r * cols;
  for (; it < total; ) {
    float4 xa[UNROLL], xb[UNROLL], da[UNROLL], db[UNROLL];
    int hps[UNROLL], c8s[UNROLL]; bool ok[UNROLL], sec[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      hps[u] = it < total ? hp : -1; c8s[u] = c8;
      const int cl = cfirst + c8 * 8;
      const int iy = y0 + r, ix = x0 + c;
      ok[u] = it < total && cl < s.cvalid && iy >= 0 && iy < H && ix >= vx0 && ix < vx1;
      sec[u] = cl + 4 < s.cvalid;
      xa[u] = z4; xb[u] = z4; da[u] = z4; db[u] = z4;
      if (ok[u]) {
        const size_t pix = img_off + (size_t)iy * W + ix;
        const float* xp = s.x + pix * s.ct + vphys(cl, s.c0, s.n0, s.gap);
        xa[u] = __ldg(reinterpret_cast<const float4*>(xp));
        if (sec[u]) xb[u] = __ldg(reinterpret_cast<const float4*>(xp + 4));
        if (MODE == CVD_XF_BNBWD) {
          const float* dp = s.dy + pix * s.dy_ct + vphys(cl, s.dy_c0, s.dy_n0, s.dy_gap);
          da[u] = __ldg(reinterpret_cast<const float4*>(dp));
          if (sec[u]) db[u] = __ldg(reinterpret_cast<const float4*>(dp + 4));
        }
      }
      // advance to this thread's next item
      it += kProducers; c8 += dch; int adv = dhp;
      if (c8 >= nchunks) { c8 -= nchunks; ++adv; }
      hp += adv; c += adv;
      while (c >= cols) { c -= cols; ++r; }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (hps[u] < 0) break;
      float v[8];
      if (ok[u]) {
        const float xv[8] = {xa[u].x, xa[u].y, xa[u].z, xa[u].w, xb[u].x, xb[u].y, xb[u].z, xb[u].w};
        const int cl = cfirst + c8s[u] * 8;
        const float4 a0 = *reinterpret_cast<const float4*>(sp + cl), a1 = *reinterpret_cast<const float4*>(sp + cl + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(sp + nch + cl), b1 = *reinterpret_cast<const float4*>(sp + nch + cl + 4);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        if (MODE == CVD_XF_AFFINE) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float t = fmaf(av[i], xv[i], bv[i]);
            v[i] = s.relu ? fmaxf(t, 0.f) : t;
          }
        } else {
          // BatchNorm(+ReLU) backward on load: y = a x + b ; g = dy * [y > 0] ; dx = c0 g - c1 - c2 y
          const float dv[8] = {da[u].x, da[u].y, da[u].z, da[u].w, db[u].x, db[u].y, db[u].z, db[u].w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float y = fmaf(av[i], xv[i], bv[i]);
            const float gq = (!s.relu || y > 0.f) ? dv[i] : 0.f;
            v[i] = sp[2 * nch + cl + i] * gq - sp[3 * nch + cl + i] - sp[4 * nch + cl + i] * y;
          }
        }
        if (!sec[u]) {
#pragma unroll
          for (int i = 4; i < 8; ++i) v[i] = 0.f;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      uint4 hi, lo;
      tc::split8(v, hi, lo);
      uint8_t* d = dst + (size_t)c8s[u] * plane_bytes + (size_t)hps[u] * 16;
      *reinterpret_cast<uint4*>(d) = hi;
      if (nsplit == 3) *reinterpret_cast<uint4*>(d + lo_off) = lo;
    }
  }
}

template <int kProducers = 128>
__device__ __forceinline__ void fill_window(const SrcView& s, uint8_t* dst, int plane_bytes, int lo_off, int nsplit,
                                            int n, int H, int W, int y0, int x0, int rows, int cols,
                                            int cfirst, int nchunks, int tid, const float* sp, int nch,
                                            int vx0 = 0, int vx1 = 1 << 30)
{
  // [vx0, vx1): image columns that may be non-zero (default: the whole image)
  vx0 = vx0 < 0 ? 0 : vx0;
  vx1 = vx1 > W ? W : vx1;
  if (s.mode == CVD_XF_AFFINE)
    fill_window_impl<CVD_XF_AFFINE, 4, kProducers>(s, dst, plane_bytes, lo_off, nsplit, n, H, W, y0, x0, rows, cols, cfirst, nchunks, tid, vx0, vx1, sp, nch);
  else
    fill_window_impl<CVD_XF_BNBWD, 2, kProducers>(s, dst, plane_bytes, lo_off, nsplit, n, H, W, y0, x0, rows, cols, cfirst, nchunks, tid, vx0, vx1, sp, nch);
}

}  // namespace fillns
