// Activation staging shared by the conv forward/dgrad and wgrad kernels:
// global NHWC fp32 (through a channel view) -> per-channel transform -> bf16 hi/lo -> shared memory in the
// UMMA SWIZZLE_NONE canonical layout  [8-channel chunk][window row][window col][16 B].
//
// kProducers (128-256) producer threads; a thread owns ONE 8-channel chunk for the whole call (kProducers % nchunks == 0), so the
// per-channel constants (BatchNorm scale/shift, backward constants) are loaded into registers once, and
// walks the window pixels with stride 128/nchunks.  Loads are issued UNROLL items ahead of the math so
// several 128-bit requests per thread are in flight (the window is read once; latency, not bandwidth, is
// what has to be hidden).
#pragma once
#include "cvd_common.cuh"
#include "tc_common.cuh"

namespace fillns {


struct SrcView {
  const float* x; const float* dy; const float* a; const float* b; const float4* bw;
  int ct, c0, n0, gap, dy_ct, dy_c0, dy_n0, dy_gap, relu, mode, cvalid;
};

__device__ __forceinline__ int vphys(int c, int c0, int n0, int gap) { return c0 + c + (c >= n0 ? gap : 0); }

// Stage window pixels (r, c), r < rows, c < cols  <->  image pixel (y0 + r, x0 + c) of image n, logical channels
// [cfirst, cfirst + 8*nchunks).  Out-of-image pixels and channels >= cvalid are zero.  nchunks in {1,2,4,8}.
template <int MODE, int UNROLL, int kProducers>
__device__ __forceinline__ void fill_window_impl(const SrcView& s, uint8_t* dst, int plane_bytes, int lo_off, int nsplit,
                                                 int n, int H, int W, int y0, int x0, int rows, int cols,
                                                 int cfirst, int nchunks, int tid, int vx0, int vx1)
{
  const int c8 = tid % nchunks;
  const int ppi = kProducers / nchunks;            // window pixels advanced per iteration
  const int npix = rows * cols;
  const int cl = cfirst + c8 * 8;
  const bool ch_ok = cl < s.cvalid;
  const bool second = cl + 4 < s.cvalid;
  const int pc = vphys(cl, s.c0, s.n0, s.gap);
  const int dc = MODE == CVD_XF_BNBWD ? vphys(cl, s.dy_c0, s.dy_n0, s.dy_gap) : 0;
  // per-channel constants, once per thread
  float av[8], bv[8];
  float c0v[8], c1v[8], c2v[8];
  const bool has_ab = s.a != nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) { av[i] = 1.f; bv[i] = 0.f; c0v[i] = 0.f; c1v[i] = 0.f; c2v[i] = 0.f; }
  if (ch_ok) {
    const int nv = second ? 8 : 4;
    if (has_ab) {
#pragma unroll
      for (int i = 0; i < 8; ++i) if (i < nv) { av[i] = __ldg(s.a + pc + i); bv[i] = __ldg(s.b + pc + i); }
    }
    if (MODE == CVD_XF_BNBWD) {
#pragma unroll
      for (int i = 0; i < 8; ++i) if (i < nv) { const float4 c = __ldg(s.bw + pc + i); c0v[i] = c.x; c1v[i] = c.y; c2v[i] = c.z; }
    }
  }
  const size_t img_off = (size_t)n * H * W;
  uint8_t* dplane = dst + (size_t)c8 * plane_bytes;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

  int hp = tid / nchunks;
  int r = hp / cols, c = hp - r * cols;
  const int dr = ppi / cols, dcn = ppi - dr * cols;          // per-iteration row/col increments
  for (; hp < npix; ) {
    float4 xa[UNROLL], xb[UNROLL], da[UNROLL], db[UNROLL];
    int hps[UNROLL]; bool ok[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      hps[u] = hp;
      const int iy = y0 + r, ix = x0 + c;
      ok[u] = hp < npix && ch_ok && iy >= 0 && iy < H && ix >= vx0 && ix < vx1;
      xa[u] = z4; xb[u] = z4; da[u] = z4; db[u] = z4;
      if (ok[u]) {
        const size_t pix = img_off + (size_t)iy * W + ix;
        const float* xp = s.x + pix * s.ct + pc;
        xa[u] = __ldg(reinterpret_cast<const float4*>(xp));
        if (second) xb[u] = __ldg(reinterpret_cast<const float4*>(xp + 4));
        if (MODE == CVD_XF_BNBWD) {
          const float* dp = s.dy + pix * s.dy_ct + dc;
          da[u] = __ldg(reinterpret_cast<const float4*>(dp));
          if (second) db[u] = __ldg(reinterpret_cast<const float4*>(dp + 4));
        }
      }
      hp += ppi; r += dr; c += dcn;
      if (c >= cols) { c -= cols; ++r; }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (hps[u] >= npix) break;
      float v[8];
      const float xv[8] = {xa[u].x, xa[u].y, xa[u].z, xa[u].w, xb[u].x, xb[u].y, xb[u].z, xb[u].w};
      if (ok[u]) {
        if (MODE == CVD_XF_AFFINE) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float t = has_ab ? fmaf(av[i], xv[i], bv[i]) : xv[i];
            v[i] = s.relu ? fmaxf(t, 0.f) : t;
          }
        } else {
          // BatchNorm(+ReLU) backward on load: y = a x + b ; g = dy * [y > 0] ; dx = c0 g - c1 - c2 y
          const float dv[8] = {da[u].x, da[u].y, da[u].z, da[u].w, db[u].x, db[u].y, db[u].z, db[u].w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float y = fmaf(av[i], xv[i], bv[i]);
            const float gq = (!s.relu || y > 0.f) ? dv[i] : 0.f;
            v[i] = c0v[i] * gq - c1v[i] - c2v[i] * y;
          }
        }
        if (!second) {
#pragma unroll
          for (int i = 4; i < 8; ++i) v[i] = 0.f;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      uint4 hi, lo;
      tc::split8(v, hi, lo);
      uint8_t* d = dplane + (size_t)hps[u] * 16;
      *reinterpret_cast<uint4*>(d) = hi;
      if (nsplit == 3) *reinterpret_cast<uint4*>(d + lo_off) = lo;
    }
  }
}

template <int kProducers = 128>
__device__ __forceinline__ void fill_window(const SrcView& s, uint8_t* dst, int plane_bytes, int lo_off, int nsplit,
                                            int n, int H, int W, int y0, int x0, int rows, int cols,
                                            int cfirst, int nchunks, int tid, int vx0 = 0, int vx1 = 1 << 30)
{
  // [vx0, vx1): image columns that may be non-zero (default: the whole image)
  vx0 = vx0 < 0 ? 0 : vx0;
  vx1 = vx1 > W ? W : vx1;
  // chunk counts that do not divide 128 (e.g. 26 chunks = 208 channels) are staged in power-of-two batches
  int done = 0;
  while (done < nchunks) {
    int batch = 8;
    while (batch > nchunks - done) batch >>= 1;
    if (s.mode == CVD_XF_AFFINE)
      fill_window_impl<CVD_XF_AFFINE, 4, kProducers>(s, dst + (size_t)done * plane_bytes, plane_bytes, lo_off, nsplit, n, H, W, y0, x0, rows, cols,
                                         cfirst + done * 8, batch, tid, vx0, vx1);
    else
      fill_window_impl<CVD_XF_BNBWD, 2, kProducers>(s, dst + (size_t)done * plane_bytes, plane_bytes, lo_off, nsplit, n, H, W, y0, x0, rows, cols,
                                        cfirst + done * 8, batch, tid, vx0, vx1);
    done += batch;
  }
}

}  // namespace fillns
