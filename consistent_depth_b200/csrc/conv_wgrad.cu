// tcgen05 weight-gradient kernel: dW[co][ci][ky][kx] = sum_{n,y,x} G[n,y,x,co] * X[n,y+ky-p,x+kx-p,ci]
//
// Replaces the wgrad half of autograd's conv backward (depth_fine_tuning.py:282 through
// hourglass.py:27,39,42,164,173).  G is the gradient wrt the conv's raw output, i.e. the
// BatchNorm+ReLU backward of the layer that follows the conv, applied ON LOAD (CVD_XF_BNBWD);
// X is the conv's input as the forward saw it (producer's BatchNorm+ReLU applied on load).
//
// GEMM view, per filter tap: D_tap[m, n] = sum_pixels Mop[pixel, m] * Nop[pixel, n]   (K = pixels)
// Both operands are "MN-major" for the tensor core (channels contiguous): the SAME smem layout as the
// forward kernel -- [8-channel chunk][row][col][16 B] -- serves as the canonical SWIZZLE_NONE MN-major
// layout (8 x-adjacent pixels x 16 B = one core matrix; next 8 pixels = LBO, next 8 channels = SBO), and a
// filter tap is again just a shifted start address into the X halo tile.
// The operand with more channels is M (64 or 128 rows), the other is N, so small-Cout layers do not waste
// the tensor core's M dimension.  Each CTA owns a group of taps (accumulators for all of them live in
// TMEM: taps x N columns <= 512) and a slab of pixel tiles; it streams the tiles through a 2-stage smem
// ring (producers: transform + bf16 hi/lo split; one thread issues the MMAs), then REDs its partial dW.
#include "cvd_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kThreads = 192;
constexpr int kProducerThreads = 128;
constexpr int TW = 16;                 // pixel-tile width = one K=16 step per tile row

struct SrcView {
  const float* x; const float* dy; const float* a; const float* b; const float4* bw;
  int ct, c0, n0, gap, dy_ct, dy_c0, dy_n0, dy_gap, relu, mode, cvalid;
};

struct WgArgs {
  SrcView g, x;
  float* dw;
  int N, H, W, cin, cout, cin_w, cout_w, k, pad;   // cin/cout padded to 16; *_w = real (dW extents)
  int nsplit;
  int x_is_m;                   // 1: M = X channels (Cin), N = G channels (Cout)
  int Mrows, mblk;              // MMA M (64/128), number of M blocks
  int Ncols;                    // MMA N
  int taps_per_group, ngroups, nslabs;
  int TH, tiles_x, tiles_y, ntiles;
  int xHP, xWP, x_plane, x_chunks, g_plane, g_chunks;
  int x_bytes, g_bytes, stage_bytes, nstages;
  int tmem_cols;
};

__device__ __forceinline__ int vphys(int c, int c0, int n0, int gap) { return c0 + c + (c >= n0 ? gap : 0); }

// Stage a (rows x cols) window of one source tensor into chunk planes at `dst`.
// Window pixel (r, c) is image pixel (y0 + r, x0 + c); out-of-image pixels are zero.
__device__ __forceinline__ void fill_planes(const SrcView& s, uint8_t* dst, int plane_bytes, int nchunks, int lo_off,
                                            int nsplit, int n, int H, int W, int y0, int x0, int rows, int cols, int tid)
{
  const int total = rows * cols * nchunks;
  const size_t img_off = (size_t)n * H * W;
  for (int it = tid; it < total; it += kProducerThreads) {
    const int c8 = it % nchunks;
    const int hp = it / nchunks;
    const int r = hp / cols, c = hp - r * cols;
    const int iy = y0 + r, ix = x0 + c;
    const int cl = c8 * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W && cl < s.cvalid) {
      const size_t pix = img_off + (size_t)iy * W + ix;
      const int pc = vphys(cl, s.c0, s.n0, s.gap);
      const float* xp = s.x + pix * s.ct + pc;
      const bool second = cl + 4 < s.cvalid;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 x0v = __ldg(reinterpret_cast<const float4*>(xp));
      float4 x1v = second ? __ldg(reinterpret_cast<const float4*>(xp + 4)) : z4;
      float xv[8] = {x0v.x, x0v.y, x0v.z, x0v.w, x1v.x, x1v.y, x1v.z, x1v.w};
      if (s.a) {
        float4 a0 = __ldg(reinterpret_cast<const float4*>(s.a + pc)), a1 = second ? __ldg(reinterpret_cast<const float4*>(s.a + pc + 4)) : z4;
        float4 b0 = __ldg(reinterpret_cast<const float4*>(s.b + pc)), b1 = second ? __ldg(reinterpret_cast<const float4*>(s.b + pc + 4)) : z4;
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) xv[i] = fmaf(av[i], xv[i], bv[i]);
      }
      if (s.mode == CVD_XF_AFFINE) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = s.relu ? fmaxf(xv[i], 0.f) : xv[i];
      } else {
        const int dc = vphys(cl, s.dy_c0, s.dy_n0, s.dy_gap);
        const float* dp = s.dy + pix * s.dy_ct + dc;
        float4 d0 = __ldg(reinterpret_cast<const float4*>(dp));
        float4 d1 = second ? __ldg(reinterpret_cast<const float4*>(dp + 4)) : z4;
        const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i >= 4 && !second) break;
          const float4 cc = __ldg(s.bw + pc + i);
          const float gq = (!s.relu || xv[i] > 0.f) ? dv[i] : 0.f;
          v[i] = cc.x * gq - cc.y - cc.z * xv[i];
        }
      }
      if (!second) {
#pragma unroll
        for (int i = 4; i < 8; ++i) v[i] = 0.f;
      }
    }
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    uint8_t* d = dst + (size_t)c8 * plane_bytes + (size_t)hp * 16;
    *reinterpret_cast<uint4*>(d) = hi;
    if (nsplit == 3) *reinterpret_cast<uint4*>(d + lo_off) = lo;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const WgArgs p)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stages = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)p.nstages * p.stage_bytes);
  uint64_t* full = bars;            // [2]
  uint64_t* empty = bars + 2;       // [2]
  uint64_t* acc_full = bars + 4;
  uint32_t* tmem_base_sh = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.y, slab = blockIdx.x;
  const int taps = p.k * p.k;
  const int t0 = group * p.taps_per_group;
  const int t1 = min(taps, t0 + p.taps_per_group);
  const int ky0 = t0 / p.k, ky1 = (t1 - 1) / p.k;
  // this CTA's tiles: slab, slab + nslabs, ...
  const int my_tiles = (p.ntiles - slab + p.nslabs - 1) / p.nslabs;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&full[i], kProducerThreads); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(acc_full, 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) { tc::tmem_alloc_dyn(tmem_base_sh, (uint32_t)p.tmem_cols); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_sh;
  const int x_lo = p.x_chunks * p.x_plane, g_lo = p.g_chunks * p.g_plane;

  if (warp == 0) {
    if (lane == 0 && my_tiles > 0) {
      const uint32_t idesc = tc::idesc_bf16(p.Mrows, p.Ncols, 1, 1);   // both operands MN-major
      const uint32_t sbase = tc::smem_u32(stages);
      for (int it = 0; it < my_tiles; ++it) {
        const int st = it % p.nstages;
        tc::mbar_wait(&full[st], (uint32_t)((it / p.nstages) & 1));
        tc::tc_fence_after();
        const uint32_t xs = sbase + (uint32_t)st * p.stage_bytes;
        const uint32_t gs = xs + (uint32_t)p.x_bytes;
        for (int tap = t0; tap < t1; ++tap) {
          const int ky = tap / p.k, kx = tap - ky * p.k;
          for (int r = 0; r < p.TH; ++r) {
            // X window pixel (r + ky - ky0, kx + c), G pixel (r, c), c = 0..15
            const uint32_t xa = xs + (uint32_t)(((r + ky - ky0) * p.xWP + kx) * 16);
            const uint32_t ga = gs + (uint32_t)(r * TW * 16);
            for (int mb = 0; mb < p.mblk; ++mb) {
              uint32_t ma, na, m_plane, n_plane, m_lo, n_lo;
              if (p.x_is_m) { ma = xa + (uint32_t)(mb * (p.Mrows / 8) * p.x_plane); na = ga; m_plane = p.x_plane; n_plane = p.g_plane; m_lo = x_lo; n_lo = g_lo; }
              else          { ma = ga + (uint32_t)(mb * (p.Mrows / 8) * p.g_plane); na = xa; m_plane = p.g_plane; n_plane = p.x_plane; m_lo = g_lo; n_lo = x_lo; }
              const uint32_t d = tmem_base + (uint32_t)(((tap - t0) * p.mblk + mb) * p.Ncols);
              const uint32_t acc = (it == 0 && r == 0) ? 0u : 1u;
              const uint64_t md_hi = tc::smem_desc(ma, 128, m_plane), nd_hi = tc::smem_desc(na, 128, n_plane);
              tc::umma_f16(d, md_hi, nd_hi, idesc, acc);
              if (p.nsplit == 3) {
                tc::umma_f16(d, tc::smem_desc(ma + m_lo, 128, m_plane), nd_hi, idesc, 1u);
                tc::umma_f16(d, md_hi, tc::smem_desc(na + n_lo, 128, n_plane), idesc, 1u);
              }
            }
          }
        }
        tc::umma_commit(&empty[st]);
      }
      tc::umma_commit(acc_full);
    }
    __syncwarp();
  } else if (warp >= 2) {
    const int tid = threadIdx.x - 64;
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % p.nstages;
      if (it >= p.nstages) tc::mbar_wait(&empty[st], (uint32_t)(((it / p.nstages) - 1) & 1));
      int t = slab + it * p.nslabs;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; const int n = t / p.tiles_y;
      const int oy = ty * p.TH, ox = tx * TW;
      uint8_t* xs = stages + (size_t)st * p.stage_bytes;
      uint8_t* gs = xs + p.x_bytes;
      fill_planes(p.x, xs, p.x_plane, p.x_chunks, x_lo, p.nsplit, n, p.H, p.W, oy + ky0 - p.pad, ox - p.pad, p.xHP, p.xWP, tid);
      fill_planes(p.g, gs, p.g_plane, p.g_chunks, g_lo, p.nsplit, n, p.H, p.W, oy, ox, p.TH, TW, tid);
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&full[st]);
    }
    if (my_tiles > 0) {
      // -------- epilogue: RED the partial dW of this CTA's taps
      tc::mbar_wait(acc_full, 0);
      tc::tc_fence_after();
      const int q = warp & 3;
      // accumulator row held by this thread (M=128: lane quarter q row 32q+lane; M=64: rows 16q+lane for lane<16)
      const int row = p.Mrows == 128 ? q * 32 + lane : q * 16 + lane;
      const bool row_ok = p.Mrows == 128 || lane < 16;
      const int kk = p.k * p.k;
      for (int tap = t0; tap < t1; ++tap) {
        for (int mb = 0; mb < p.mblk; ++mb) {
          const int m = mb * p.Mrows + row;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(((tap - t0) * p.mblk + mb) * p.Ncols);
          for (int c16 = 0; c16 < p.Ncols; c16 += 16) {
            float v[16];
            tc::tmem_ld16(taddr + (uint32_t)c16, v);
            if (!row_ok) continue;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int nn = c16 + i;
              const int ci = p.x_is_m ? m : nn, co = p.x_is_m ? nn : m;
              if (ci < p.cin_w && co < p.cout_w) atomicAdd(p.dw + ((size_t)co * p.cin_w + ci) * kk + tap, v[i]);
            }
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

SrcView make_view(const cvd_src_t* s, int cvalid) {
  SrcView v{};
  v.x = s->x; v.dy = s->dy; v.a = s->a; v.b = s->b; v.bw = reinterpret_cast<const float4*>(s->bw);
  v.ct = s->c_total; v.c0 = s->c_off; v.n0 = s->n0 > 0 ? s->n0 : (1 << 30); v.gap = s->gap;
  v.dy_ct = s->dy_ctotal; v.dy_c0 = s->dy_coff; v.dy_n0 = s->dy_n0 > 0 ? s->dy_n0 : (1 << 30); v.dy_gap = s->dy_gap;
  v.relu = s->relu; v.mode = s->mode; v.cvalid = cvalid;
  return v;
}

}  // namespace

extern "C" int cvd_conv_wgrad(const cvd_src_t* gsrc, const cvd_src_t* xsrc, float* dw_oihw,
                              int N, int H, int W, int cin, int cout, int k, int precision, void* stream)
{
  CVD_CHECK_ARG(gsrc && xsrc && dw_oihw && gsrc->x && xsrc->x, "cvd_conv_wgrad: null pointer");
  CVD_CHECK_ARG(precision == 1 || precision == 3, "cvd_conv_wgrad: precision must be 1 or 3");
  CVD_CHECK_ARG(k >= 1 && k <= 11 && (k & 1), "cvd_conv_wgrad: k=%d unsupported", k);
  WgArgs p{};
  p.g = make_view(gsrc, round_up(cout, 4)); p.x = make_view(xsrc, round_up(cin, 4));
  p.dw = dw_oihw; p.N = N; p.H = H; p.W = W; p.k = k; p.pad = (k - 1) / 2;
  p.cin_w = cin; p.cout_w = cout; p.cin = round_up(cin, 16); p.cout = round_up(cout, 16);
  p.nsplit = precision;
  CVD_CHECK_ARG(p.cin <= 256 && p.cout <= 256, "cvd_conv_wgrad: channel counts above 256 unsupported");
  p.x_is_m = p.cin >= p.cout;
  const int cm = p.x_is_m ? p.cin : p.cout, cn = p.x_is_m ? p.cout : p.cin;
  p.Mrows = cm <= 64 ? 64 : 128;
  p.mblk = (cm + p.Mrows - 1) / p.Mrows;
  p.Ncols = cn;
  CVD_CHECK_ARG(p.mblk * p.Ncols <= 512, "cvd_conv_wgrad: accumulator does not fit TMEM");
  const int taps = k * k;
  p.taps_per_group = 512 / (p.mblk * p.Ncols);
  if (p.taps_per_group > taps) p.taps_per_group = taps;
  p.ngroups = (taps + p.taps_per_group - 1) / p.taps_per_group;
  // smem: choose the tile height so two stages fit
  const int budget = 200 * 1024;
  p.x_chunks = p.cin / 8; p.g_chunks = p.cout / 8;
  // M operand is read with Mrows/8 chunk planes per block: make sure those reads stay inside the stage
  const int max_ky_span = (p.taps_per_group + k - 2) / k + 1;     // rows of taps a group can touch
  int TH = 0;
  for (int th = 16; th >= 1; th >>= 1) {
    if (th > round_up(H, 1) && th > 1) continue;
    const int xHP = th + (max_ky_span - 1 < k - 1 ? max_ky_span - 1 : k - 1), xWP = TW + k - 1;
    const int xpl = round_up(xHP * xWP * 16, 128) + 16, gpl = round_up(th * TW * 16, 128) + 16;
    const int m_chunks_read = p.mblk * (p.Mrows / 8);
    const int xb = xpl * (p.x_is_m ? (m_chunks_read > p.x_chunks ? m_chunks_read : p.x_chunks) : p.x_chunks) * (precision == 3 ? 2 : 1);
    const int gb = gpl * (!p.x_is_m ? (m_chunks_read > p.g_chunks ? m_chunks_read : p.g_chunks) : p.g_chunks) * (precision == 3 ? 2 : 1);
    if (2 * (xb + gb) + 1024 <= budget) {
      TH = th; p.xHP = xHP; p.xWP = xWP; p.x_plane = xpl; p.g_plane = gpl; p.x_bytes = xb; p.g_bytes = gb;
      break;
    }
  }
  CVD_CHECK_ARG(TH > 0, "cvd_conv_wgrad: no tile fits shared memory (cin=%d cout=%d k=%d)", cin, cout, k);
  p.TH = TH; p.stage_bytes = p.x_bytes + p.g_bytes; p.nstages = 2;
  // lo planes start after the REAL chunk planes; the padded M reads may run into them (garbage rows, ignored)
  p.tiles_x = (W + TW - 1) / TW; p.tiles_y = (H + TH - 1) / TH;
  p.ntiles = N * p.tiles_x * p.tiles_y;
  int slabs = (cvd_num_sms() + p.ngroups - 1) / p.ngroups;
  if (slabs > p.ntiles) slabs = p.ntiles;
  if (slabs < 1) slabs = 1;
  p.nslabs = slabs;
  int cols = p.taps_per_group * p.mblk * p.Ncols, pw = 32;
  while (pw < cols) pw <<= 1;
  p.tmem_cols = pw;
  const size_t smem = (size_t)p.nstages * p.stage_bytes + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) return cvd_fail("cvd_conv_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  wgrad_tc_kernel<<<dim3(p.nslabs, p.ngroups), kThreads, smem, (cudaStream_t)stream>>>(p);
  CVD_LAUNCH_OK("wgrad_tc_kernel");
  return 0;
}
