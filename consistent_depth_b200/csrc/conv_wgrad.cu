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
#include "fill.cuh"
#include <cstdlib>

namespace {

constexpr int kIssuers = 4;                // MMA issuer warps (taps split between them)
constexpr int kThreads = 32 * (kIssuers + 8);   // warps [0,kIssuers) issue, then 8 producer warps (first 4 also epilogue)
constexpr int kProducerThreads = 256;
constexpr int TW = 16;                 // pixel-tile width = one K=16 step per tile row

using fillns::SrcView;

struct WgArgs {
  SrcView g, x;
  float* dw;
  int N, H, W, cin, cout, cin_w, cout_w, k, pad;   // cin/cout padded to 16; *_w = real (dW extents)
  int dw_ci_stride;                                // Cin of the whole dW tensor (>= cin_w when this launch is a chunk)
  int zb_ci, zb_co, tot_cin, tot_cout;             // blockIdx.z enumerates 256 x 256 blocks of a larger dW (1, 1: single block)
  int group;                                       // > 0: grouped conv chunk, dW is (Cout, group, k, k): only same-group entries are kept
  int red_v4;                                      // 1x1, X is M: transposed epilogue issuing coalesced 4-wide REDs
  int diag;                                        // > 0: blockIdx.z enumerates the diagonal chunks of a grouped conv (chunk = cin_w channels)
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

template <int MBLK, int NSPLIT>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const WgArgs p)
{
  // blockIdx.z = (Cout block, Cin block) of a dW larger than one launch's 256 x 256 limit: same tiling, shifted views
  const int cob = (int)blockIdx.z / p.zb_ci, cib = (int)blockIdx.z - cob * p.zb_ci;
  SrcView xv = p.x, gv = p.g;
  int cin_w = p.cin_w, cout_w = p.cout_w;
  float* dwp = p.dw;
  if (p.diag > 0) {
    // grouped conv: chunk z of the block-diagonal weight -- both operands and the dW rows shift by z chunks
    const int c0 = (int)blockIdx.z * cin_w;
    xv.c0 += c0; xv.dy_c0 += c0; gv.c0 += c0; gv.dy_c0 += c0;
    dwp += (size_t)c0 * p.group * p.k * p.k;
  } else if (p.zb_ci * p.zb_co > 1) {
    const int ci0 = cib * 256, co0 = cob * 256;
    xv.c0 += ci0; xv.dy_c0 += ci0; gv.c0 += co0; gv.dy_c0 += co0;       // (chunked launches require gap-free views)
    cin_w = min(256, p.tot_cin - ci0); cout_w = min(256, p.tot_cout - co0);
    xv.cvalid = (cin_w + 3) & ~3; gv.cvalid = (cout_w + 3) & ~3;
    dwp += ((size_t)co0 * p.dw_ci_stride + ci0) * p.k * p.k;
  }
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stages = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)p.nstages * p.stage_bytes);
  uint64_t* full = bars;            // [2]
  uint64_t* empty = bars + 2;       // [2]
  uint64_t* acc_full = bars + 4;
  uint32_t* tmem_base_sh = reinterpret_cast<uint32_t*>(bars + 5);
  float* xparams = reinterpret_cast<float*>(bars + 16);           // per-channel constants of X (5 x cin) then G (5 x cout)
  float* gparams = xparams + 5 * p.cin;
  fillns::stage_params(xv, xparams, p.cin, threadIdx.x, kThreads);
  fillns::stage_params(gv, gparams, p.cout, threadIdx.x, kThreads);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.y, slab = blockIdx.x;
  const int taps = p.k * p.k;
  const int t0 = group * p.taps_per_group;
  const int t1 = min(taps, t0 + p.taps_per_group);
  const int ky0 = t0 / p.k;
  // this CTA's tiles: slab, slab + nslabs, ...
  const int my_tiles = (p.ntiles - slab + p.nslabs - 1) / p.nslabs;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&full[i], kProducerThreads); tc::mbar_init(&empty[i], kIssuers); }
    tc::mbar_init(acc_full, kIssuers);
    tc::mbar_fence_init();
  }
  if (warp == 0) { tc::tmem_alloc_dyn(tmem_base_sh, (uint32_t)p.tmem_cols); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_sh;
  const int x_lo = p.x_chunks * p.x_plane, g_lo = p.g_chunks * p.g_plane;

  if (warp < kIssuers) {
    // SEVERAL MMA issuer warps (each: uniform loop nest, one elected lane issues): the taps of this CTA are split
    // between them -- a single thread cannot issue these small MMAs fast enough to keep the tensor pipe busy
    // (measured: 1 -> 2 issuers = 1.56x on the 64->16 11x11 layer).
    const int th = (t1 - t0 + kIssuers - 1) / kIssuers;
    const int w_t0 = min(t1, t0 + warp * th), w_t1 = min(t1, t0 + (warp + 1) * th);
    const uint32_t idesc = tc::idesc_bf16(p.Mrows, p.Ncols, 1, 1);   // both operands MN-major
    const uint32_t sbase = tc::smem_u32(stages);
    const uint32_t m_plane = p.x_is_m ? p.x_plane : p.g_plane, n_plane = p.x_is_m ? p.g_plane : p.x_plane;
    const uint32_t m_lo = p.x_is_m ? x_lo : g_lo, n_lo = p.x_is_m ? g_lo : x_lo;
    const uint64_t mdesc0 = tc::smem_desc_base(128, m_plane), ndesc0 = tc::smem_desc_base(128, n_plane);
    const uint32_t mb_stride = (uint32_t)(p.Mrows / 8) * m_plane;
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % p.nstages;
      tc::mbar_wait(&full[st], (uint32_t)((it / p.nstages) & 1));
      tc::tc_fence_after();
      const uint32_t xs = sbase + (uint32_t)st * p.stage_bytes;
      const uint32_t gs = xs + (uint32_t)p.x_bytes;
      if (tc::elect_one()) {
        // taps t0..t1-1 in raster order without divisions; per tap a straight run of TH x MBLK x NSPLIT MMAs
        int ky = w_t0 / p.k, kx = w_t0 - ky * p.k;
        uint32_t dtap = tmem_base + (uint32_t)((w_t0 - t0) * MBLK * p.Ncols);
        const uint32_t x_row = (uint32_t)p.xWP * 16;
        for (int tap = w_t0; tap < w_t1; ++tap) {
          // X window pixel (r + ky - ky0, kx + c), G pixel (r, c), c = 0..15
          uint32_t xa = xs + (uint32_t)(((ky - ky0) * p.xWP + kx) * 16);
          uint32_t ga = gs;
          uint32_t acc = it ? 1u : 0u;
          for (int r = 0; r < p.TH; ++r, xa += x_row, ga += TW * 16) {
            const uint32_t ma0 = p.x_is_m ? xa : ga, na = p.x_is_m ? ga : xa;
            const uint64_t nd_hi = tc::smem_desc_at(ndesc0, na), nd_lo = tc::smem_desc_at(ndesc0, na + n_lo);
#pragma unroll
            for (int mb = 0; mb < MBLK; ++mb) {
              const uint32_t ma = ma0 + (uint32_t)mb * mb_stride;
              const uint32_t d = dtap + (uint32_t)(mb * p.Ncols);
              const uint64_t md_hi = tc::smem_desc_at(mdesc0, ma);
              tc::umma_f16(d, md_hi, nd_hi, idesc, acc);
              if (NSPLIT == 3) {
                tc::umma_f16(d, tc::smem_desc_at(mdesc0, ma + m_lo), nd_hi, idesc, 1u);
                tc::umma_f16(d, md_hi, nd_lo, idesc, 1u);
              }
            }
            acc = 1u;
          }
          dtap += (uint32_t)(MBLK * p.Ncols);
          if (++kx == p.k) { kx = 0; ++ky; }
        }
        tc::umma_commit(&empty[st]);
      }
      __syncwarp();
    }
    if (my_tiles > 0 && tc::elect_one()) tc::umma_commit(acc_full);
    __syncwarp();
  } else {
    const int tid = threadIdx.x - 32 * kIssuers;
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % p.nstages;
      if (it >= p.nstages) tc::mbar_wait(&empty[st], (uint32_t)(((it / p.nstages) - 1) & 1));
      int t = slab + it * p.nslabs;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; const int n = t / p.tiles_y;
      const int oy = ty * p.TH, ox = tx * TW;
      uint8_t* xs = stages + (size_t)st * p.stage_bytes;
      uint8_t* gs = xs + p.x_bytes;
      fillns::fill_window<kProducerThreads>(xv, xs, p.x_plane, x_lo, p.nsplit, n, p.H, p.W, oy + ky0 - p.pad, ox - p.pad, p.xHP, p.xWP, 0, p.x_chunks, tid, xparams, p.cin);
      fillns::fill_window<kProducerThreads>(gv, gs, p.g_plane, g_lo, p.nsplit, n, p.H, p.W, oy, ox, p.TH, TW, 0, p.g_chunks, tid, gparams, p.cout);
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&full[st]);
    }
    if (my_tiles > 0 && warp < kIssuers + 4) {
      // -------- epilogue: RED the partial dW of this CTA's taps
      tc::mbar_wait(acc_full, 0);
      tc::tc_fence_after();
      const int q = warp & 3;
      // accumulator row held by this thread (M=128: lane quarter q row 32q+lane; M=64: rows 16q+lane for lane<16)
      const int row = p.Mrows == 128 ? q * 32 + lane : q * 16 + lane;
      const bool row_ok = p.Mrows == 128 || lane < 16;
      const int kk = p.k * p.k;
      if (p.red_v4) {
        // 1x1 conv, X = M: a thread holds 16 output channels of ONE input channel (stride dw_ci_stride in dW), the
        // warp 32 consecutive input channels.  Transposing the 32 x 16 block through shared memory (the stage ring is
        // idle by now) lets each lane add 4 CONSECUTIVE input channels with one red.v4: a quarter of the RED lane
        // operations, still 128 B-coalesced (measured: the scalar-RED epilogue was ~40 us of a 78 us launch).
        float* tr = reinterpret_cast<float*>(stages) + q * (16 * 36);
        const int rr = lane >> 3, cc = (lane & 7) * 4;
        for (int mb = 0; mb < p.mblk; ++mb) {
          const int m0 = mb * 128 + q * 32;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * p.Ncols);
          for (int c16 = 0; c16 < p.Ncols; c16 += 16) {
            float v[16];
            tc::tmem_ld16(taddr + (uint32_t)c16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) tr[i * 36 + lane] = v[i];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int co = c16 + rr + 4 * j, ci = m0 + cc;
              const float4 o = *reinterpret_cast<const float4*>(tr + (rr + 4 * j) * 36 + cc);
              if (co < cout_w && ci < cin_w)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                             ::"l"(dwp + (size_t)co * p.dw_ci_stride + ci), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
            }
            __syncwarp();
          }
        }
      } else
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
              if (ci < cin_w && co < cout_w) {
                if (p.group == 0) atomicAdd(dwp + ((size_t)co * p.dw_ci_stride + ci) * kk + tap, v[i]);
                else if (ci / p.group == co / p.group) atomicAdd(dwp + ((size_t)co * p.group + ci % p.group) * kk + tap, v[i]);
              }
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

int cvd_conv_wgrad_kx(const cvd_src_t* gsrc, const cvd_src_t* xsrc, float* dw_oihw,
                      int N, int H, int W, int cin, int cout, int k, int precision, void* stream);   // conv_wgrad_kx.cu

static int wgrad_impl(const cvd_src_t* gsrc, const cvd_src_t* xsrc, float* dw_oihw,
                      int N, int H, int W, int cin, int cout, int k, int precision, int dw_ci_stride, void* stream, int group = 0,
                      int tot_cin = 0, int tot_cout = 0, int diag_chunks = 0);

extern "C" int cvd_conv_wgrad_grouped(const cvd_src_t* gsrc, const cvd_src_t* xsrc, float* dw_ogkk,
                                      int N, int H, int W, int c, int group_size, int k, int precision, void* stream)
{
  CVD_CHECK_ARG(gsrc && xsrc && dw_ogkk && gsrc->x && xsrc->x, "cvd_conv_wgrad_grouped: null pointer");
  CVD_CHECK_ARG(c > 0 && c <= 256 && group_size > 0 && c % group_size == 0, "cvd_conv_wgrad_grouped: c=%d group_size=%d", c, group_size);
  return wgrad_impl(gsrc, xsrc, dw_ogkk, N, H, W, c, c, k, precision, -1, stream, group_size);
}

extern "C" int cvd_conv_wgrad_grouped_chunks(const cvd_src_t* gsrc, const cvd_src_t* xsrc, float* dw_ogkk,
                                             int N, int H, int W, int c, int nchunks, int group_size, int k, int precision,
                                             void* stream)
{
  CVD_CHECK_ARG(gsrc && xsrc && dw_ogkk && gsrc->x && xsrc->x, "cvd_conv_wgrad_grouped_chunks: null pointer");
  CVD_CHECK_ARG(c > 0 && c <= 256 && c % 16 == 0 && group_size > 0 && c % group_size == 0 && nchunks >= 1 && nchunks <= 65535,
                "cvd_conv_wgrad_grouped_chunks: c=%d group_size=%d nchunks=%d", c, group_size, nchunks);
  CVD_CHECK_ARG(gsrc->gap == 0 && xsrc->gap == 0 && gsrc->dy_gap == 0 && xsrc->dy_gap == 0,
                "cvd_conv_wgrad_grouped_chunks: gap-free views required");
  return wgrad_impl(gsrc, xsrc, dw_ogkk, N, H, W, c, c, k, precision, -1, stream, group_size, 0, 0, nchunks);
}

extern "C" int cvd_conv_wgrad(const cvd_src_t* gsrc, const cvd_src_t* xsrc, float* dw_oihw,
                              int N, int H, int W, int cin, int cout, int k, int precision, void* stream)
{
  CVD_CHECK_ARG(gsrc && xsrc && dw_oihw && gsrc->x && xsrc->x, "cvd_conv_wgrad: null pointer");
  if (cin > 256 || cout > 256) {
    // channel counts above one launch's limits: ONE launch whose blockIdx.z enumerates the 256 x 256 blocks of dW
    // (operands read through shifted views); the pixel tiles are split over fewer slabs so the grid still fills the GPU
    // and each dW element receives fewer RED contributions
    CVD_CHECK_ARG(gsrc->gap == 0 && xsrc->gap == 0 && gsrc->dy_gap == 0 && xsrc->dy_gap == 0,
                  "cvd_conv_wgrad: channel counts above 256 need gap-free views");
    return wgrad_impl(gsrc, xsrc, dw_oihw, N, H, W, cin < 256 ? cin : 256, cout < 256 ? cout : 256, k, precision,
                      cin, stream, 0, cin, cout);
  }
  return wgrad_impl(gsrc, xsrc, dw_oihw, N, H, W, cin, cout, k, precision, cin, stream);
}

static int wgrad_impl(const cvd_src_t* gsrc, const cvd_src_t* xsrc, float* dw_oihw,
                      int N, int H, int W, int cin, int cout, int k, int precision, int dw_ci_stride, void* stream, int group,
                      int tot_cin, int tot_cout, int diag_chunks)
{
  CVD_CHECK_ARG(precision == 1 || precision == 3, "cvd_conv_wgrad: precision must be 1 or 3");
  CVD_CHECK_ARG(k >= 1 && k <= 11 && (k & 1), "cvd_conv_wgrad: k=%d unsupported", k);
  // kx-fused variant (conv_wgrad_kx.cu): measured faster only when the shifted operand is a single 8-channel chunk
  // (the 3-channel image of the first layer: 0.45 ms vs 1.58 ms); elsewhere its extra staging (more tap groups,
  // wider halo) makes it producer-bound, so the per-tap kernel stays the default.  CVD_WGRAD_KX=all forces it.
  const char* kxenv = getenv("CVD_WGRAD_KX");
  const bool kx_all = kxenv && kxenv[0] == 'a';
  const int cmin = cin < cout ? cin : cout;
  if (k >= 3 && dw_ci_stride == cin && !getenv("CVD_WGRAD_PER_TAP") && (kx_all || cmin <= 8)) {
    const int rc = cvd_conv_wgrad_kx(gsrc, xsrc, dw_oihw, N, H, W, cin, cout, k, precision, stream);
    if (rc != 2) return rc;
  }
  WgArgs p{};
  p.g = make_view(gsrc, round_up(cout, 4)); p.x = make_view(xsrc, round_up(cin, 4));
  p.dw = dw_oihw; p.N = N; p.H = H; p.W = W; p.k = k; p.pad = (k - 1) / 2;
  p.cin_w = cin; p.cout_w = cout; p.cin = round_up(cin, 16); p.cout = round_up(cout, 16);
  p.dw_ci_stride = dw_ci_stride; p.group = group;
  p.tot_cin = tot_cin > 0 ? tot_cin : cin; p.tot_cout = tot_cout > 0 ? tot_cout : cout;
  p.zb_ci = (p.tot_cin + 255) / 256; p.zb_co = (p.tot_cout + 255) / 256;
  if (tot_cin <= 0) { p.zb_ci = 1; p.zb_co = 1; }
  p.diag = diag_chunks;
  const int zb = diag_chunks > 0 ? diag_chunks : p.zb_ci * p.zb_co;
  p.nsplit = precision;
  CVD_CHECK_ARG(p.cin <= 256 && p.cout <= 256, "cvd_conv_wgrad: channel counts above 256 unsupported");
  p.x_is_m = p.cin >= p.cout;
  const int cm = p.x_is_m ? p.cin : p.cout, cn = p.x_is_m ? p.cout : p.cin;
  p.Mrows = cm <= 64 ? 64 : 128;
  p.red_v4 = k == 1 && p.x_is_m && group == 0 && p.Mrows == 128 && (cin & 3) == 0 && (dw_ci_stride & 3) == 0 &&
             (p.tot_cin & 3) == 0 && (reinterpret_cast<uintptr_t>(dw_oihw) & 15) == 0 && !getenv("CVD_WG_NO_V4");
  p.mblk = (cm + p.Mrows - 1) / p.Mrows;
  p.Ncols = cn;
  CVD_CHECK_ARG(p.mblk * p.Ncols <= 512 && p.mblk <= 2, "cvd_conv_wgrad: accumulator does not fit TMEM");
  const int taps = k * k;
  p.taps_per_group = 512 / (p.mblk * p.Ncols);
  if (p.taps_per_group > taps) p.taps_per_group = taps;
  p.ngroups = (taps + p.taps_per_group - 1) / p.taps_per_group;
  // smem: choose the tile height so two stages fit
  const int budget = 218 * 1024;
  p.x_chunks = p.cin / 8; p.g_chunks = p.cout / 8;
  // M operand is read with Mrows/8 chunk planes per block: make sure those reads stay inside the stage
  const int max_ky_span = (p.taps_per_group + k - 2) / k + 1;     // rows of taps a group can touch
  int TH = 0;
  const int th_max = getenv("CVD_WG_TH") ? atoi(getenv("CVD_WG_TH")) : 16;
  for (int th = 16; th >= 1; th >>= 1) {
    if (th > th_max && th > 1) continue;
    if (th > round_up(H, 1) && th > 1) continue;
    const int xHP = th + (max_ky_span - 1 < k - 1 ? max_ky_span - 1 : k - 1), xWP = TW + k - 1;
    const int xpl = round_up(xHP * xWP * 16, 128) + 16, gpl = round_up(th * TW * 16, 128) + 16;
    const int m_chunks_read = p.mblk * (p.Mrows / 8);
    const int xb = xpl * (p.x_is_m ? (m_chunks_read > p.x_chunks ? m_chunks_read : p.x_chunks) : p.x_chunks) * (precision == 3 ? 2 : 1);
    const int gb = gpl * (!p.x_is_m ? (m_chunks_read > p.g_chunks ? m_chunks_read : p.g_chunks) : p.g_chunks) * (precision == 3 ? 2 : 1);
    if (2 * (xb + gb) + 1024 + fillns::param_bytes(p.cin + p.cout) <= budget) {
      TH = th; p.xHP = xHP; p.xWP = xWP; p.x_plane = xpl; p.g_plane = gpl; p.x_bytes = xb; p.g_bytes = gb;
      break;
    }
  }
  CVD_CHECK_ARG(TH > 0, "cvd_conv_wgrad: no tile fits shared memory (cin=%d cout=%d k=%d)", cin, cout, k);
  p.TH = TH; p.stage_bytes = p.x_bytes + p.g_bytes; p.nstages = 2;
  if (p.nstages * p.stage_bytes < 4 * 16 * 36 * 4) p.red_v4 = 0;     // transpose tiles live in the idle stage ring
  // lo planes start after the REAL chunk planes; the padded M reads may run into them (garbage rows, ignored)
  p.tiles_x = (W + TW - 1) / TW; p.tiles_y = (H + TH - 1) / TH;
  p.ntiles = N * p.tiles_x * p.tiles_y;
  // one CTA per SM: never more CTAs than SMs (a second wave of a few CTAs doubles the launch time)
  int slabs = cvd_num_sms() / (p.ngroups * zb);
  if (const char* e = getenv("CVD_WG_SLABS")) { if (atoi(e) > 0) slabs = atoi(e); }
  if (slabs > p.ntiles) slabs = p.ntiles;
  if (slabs < 1) slabs = 1;
  p.nslabs = slabs;
  int cols = p.taps_per_group * p.mblk * p.Ncols, pw = 32;
  while (pw < cols) pw <<= 1;
  p.tmem_cols = pw;
  const size_t smem = (size_t)p.nstages * p.stage_bytes + 1024 + fillns::param_bytes(p.cin + p.cout);
  cudaError_t e = cudaSuccess;
#define CVD_WG_LAUNCH(MB, NS)                                                                                 \
  do {                                                                                                        \
    static bool cfg = false;                                                                                  \
    if (!cfg) {                                                                                               \
      e = cudaFuncSetAttribute(wgrad_tc_kernel<MB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)); \
      cfg = true;                                                                                             \
    }                                                                                                         \
    if (e == cudaSuccess) wgrad_tc_kernel<MB, NS><<<dim3(p.nslabs, p.ngroups, zb), kThreads, smem, (cudaStream_t)stream>>>(p); \
  } while (0)
  if (p.mblk == 1) { if (precision == 3) CVD_WG_LAUNCH(1, 3); else CVD_WG_LAUNCH(1, 1); }
  else             { if (precision == 3) CVD_WG_LAUNCH(2, 3); else CVD_WG_LAUNCH(2, 1); }
  if (e != cudaSuccess) return cvd_fail("cvd_conv_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  CVD_LAUNCH_OK("wgrad_tc_kernel");
  return 0;
}
