// tcgen05 weight gradient for k x k filters (k >= 3) with the kx taps FUSED INTO THE GEMM N DIMENSION.
//
//   dW[co][ci][ky][kx] = sum_p G[p, co] * X[p + (ky - pad, kx - pad), ci]
//
// The per-tap formulation (conv_wgrad.cu) issues one tiny MMA (N = 16/32 columns) per tap and 16 pixels and is
// bound by MMA issue / operand fetch.  Here ONE MMA covers all k horizontal taps of an 8-channel chunk:
// the N operand is read MN-major from the [chunk][row][col][16 B] tile with a *stride of 16 bytes between its
// 8-channel groups* (descriptor SBO = 16 B): N-group j starts one pixel later than group j-1, i.e. group j IS
// the tile shifted by j pixels.  The hardware only computes addresses, so the k shifted views overlap in shared
// memory and nothing is copied:   D[m, j*8 + c] = sum_q Mop[q, m] * Nop[q + j, c]   (N = 8k columns per chunk).
// The operand with MORE channels is M (unshifted); the one with fewer channels is the shifted N operand:
//   Cout > Cin  : M = G (tile), N = X halo window, j = kx
//   Cin >= Cout : M = X halo window, N = G zero-padded tile, j = k-1-kx, K runs over the tile plus the halo columns
// ky stays a start-address shift of the X window (as in the forward kernel).  Accumulators: one per (ky, N chunk),
// 8k TMEM columns each; a CTA owns a contiguous range of them and a slab of pixel tiles (2-stage smem ring,
// same producer / issuer / epilogue roles as conv_wgrad.cu); partial dW is RED-accumulated.
#include "cvd_common.cuh"
#include "tc_common.cuh"
#include "fill.cuh"
#include <cstdlib>

namespace {

constexpr int kThreads = 320;            // warp 0 issuer, warp 1 idle, warps 2-9 producers (2-5 also epilogue)
constexpr int kProducerThreads = 256;

using fillns::SrcView;

struct KxArgs {
  SrcView g, x;
  float* dw;
  int N, H, W, cin_w, cout_w, k, pad, cin_p, cout_p;
  int nsplit;
  int n_is_g;                 // 1: N operand = G (Cin >= Cout), 0: N operand = X
  int Mrows;                  // 64 or 128 (channels of the M operand, padded)
  int ncn;                    // 8-channel chunks of the N operand
  int NC;                     // TMEM columns per accumulator (8k, rounded to 16 when M = 128)
  int acc_per_cta, ngroups, nslabs, nacc;
  int TH, TW, tiles_x, tiles_y, ntiles;
  int ksteps;                 // 16-pixel K steps per tile row
  int x_rows, x_cols, x_plane, x_chunks;   // X window
  int g_cols, g_plane, g_chunks;           // G tile (TH rows)
  int x_bytes, g_bytes, stage_bytes, nstages;
  int tmem_cols;
};

template <int NSPLIT>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_kx_kernel(const KxArgs p)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stages = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)p.nstages * p.stage_bytes);
  uint64_t* full = bars;            // [2]
  uint64_t* empty = bars + 2;       // [2]
  uint64_t* acc_full = bars + 4;
  uint32_t* tmem_base_sh = reinterpret_cast<uint32_t*>(bars + 5);
  float* xparams = reinterpret_cast<float*>(bars + 16);
  float* gparams = xparams + 5 * p.cin_p;
  fillns::stage_params(p.x, xparams, p.cin_p, threadIdx.x, kThreads);
  fillns::stage_params(p.g, gparams, p.cout_p, threadIdx.x, kThreads);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.y, slab = blockIdx.x;
  const int a0 = group * p.acc_per_cta;
  const int a1 = min(p.nacc, a0 + p.acc_per_cta);
  const int ky0 = a0 / p.ncn;                                   // first filter row this CTA touches
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
    // ---------------- MMA issuer (uniform loops, one elected lane issues)
    const uint32_t idesc = tc::idesc_bf16(p.Mrows, p.NC, 1, 1);      // both operands MN-major
    const uint32_t sbase = tc::smem_u32(stages);
    // M operand: 8-channel chunks one plane apart; N operand: "chunks" 16 B apart = the tile shifted by one pixel
    const uint32_t m_plane = p.n_is_g ? p.x_plane : p.g_plane;
    const uint32_t m_lo = p.n_is_g ? x_lo : g_lo, n_lo = p.n_is_g ? g_lo : x_lo;
    const uint32_t n_plane = p.n_is_g ? p.g_plane : p.x_plane;
    const uint64_t mdesc0 = tc::smem_desc_base(128, m_plane), ndesc0 = tc::smem_desc_base(128, 16);
    const uint32_t x_row = (uint32_t)p.x_cols * 16, g_row = (uint32_t)p.g_cols * 16;
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % p.nstages;
      tc::mbar_wait(&full[st], (uint32_t)((it / p.nstages) & 1));
      tc::tc_fence_after();
      const uint32_t xs = sbase + (uint32_t)st * p.stage_bytes;
      const uint32_t gs = xs + (uint32_t)p.x_bytes;
      if (tc::elect_one()) {
        int ky = ky0, c8 = a0 - ky0 * p.ncn;
        uint32_t dacc = tmem_base;
        for (int a = a0; a < a1; ++a) {
          // row r of the tile: X window row (r + ky - ky0); all addresses at column 0 of the K range
          uint32_t xa = xs + (uint32_t)(ky - ky0) * x_row + (p.n_is_g ? 0u : (uint32_t)c8 * n_plane);
          uint32_t ga = gs + (p.n_is_g ? (uint32_t)c8 * n_plane : 0u);
          uint32_t acc = it ? 1u : 0u;
          for (int r = 0; r < p.TH; ++r, xa += x_row, ga += g_row) {
            uint32_t ma = p.n_is_g ? xa : ga, na = p.n_is_g ? ga : xa;
            for (int s = 0; s < p.ksteps; ++s, ma += 256, na += 256) {
              const uint64_t md_hi = tc::smem_desc_at(mdesc0, ma), nd_hi = tc::smem_desc_at(ndesc0, na);
              tc::umma_f16(dacc, md_hi, nd_hi, idesc, acc);
              if (NSPLIT == 3) {
                tc::umma_f16(dacc, tc::smem_desc_at(mdesc0, ma + m_lo), nd_hi, idesc, 1u);
                tc::umma_f16(dacc, md_hi, tc::smem_desc_at(ndesc0, na + n_lo), idesc, 1u);
              }
              acc = 1u;
            }
          }
          dacc += (uint32_t)p.NC;
          if (++c8 == p.ncn) { c8 = 0; ++ky; }
        }
        tc::umma_commit(&empty[st]);
      }
      __syncwarp();
    }
    if (my_tiles > 0 && tc::elect_one()) tc::umma_commit(acc_full);
    __syncwarp();
  } else if (warp >= 2) {
    // ---------------- producers
    const int tid = threadIdx.x - 64;
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % p.nstages;
      if (it >= p.nstages) tc::mbar_wait(&empty[st], (uint32_t)(((it / p.nstages) - 1) & 1));
      int t = slab + it * p.nslabs;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; const int n = t / p.tiles_y;
      const int oy = ty * p.TH, ox = tx * p.TW;
      uint8_t* xs = stages + (size_t)st * p.stage_bytes;
      uint8_t* gs = xs + p.x_bytes;
      // X window: rows oy + ky0 - pad ..., columns from ox - pad (zero outside the image)
      fillns::fill_window<kProducerThreads>(p.x, xs, p.x_plane, x_lo, p.nsplit, n, p.H, p.W, oy + ky0 - p.pad, ox - p.pad, p.x_rows, p.x_cols, 0, p.x_chunks, tid, xparams, p.cin_p);
      // G tile: zero outside this tile's own columns [ox, ox + TW); as the N operand it is stored with k-1 zero
      // columns in front so that N-group j (start + j pixels) is the tile shifted by kx = k-1-j
      fillns::fill_window<kProducerThreads>(p.g, gs, p.g_plane, g_lo, p.nsplit, n, p.H, p.W, oy, p.n_is_g ? ox - (p.k - 1) : ox, p.TH, p.g_cols, 0, p.g_chunks, tid, gparams, p.cout_p,
                          ox, ox + p.TW);
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&full[st]);
    }
    if (my_tiles > 0 && warp < 6) {
      // ---------------- epilogue: RED the partial dW
      tc::mbar_wait(acc_full, 0);
      tc::tc_fence_after();
      const int q = warp & 3;
      const int row = p.Mrows == 128 ? q * 32 + lane : q * 16 + lane;
      const bool row_ok = p.Mrows == 128 || lane < 16;
      const int kk = p.k * p.k;
      int ky = ky0, c8 = a0 - ky0 * p.ncn;
      for (int a = a0; a < a1; ++a) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((a - a0) * p.NC);
        for (int c16 = 0; c16 < p.NC; c16 += 16) {
          float v[16];
          tc::tmem_ld16(taddr + (uint32_t)c16, v);
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int nn = c16 + i, j = nn >> 3, cc = nn & 7;
              if (j < p.k) {
                const int kx = p.n_is_g ? p.k - 1 - j : j;
                const int ci = p.n_is_g ? row : c8 * 8 + cc;
                const int co = p.n_is_g ? c8 * 8 + cc : row;
                if (ci < p.cin_w && co < p.cout_w) atomicAdd(p.dw + ((size_t)co * p.cin_w + ci) * kk + ky * p.k + kx, v[i]);
              }
            }
          }
        }
        if (++c8 == p.ncn) { c8 = 0; ++ky; }
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

// returns 0 = launched, 1 = error, 2 = shape not supported by this kernel (caller falls back to the per-tap kernel)
int cvd_conv_wgrad_kx(const cvd_src_t* gsrc, const cvd_src_t* xsrc, float* dw_oihw,
                      int N, int H, int W, int cin, int cout, int k, int precision, void* stream)
{
  if (k < 3) return 2;
  KxArgs p{};
  p.g = make_view(gsrc, round_up(cout, 4)); p.x = make_view(xsrc, round_up(cin, 4));
  p.dw = dw_oihw; p.N = N; p.H = H; p.W = W; p.k = k; p.pad = (k - 1) / 2;
  p.cin_w = cin; p.cout_w = cout; p.nsplit = precision;
  const int cin_p = round_up(cin, 8), cout_p = round_up(cout, 8);
  p.cin_p = cin_p; p.cout_p = cout_p;
  p.n_is_g = cin_p >= cout_p;
  const int cm = p.n_is_g ? cin_p : cout_p, cn = p.n_is_g ? cout_p : cin_p;
  if (cm > 128) return 2;
  p.Mrows = cm <= 64 ? 64 : 128;
  p.ncn = cn / 8;
  p.NC = p.Mrows == 128 ? round_up(8 * k, 16) : 8 * k;
  p.nacc = k * p.ncn;
  p.acc_per_cta = 512 / p.NC;
  if (p.acc_per_cta > p.nacc) p.acc_per_cta = p.nacc;
  p.ngroups = (p.nacc + p.acc_per_cta - 1) / p.acc_per_cta;
  p.x_chunks = cin_p / 8; p.g_chunks = cout_p / 8;
  const int m_chunks_read = p.Mrows / 8;                       // the M descriptor always walks Mrows/8 planes
  const int max_ky_span = (p.acc_per_cta + p.ncn - 2) / p.ncn + 1;
  const int budget = 218 * 1024;
  bool found = false;
  const int tws[3] = {64, 32, 16};
  const int force_tw = getenv("CVD_KX_TW") ? atoi(getenv("CVD_KX_TW")) : 0;      // tuning overrides
  const int force_th = getenv("CVD_KX_TH") ? atoi(getenv("CVD_KX_TH")) : 0;
  for (int ti = 0; ti < 3 && !found; ++ti) {
    const int TW = tws[ti];
    if (TW > round_up(W, 16) && TW > 16) continue;
    if (force_tw && TW != force_tw) continue;
    for (int th = 16; th >= 1 && !found; th >>= 1) {
      if (force_th && th != force_th) continue;
      const int kcols = p.n_is_g ? round_up(TW + k - 1, 16) : TW;           // K range per row (pixels)
      const int x_cols = p.n_is_g ? kcols : TW + k - 1 + 7;                  // N = X reads up to start + k-1 + 15
      const int g_cols = p.n_is_g ? kcols + k - 1 + 7 : TW;
      const int x_rows = th + (max_ky_span - 1 < k - 1 ? max_ky_span - 1 : k - 1);
      const int xpl = round_up(x_rows * x_cols * 16, 128) + 16, gpl = round_up(th * g_cols * 16, 128) + 16;
      const int xch = p.n_is_g ? (m_chunks_read > p.x_chunks ? m_chunks_read : p.x_chunks) : p.x_chunks;
      const int gch = !p.n_is_g ? (m_chunks_read > p.g_chunks ? m_chunks_read : p.g_chunks) : p.g_chunks;
      // the shifted N views run up to (k-1)*16 + 15*16 bytes past the last K step of a row: keep one spare plane row
      const int xb = xpl * xch * (precision == 3 ? 2 : 1) + 512, gb = gpl * gch * (precision == 3 ? 2 : 1) + 512;
      if (2 * (round_up(xb, 128) + round_up(gb, 128)) + 1024 + fillns::param_bytes(cin_p + cout_p) > budget) continue;
      p.TW = TW; p.TH = th; p.ksteps = kcols / 16;
      p.x_rows = x_rows; p.x_cols = x_cols; p.x_plane = xpl; p.g_cols = g_cols; p.g_plane = gpl;
      p.x_bytes = round_up(xb, 128); p.g_bytes = round_up(gb, 128);
      found = true;
    }
  }
  if (!found) return 2;
  p.stage_bytes = p.x_bytes + p.g_bytes; p.nstages = 2;
  p.tiles_x = (W + p.TW - 1) / p.TW; p.tiles_y = (H + p.TH - 1) / p.TH;
  p.ntiles = N * p.tiles_x * p.tiles_y;
  int slabs = (cvd_num_sms() + p.ngroups - 1) / p.ngroups;
  if (slabs > p.ntiles) slabs = p.ntiles;
  if (slabs < 1) slabs = 1;
  p.nslabs = slabs;
  int cols = p.acc_per_cta * p.NC, pw = 32;
  while (pw < cols) pw <<= 1;
  p.tmem_cols = pw;
  const size_t smem = (size_t)p.nstages * p.stage_bytes + 1024 + fillns::param_bytes(p.cin_p + p.cout_p);
  cudaError_t e = cudaSuccess;
#define CVD_KX_LAUNCH(NS)                                                                                      \
  do {                                                                                                         \
    static bool cfg = false;                                                                                   \
    if (!cfg) {                                                                                                \
      e = cudaFuncSetAttribute(wgrad_kx_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)); \
      cfg = true;                                                                                              \
    }                                                                                                          \
    if (e == cudaSuccess) wgrad_kx_kernel<NS><<<dim3(p.nslabs, p.ngroups), kThreads, smem, (cudaStream_t)stream>>>(p); \
  } while (0)
  if (precision == 3) CVD_KX_LAUNCH(3); else CVD_KX_LAUNCH(1);
  if (e != cudaSuccess) return cvd_fail("cvd_conv_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  CVD_LAUNCH_OK("wgrad_kx_kernel");
  return 0;
}
