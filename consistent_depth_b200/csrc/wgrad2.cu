// Second-generation tcgen05 weight-gradient kernel: TMA-fed, kx taps fused into GEMM N, ky taps stacked into GEMM M.
//
//   dW[co][ci][ky][kx] = sum_{n,y,x} G[n,y,x,co] * X[n, y+ky-p, x+kx-p, ci]        (wgrad half of conv backward,
//                                                                                    depth_fine_tuning.py:282)
// Both operands are the pre-split bf16 hi/lo planes written by cvd_prep_operand (prep.cu) -- the SAME planes the
// forward conv (X) and the dgrad conv (G) already consume, so the weight gradient needs no operand pass of its own.
//
// The first-generation kernel (conv_wgrad.cu) issues one M64 x N=16/32 MMA per tap and 16 pixels: half-rate M = 64
// instructions whose 2 KB operand fetch buys 4-8 cycles of tensor work.  Here, with K = 16 pixels of one image row:
//   * N = kx taps x 8 output channels: the G tile is stored with k-1 zero columns on both sides and read MN-major with
//     a 16-byte stride between its 8-channel N-groups, so N-group j IS the tile shifted by j pixels (kx = k-1-j);
//     nothing is copied (the descriptor trick of conv_wgrad_kx.cu).  N = 8(k+1) columns.
//   * M = ky taps x input channels: the X window is staged as [row][chunk][col], which makes "next 8-channel chunk" and
//     "next window row" the same uniform M-group stride, so one M = 128 instruction covers 16/nch consecutive ky rows
//     (2 for 64 input channels, 4 for 32): full-rate M = 128 MMAs.
//   D[(kyl, ci)][(j, co8)] += sum_q X[r + kyM*g + kyl][q][ci] * G[r][q + j - (k-1)][co8]
// Accumulators (one per ky-group x G chunk, 8(k+1) TMEM columns each) are split into passes of <= 512 columns
// (blockIdx.y); a CTA owns a slab of pixel tiles (2-stage TMA ring), then REDs its partial dW.
// The G tile must be zero outside its own columns (each G pixel contributes once): the stage buffers are zeroed once,
// then only tile interiors are (re)written, one small TMA box per (row, chunk, plane); image borders are TMA zero fill.
// 1x1 convolutions: K runs over the flattened image, N-groups are the G channel chunks (no shifts, dense boxes).
#include <cuda.h>
#include "cvd_common.cuh"
#include "tc_common.cuh"
#include <cstdlib>

namespace {

constexpr int kIssuers = 2;                // MMA issuer warps (k > 1: the accumulators of a pass alternate between them)
constexpr int kThreads = 32 * 8;           // warp 0 TMA producer, warps 1-2 MMA issuers (warp 1 owns TMEM), warp 3 idle, warps 4-7 epilogue
constexpr int kMaxStages = 4;
constexpr int kLeft = 16;                    // zero columns in front of a G tile's interior (>= k-1, keeps TMA destinations 128-B aligned)

struct W2Args {
  float* dw; int cin_w, cout_w, dw_ci_stride;
  int N, Hv, Wv;                            // image (k > 1) or flattened image (k == 1: Hv = 1, Wv = H*W)
  int k, pad;
  int nch, kyM, mblk;                       // X chunks per M block (<= 16), ky rows per MMA, M blocks (k == 1, Cin > 128)
  int gch, x_off, g_off;                    // G chunks; first chunk of X / G inside their planes
  int NC;                                   // TMEM columns per accumulator = MMA N
  int nacc, acc_per_pass, npass, nslabs;
  int TH, TW, KQ, GC, XR;                   // tile rows / cols, K range per row (multiple of 16), G row pitch, X window rows
  int tiles_x, tiles_y, ntiles;
  int x_plane_bytes, g_plane_bytes, x_bytes, g_bytes, stage_bytes;   // per stage: [X hi][X lo][G hi][G lo]
  int n_sbo;                                // N-group stride of the G descriptor (16: shifted views; k == 1: chunk plane)
  int tmem_cols;
  int nst;                                  // smem stages of the (X window, G tile) ring
};

__device__ __forceinline__ void tma_load_5d(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t* bar)
{
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(tc::smem_u32(bar)),
        "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad2_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap gmap, const W2Args p)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stages = smem;
  const int kStages = p.nst;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)kStages * p.stage_bytes);
  uint64_t* full = bars;                 // [kMaxStages]
  uint64_t* empty = bars + kMaxStages;   // [kMaxStages]
  uint64_t* acc_full = bars + 2 * kMaxStages;
  uint32_t* tmem_base_sh = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slab = blockIdx.x, pass = blockIdx.y;
  const int a0 = pass * p.acc_per_pass, a1 = min(p.nacc, a0 + p.acc_per_pass);
  // accumulator a <-> (ky group g = a / gch, G chunk c = a % gch) for k > 1; M block a for k == 1
  const int g_lo = p.k > 1 ? a0 / p.gch : 0;
  const int my_tiles = (p.ntiles - slab + p.nslabs - 1) / p.nslabs;

  // zero the G regions once: the halo columns around a tile's interior are never written afterwards
  if (p.k > 1) {
    for (int st = 0; st < kStages; ++st) {
      uint4* g = reinterpret_cast<uint4*>(stages + (size_t)st * p.stage_bytes + p.x_bytes);
      for (int i = threadIdx.x; i < p.g_bytes / 16; i += kThreads) g[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    tc::fence_proxy_async_smem();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], kIssuers); }
    tc::mbar_init(acc_full, kIssuers);
    tc::mbar_fence_init();
  }
  if (warp == 1) { tc::tmem_alloc_dyn(tmem_base_sh, (uint32_t)p.tmem_cols); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_sh;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&xmap)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&gmap)) : "memory");
      for (int it = 0; it < my_tiles; ++it) {
        const int st = it % kStages;
        if (it >= kStages) tc::mbar_wait(&empty[st], (uint32_t)(((it / kStages) - 1) & 1));
        int t = slab + it * p.nslabs;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y; const int n = t / p.tiles_y;
        const int oy = ty * p.TH, ox = tx * p.TW;
        const uint32_t xs = tc::smem_u32(stages + (size_t)st * p.stage_bytes);
        const uint32_t gs = xs + (uint32_t)p.x_bytes;
        if (p.k > 1) {
          const uint32_t g_tx = (uint32_t)(p.TH * p.gch * 2 * p.TW * 16);
          tc::mbar_arrive_expect_tx(&full[st], (uint32_t)p.x_bytes + g_tx);
          // X window: columns ox - pad .. (+KQ), rows oy - pad + kyM*g_lo .. (+XR), all nch chunks, both planes: [plane][row][chunk][col]
          tma_load_5d(xs, &xmap, 2 * (ox - p.pad), p.x_off, oy - p.pad + p.kyM * g_lo, n, 0, &full[st]);
          // G tile interior [kLeft, kLeft+TW) of every (plane, chunk, row): [plane][chunk][row][GC cols]
          for (int pl = 0; pl < 2; ++pl)
            for (int c = 0; c < p.gch; ++c)
              for (int r = 0; r < p.TH; ++r)
                tma_load_5d(gs + (uint32_t)(pl * p.g_plane_bytes + ((c * p.TH + r) * p.GC + kLeft) * 16), &gmap,
                            2 * ox, p.g_off + c, oy + r, n, pl, &full[st]);
        } else {
          tc::mbar_arrive_expect_tx(&full[st], (uint32_t)(p.x_bytes + p.g_bytes));
          tma_load_5d(xs, &xmap, 2 * ox, p.x_off, 0, n, 0, &full[st]);      // [plane][chunk][KQ pixels]
          tma_load_5d(gs, &gmap, 2 * ox, p.g_off, 0, n, 0, &full[st]);
        }
      }
    }
    __syncwarp();
  } else if (warp <= kIssuers) {
    // ============================ MMA issuers ============================
    // k > 1: issuer iw owns the accumulators a0 + iw, a0 + iw + 2, ...; k == 1 (one accumulator per M block, every K step
    // accumulates into it): only issuer 0 issues, the other keeps the barrier protocol
    const int iw = warp - 1;
    const int a_first = p.k > 1 ? a0 + iw : (iw == 0 ? a0 : a1), a_step = p.k > 1 ? kIssuers : 1;
    const uint32_t idesc = tc::idesc_bf16(128, p.NC, 1, 1);                  // both operands MN-major
    const uint32_t sbase = tc::smem_u32(stages);
    const uint64_t mdesc0 = tc::smem_desc_base(128, (uint32_t)p.KQ * 16);    // M groups: next chunk / next window row
    const uint64_t ndesc0 = tc::smem_desc_base(128, (uint32_t)p.n_sbo);
    const uint32_t x_lo = (uint32_t)p.x_plane_bytes, g_lo_off = (uint32_t)p.g_plane_bytes;
    const uint32_t x_rowb = (uint32_t)(p.nch * p.KQ * 16), g_rowb = (uint32_t)(p.GC * 16);
    const int ksteps = p.KQ / 16;
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % kStages;
      tc::mbar_wait(&full[st], (uint32_t)((it / kStages) & 1));
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t xs = sbase + (uint32_t)st * p.stage_bytes;
        const uint32_t gs = xs + (uint32_t)p.x_bytes;
        uint32_t dacc = tmem_base + (uint32_t)((a_first - a0) * p.NC);
        for (int a = a_first; a < a1; a += a_step, dacc += (uint32_t)(a_step * p.NC)) {
          uint32_t xa, ga;
          if (p.k > 1) {
            const int g = a / p.gch, c = a - g * p.gch;
            xa = xs + (uint32_t)(p.kyM * (g - g_lo)) * x_rowb;
            ga = gs + (uint32_t)(c * p.TH) * g_rowb + (uint32_t)((kLeft - (p.k - 1)) * 16);   // column of tap shift j = 0
          } else {
            xa = xs + (uint32_t)(a * 16 * p.KQ * 16);                        // M block a: chunks 16a ..
            ga = gs;
          }
          uint32_t acc = it ? 1u : 0u;
          for (int r = 0; r < p.TH; ++r, xa += x_rowb, ga += g_rowb) {
            uint32_t ma = xa, na = ga;
            for (int s = 0; s < ksteps; ++s, ma += 256, na += 256) {
              const uint64_t md_hi = tc::smem_desc_at(mdesc0, ma), nd_hi = tc::smem_desc_at(ndesc0, na);
              tc::umma_f16(dacc, md_hi, nd_hi, idesc, acc);
              tc::umma_f16(dacc, tc::smem_desc_at(mdesc0, ma + x_lo), nd_hi, idesc, 1u);
              tc::umma_f16(dacc, md_hi, tc::smem_desc_at(ndesc0, na + g_lo_off), idesc, 1u);
              acc = 1u;
            }
          }
        }
        tc::umma_commit(&empty[st]);
      }
      __syncwarp();
    }
    if (my_tiles > 0 && tc::elect_one()) tc::umma_commit(acc_full);
    __syncwarp();
  } else if (my_tiles > 0 && warp >= 4) {
    // ============================ epilogue: RED the partial dW ============================
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
    const int q = warp & 3;
    const int m = q * 32 + lane;                           // accumulator row
    const int kk = p.k * p.k;
    for (int a = a0; a < a1; ++a) {
      int ky = 0, ci, cbase = 0;
      if (p.k > 1) {
        const int g = a / p.gch;
        cbase = (a - g * p.gch) * 8;
        const int kyl = m / (8 * p.nch);
        ci = m - kyl * 8 * p.nch;
        ky = p.kyM * g + kyl;
      } else {
        ci = a * 128 + m;
      }
      const bool row_ok = ky < p.k && ci < p.cin_w;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((a - a0) * p.NC);
      for (int c16 = 0; c16 < p.NC; c16 += 16) {
        float v[16];
        tc::tmem_ld16(taddr + (uint32_t)c16, v);
        if (!row_ok) continue;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int nn = c16 + i;
          int co, kx = 0;
          if (p.k > 1) { const int j = nn >> 3; kx = p.k - 1 - j; co = cbase + (nn & 7); }
          else co = nn;
          if (kx >= 0 && co < p.cout_w) atomicAdd(p.dw + ((size_t)co * p.dw_ci_stride + ci) * kk + ky * p.k + kx, v[i]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// planes z[2][N][zc8][Hv*Wv] x 16 B seen as (x in u64 pairs, chunk, y, image, plane)
int make_map(CUtensorMap* map, const void* z, int zc8, int N, int Hv, int Wv, const cuuint32_t box[5])
{
  EncodeTiledFn enc = encode_fn();
  if (!enc) return -1;
  const cuuint64_t hw = (cuuint64_t)Hv * Wv;
  const cuuint64_t gdim[5] = {(cuuint64_t)(2 * (cuuint64_t)Wv), (cuuint64_t)zc8, (cuuint64_t)Hv, (cuuint64_t)N, 2};
  const cuuint64_t gstr[4] = {hw * 16, (cuuint64_t)Wv * 16, (cuuint64_t)zc8 * hw * 16, (cuuint64_t)N * zc8 * hw * 16};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return (int)enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, const_cast<void*>(z), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace

// Returns 0 = launched, 1 = error, 2 = shape not supported by this kernel (caller uses cvd_conv_wgrad).
// xz / gz: operand planes from cvd_prep_operand of the conv's input X (cin channels from chunk x_off) and of the
// gradient G wrt its raw output (cout channels from chunk g_off); dw_oihw is accumulated into (caller zeroes it).
extern "C" int cvd_conv2_wgrad(const void* xz, int xc8, int x_off, const void* gz, int gc8, int g_off, float* dw_oihw,
                               int N, int H, int W, int cin, int cout, int k, void* stream)
{
  CVD_CHECK_ARG(xz && gz && dw_oihw, "cvd_conv2_wgrad: null pointer");
  CVD_CHECK_ARG(N > 0 && H > 0 && W > 0 && cin > 0 && cout > 0 && k >= 1 && k <= 11 && (k & 1), "cvd_conv2_wgrad: bad shape");
  W2Args p{};
  p.dw = dw_oihw; p.cin_w = cin; p.cout_w = cout; p.dw_ci_stride = cin;
  p.N = N; p.k = k; p.pad = (k - 1) / 2; p.x_off = x_off; p.g_off = g_off;
  const int nch_all = round_up(cin, 16) / 8;
  p.gch = round_up(cout, 16) / 8;
  CVD_CHECK_ARG(x_off >= 0 && x_off + nch_all <= xc8 && g_off >= 0 && g_off + p.gch <= gc8, "cvd_conv2_wgrad: channel range exceeds the operand planes");
  const int budget = 220 * 1024;
  int kStages = 2;
  if (const char* e = getenv("CVD2_WG_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= kMaxStages) kStages = v; }
  p.nst = kStages;
  if (k == 1) {
    if (nch_all % 16 != 0 || round_up(cout, 16) > 256) return 2;             // M blocks of 128 input channels
    p.mblk = nch_all / 16; p.nch = 16; p.kyM = 1;
    p.NC = round_up(cout, 16);
    p.nacc = p.mblk;
    if (p.nacc * p.NC > 512) return 2;
    p.acc_per_pass = p.nacc; p.npass = 1;
    p.Hv = 1; p.Wv = H * W;
    p.TH = 1; p.XR = 1;
    int kp = 128;
    while (kp >= 16) {
      const long long xb = 2ll * nch_all * kp * 16, gb = 2ll * p.gch * kp * 16;
      if (kStages * (xb + gb) + 1024 <= budget) break;
      kp >>= 1;
    }
    if (kp < 16) return 2;
    p.TW = kp; p.KQ = kp; p.GC = kp;
    p.x_plane_bytes = nch_all * kp * 16; p.g_plane_bytes = p.gch * kp * 16;
    p.n_sbo = kp * 16;
  } else {
    if (nch_all != 2 && nch_all != 4 && nch_all != 8 && nch_all != 16) return 2;
    p.nch = nch_all; p.kyM = 16 / nch_all; p.mblk = 1;
    p.NC = 8 * (k + 1);
    const int ngky = (k + p.kyM - 1) / p.kyM;
    p.nacc = ngky * p.gch;
    p.Hv = H; p.Wv = W;
    p.n_sbo = 16;
    // passes of <= 512 TMEM columns: whole ky groups (all G chunks of each) when a group fits, else an even split of ONE
    // group's chunks; then the largest tile (rows x K range) whose two stages fit shared memory
    // ky groups per pass: ONE by default.  More groups per pass mean fewer passes over the pixels but a taller X window per
    // tile (kyM * gpp - 1 extra rows), i.e. smaller tiles in the same shared memory; measured on B200 one group per pass
    // is faster on every hourglass shape (64->16 7x7: 0.68 -> 0.35 ms, 11x11: 0.63 -> 0.57 ms, 32->32 5x5: 0.147 -> 0.112 ms).
    const int gpp_max = 512 / (p.NC * p.gch) < ngky ? 512 / (p.NC * p.gch) : ngky;
    int gpp = gpp_max >= 1 ? 1 : 0;
    if (const char* e = getenv("CVD2_WG_GPP")) { const int v = atoi(e); if (v >= 1 && v <= gpp_max) gpp = v; }
    int app;
    if (gpp >= 1) app = gpp * p.gch;
    else {
      app = 512 / p.NC;
      while (app > 1 && p.gch % app != 0) --app;           // divisor of gch: a pass never straddles two ky groups
      gpp = 1;
    }
    bool found = false;
    const int tw_cand[6] = {118, 112, 96, 80, 64, 48};
    for (; gpp >= 1 && !found; --gpp) {
      if (app >= p.gch) app = gpp * p.gch;
      const int rows_extra = p.kyM * gpp - 1;
      for (int th = 4; th >= 1 && !found; th >>= 1) {
        if (th > H && th > 1) continue;
        for (int ti = 0; ti < 6 && !found; ++ti) {
          int tw = tw_cand[ti];
          if (tw + k - 1 > 128) tw = 128 - (k - 1);
          if (tw > W) tw = W;
          const int kq = round_up(tw + k - 1, 16);
          if (kq > 128) continue;
          const int gc = round_up(kLeft + kq + 1, 8);
          const int xr = th + rows_extra;
          const long long xpl = (long long)xr * p.nch * kq * 16, gpl = (long long)p.gch * th * gc * 16 + 512;
          if (kStages * 2 * (xpl + gpl) + 1024 > budget) continue;
          p.TH = th; p.TW = tw; p.KQ = kq; p.GC = gc; p.XR = xr;
          p.x_plane_bytes = (int)xpl; p.g_plane_bytes = (int)gpl;
          p.acc_per_pass = app;
          found = true;
        }
      }
    }
    if (!found) return 2;
    p.npass = (p.nacc + p.acc_per_pass - 1) / p.acc_per_pass;
  }
  p.x_bytes = 2 * p.x_plane_bytes; p.g_bytes = 2 * p.g_plane_bytes;
  p.stage_bytes = round_up(p.x_bytes + p.g_bytes, 1024);
  p.tiles_x = (p.Wv + p.TW - 1) / p.TW; p.tiles_y = (p.Hv + p.TH - 1) / p.TH;
  p.ntiles = N * p.tiles_x * p.tiles_y;
  int slabs = cvd_num_sms() / p.npass;
  if (slabs < 1) slabs = 1;
  if (slabs > p.ntiles) slabs = p.ntiles;
  p.nslabs = slabs;
  int cols = p.acc_per_pass * p.NC, pw = 32;
  while (pw < cols) pw <<= 1;
  CVD_CHECK_ARG(pw <= 512, "cvd_conv2_wgrad: accumulators exceed TMEM");
  p.tmem_cols = pw;
  CVD_CHECK_ARG(p.KQ * 16 < (1 << 18), "cvd_conv2_wgrad: descriptor stride overflow");

  alignas(64) CUtensorMap xmap, gmap;
  const cuuint32_t xbox[5] = {(cuuint32_t)(2 * p.KQ), (cuuint32_t)(k == 1 ? nch_all : p.nch), (cuuint32_t)p.XR, 1, 2};
  const cuuint32_t gbox1[5] = {(cuuint32_t)(2 * p.KQ), (cuuint32_t)p.gch, 1, 1, 2};       // k == 1: dense [plane][chunk][pixels]
  const cuuint32_t gboxk[5] = {(cuuint32_t)(2 * p.TW), 1, 1, 1, 1};                        // k > 1: one row segment
  int rc = make_map(&xmap, xz, xc8, N, p.Hv, p.Wv, xbox);
  if (rc == 0) rc = make_map(&gmap, gz, gc8, N, p.Hv, p.Wv, k == 1 ? gbox1 : gboxk);
  CVD_CHECK_ARG(rc == 0, "cvd_conv2_wgrad: cuTensorMapEncodeTiled failed (%d) [Wv=%d Hv=%d KQ=%d XR=%d TW=%d]", rc, p.Wv, p.Hv, p.KQ, p.XR, p.TW);

  const size_t smem = (size_t)kStages * p.stage_bytes + 256;
  static bool cfg = false;
  if (!cfg) {
    const cudaError_t e = cudaFuncSetAttribute(wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) return cvd_fail("cvd_conv2_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  wgrad2_kernel<<<dim3(p.nslabs, p.npass), kThreads, smem, (cudaStream_t)stream>>>(xmap, gmap, p);
  CVD_LAUNCH_OK("wgrad2_kernel");
  return 0;
}
