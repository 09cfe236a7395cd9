// Second-generation tcgen05 implicit-GEMM convolution (forward and input-gradient): TMA-fed, kx taps fused into N.
//
// Replaces nn.Conv2d forward (monodepth/mannequin_challenge/models/hourglass.py:27,39,42) and the dgrad half of its
// autograd backward (depth_fine_tuning.py:282) for the stride-1 "same" convolutions of the hourglass.  Differences to
// the first-generation kernel (conv_tc.cu), each answering a measured bottleneck of it:
//
//  * Operands are PRE-SPLIT bf16 hi/lo planes in the chunk-planar layout Z[plane][n][c/8][y][x][8] written once per
//    tensor by cvd_prep_operand (prep.cu).  One 5-D TMA box load (cp.async.bulk.tensor, SASS UTMALDG) drops the
//    (rows x cols x 16 channels x hi/lo) window of a k-block into shared memory in the UMMA SWIZZLE_NONE K-major
//    canonical layout; out-of-image elements are zero-filled by the TMA unit ("same" padding).  No producer warps, no
//    per-consumer fp32->bf16 conversion, no LSU traffic for staging.
//  * The kx filter taps are fused into the GEMM N dimension.  conv_tc issues one M128 x N=Cout MMA per tap; with
//    Cout = 16/32 each fetches a 4 KB A operand for 8-16 cycles of tensor work and the shared-memory operand path
//    (~32 cycles / MMA) is the bound.  Here, for a fixed ky and a group of G horizontal taps kx = k0+j, the weights form
//    ONE B operand [K = 16 channels][N = j*Cout + co] and
//          D[slot s][j*Cout + co] += sum_ci  Z[row + ky][s + k0][ci] * W[co][ci][ky][k0 + j]
//    where an M-tile is 128 consecutive window SLOTS (R rows x WS slots, WS = 64/32/128): the A operand is the window
//    itself (a tap group is a start-address shift of k0 slots), so N = G*Cout = 96..256 columns per MMA.  The epilogue
//    forms   out[x][co] = sum_j D[x + j][j*Cout + co]   through a shared-memory row buffer.  Slots s > WS-k of a window
//    row produce no output (k-1 of WS lanes idle) -- the price of the fusion.
//  * A CTA tile stacks MT M-tiles vertically so that each streamed weight tile (cp.async.bulk through an mbarrier
//    ring, or RESIDENT in shared memory for 1x1 convolutions) feeds MT MMAs x 3 (bf16x3 split).
//
// 1x1 convolutions run the same kernel on the flattened image (H*W pixels as rows of WS slots, k = G = 1).
// Warp roles (224 threads): warp 0 TMA producer (activation windows), warp 1 weight streamer, warp 2 MMA issuer
// (+ TMEM owner), warps 3-6 epilogue (tcgen05.ld -> shifted sum -> bias / BN statistics / exp -> NHWC fp32 stores).
#include <cuda.h>
#include "cvd_common.cuh"
#include "tc_common.cuh"
#include "bn_epilogue.cuh"
#include <cstdlib>

namespace {

constexpr int kEpiGroups = 2;                 // epilogue warp groups (4 warps = the 4 TMEM lane quarters each); M-tiles alternate
constexpr int kIssuers = 2;                   // MMA issuer warps: the M-tiles of a CTA tile alternate between them (one thread
                                               // issues ~1 UTCHMMA per 80 cycles incl. descriptor moves; N = 96 MMAs execute in 48)
constexpr int kEpiWarp0 = 2 + kIssuers;       // first epilogue warp (a multiple of 4: TMEM lane quarter = warp % 4)
constexpr int kThreads = 32 * (kEpiWarp0 + 4 * kEpiGroups);
static_assert(kEpiWarp0 % 4 == 0, "epilogue warps must start at a multiple of 4");
constexpr int kMaxA = 12, kMaxB = 32;        // A stages: 1x1 convs need ~80 KB of activation loads in flight per SM to cover
                                               // the HBM latency (8 KB stages), k x k windows are 40-70 KB each (2 stages)
constexpr int kAPad = 256;                     // zeroed bytes behind every A stage: a zero-weight (padded) tap of the last
                                               // window row reads up to G slots past the stage and must not meet NaN / Inf garbage
constexpr int kOPitch = 20;                    // floats per obuf row (16 + 4: conflict-free float4 access)

struct C2Args {
  const uint8_t* wp; const float* bias;
  float* y; int y_ct, y_c0, y_n0, y_gap, cout_valid;
  int N, Hv, Wv; long long HW;                 // virtual image (flattened for k == 1), pixels per image
  int k, pad, nkb, zc8_off;
  int Cp, G, ng, Ncols;
  int WS, R, MT, TR, WR, VW, tiles_x, tiles_y, ntiles;
  int plane_bytes, a_stage_bytes, NA;
  int b_tile_bytes, NB, b_resident, b_tiles;   // b_tiles = weight tiles per CTA tile (nkb * k * ng)
  int nbuf, tmem_cols;
  int a_split;                                 // 1: the window of a k-block is loaded as 4 boxes (plane x chunk) instead of one
  int flags;
  bnepi::Stats st;                             // st.scratch == nullptr: no fused BatchNorm statistics
};

__device__ __forceinline__ int view_phys(int c, int c0, int n0, int gap) { return c0 + c + (c >= n0 ? gap : 0); }

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t* bar)
{
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(tc::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(tc::smem_u32(bar)),
        "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
conv2_kernel(const __grid_constant__ CUtensorMap zmap, const C2Args p)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_ring = smem;                                                      // NA x (a_stage_bytes + kAPad)
  uint8_t* b_ring = a_ring + (size_t)p.NA * (p.a_stage_bytes + kAPad);          // NB x b_tile_bytes
  float* obuf = reinterpret_cast<float*>(b_ring + (size_t)p.NB * p.b_tile_bytes);   // [groups][128][kOPitch]
  float* sstat = obuf + kEpiGroups * 128 * kOPitch;                            // [4 * groups warps][2][Cp]
  float* sbias = sstat + kEpiGroups * 8 * p.Cp;                                             // [Cp] (zeros when the conv has no bias)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + p.Cp);
  uint64_t* a_full = bars;                      // [kMaxA]
  uint64_t* a_empty = a_full + kMaxA;           // [kMaxA]
  uint64_t* b_full = a_empty + kMaxA;           // [kMaxB]
  uint64_t* b_empty = b_full + kMaxB;           // [kMaxB]
  uint64_t* acc_full = b_empty + kMaxB;         // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_base_sh = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.st.scratch) for (int i = threadIdx.x; i < kEpiGroups * 8 * p.Cp; i += kThreads) sstat[i] = 0.f;
  for (int i = threadIdx.x; i < p.Cp; i += kThreads) sbias[i] = (p.bias && i < p.cout_valid) ? __ldg(p.bias + i) : 0.f;
  for (int i = threadIdx.x; i < p.NA * (kAPad / 16); i += kThreads)
    *reinterpret_cast<uint4*>(a_ring + (size_t)(i / (kAPad / 16)) * (p.a_stage_bytes + kAPad) + p.a_stage_bytes + (i % (kAPad / 16)) * 16) = make_uint4(0u, 0u, 0u, 0u);
  tc::fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.NA; ++i) { tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], kIssuers); }
    for (int i = 0; i < p.NB; ++i) { tc::mbar_init(&b_full[i], 1); tc::mbar_init(&b_empty[i], kIssuers); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], kIssuers); tc::mbar_init(&acc_empty[i], 4 * kEpiGroups); }
    tc::mbar_fence_init();
  }
  if (warp == 2) { tc::tmem_alloc_dyn(tmem_base_sh, (uint32_t)p.tmem_cols); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_sh;

  // Ring positions are RUNNING counters with an explicit wrap / phase bit: the issuer warp is a single thread whose
  // scalar bookkeeping, not the tensor pipe, was the bound of the first version (ncu: ~250 instructions incl. four integer
  // divisions per weight tile, tensor pipe 49 % active with no data waits).
  if (warp == 0) {
    // ============================ TMA producer: one k-block window (hi + lo planes) per stage ============================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&zmap)) : "memory");
      const int NA = p.NA, nkb = p.nkb, a_stride = p.a_stage_bytes + kAPad, tiles_x = p.tiles_x, tiles_y = p.tiles_y;
      int slot = 0; uint32_t eph = 1u;                  // first pass over the ring: nothing to wait for
      bool wrapped = false;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        int t = tile;
        const int tx = t % tiles_x; t /= tiles_x;
        const int ty = t % tiles_y; const int n = t / tiles_y;
        const int iy0 = ty * p.TR - p.pad, ix0 = tx * p.VW - p.pad;
        for (int kb = 0; kb < nkb; ++kb) {
          if (wrapped) tc::mbar_wait(&a_empty[slot], eph);
          tc::mbar_arrive_expect_tx(&a_full[slot], (uint32_t)p.a_stage_bytes);
          // box = (2*WS u64 per row, WR rows, 2 chunks, 1 image, 2 planes): [hi c0][hi c1][lo c0][lo c1], each [row][slot][16 B]
          uint8_t* dst = a_ring + (size_t)slot * a_stride;
          if (!p.a_split) tma_load_5d(dst, &zmap, 2 * ix0, iy0, p.zc8_off + 2 * kb, n, 0, &a_full[slot]);
          else
            for (int q4 = 0; q4 < 4; ++q4)       // [hi c0][hi c1][lo c0][lo c1]: smaller requests interleave with the weight stream
              tma_load_5d(dst + (size_t)q4 * p.plane_bytes, &zmap, 2 * ix0, iy0, p.zc8_off + 2 * kb + (q4 & 1), n, q4 >> 1, &a_full[slot]);
          if (++slot == NA) { slot = 0; if (wrapped) eph ^= 1u; else { wrapped = true; eph = 0u; } }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================ weight streamer ============================
    if (lane == 0) {
      const int b_tiles = p.b_tiles, NB = p.NB; const uint32_t tb = (uint32_t)p.b_tile_bytes;
      if (p.b_resident) {
        for (int t = 0; t < b_tiles; ++t) {
          tc::mbar_arrive_expect_tx(&b_full[t], tb);
          tc::bulk_g2s(b_ring + (size_t)t * tb, p.wp + (size_t)t * tb, tb, &b_full[t]);
        }
      } else {
        int st = 0; uint32_t eph = 0u; bool wrapped = false;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
          const uint8_t* src = p.wp;
          for (int t = 0; t < b_tiles; ++t, src += tb) {
            if (wrapped) tc::mbar_wait(&b_empty[st], eph);
            tc::mbar_arrive_expect_tx(&b_full[st], tb);
            tc::bulk_g2s(b_ring + (size_t)st * tb, src, tb, &b_full[st]);
            if (++st == NB) { st = 0; if (wrapped) eph ^= 1u; else wrapped = true; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < kEpiWarp0) {
    // ============================ MMA issuers (warp 2 also owns TMEM) ============================
    const int iw = warp - 2;                         // this issuer handles the M-tiles mt with mt % kIssuers == iw
    const int NA = p.NA, NB = p.NB, nkb = p.nkb, kk = p.k, ng = p.ng, MT = p.MT, nbuf = p.nbuf, resident = p.b_resident;
    const uint32_t Ncols = (uint32_t)p.Ncols, tb = (uint32_t)p.b_tile_bytes;
    const uint32_t idesc = tc::idesc_bf16(128, p.Ncols, 0, 0);
    const uint32_t a_base = tc::smem_u32(a_ring), b_base = tc::smem_u32(b_ring);
    const uint64_t adesc0 = tc::smem_desc_base((uint32_t)p.plane_bytes, 128);      // LBO: next 8-channel chunk; SBO: next 8 slots
    const uint64_t bdesc0 = tc::smem_desc_base(128, 256);
    const uint32_t lo_a = 2u * (uint32_t)p.plane_bytes, lo_b = 32u * Ncols;
    const uint32_t mt_stride = (uint32_t)(p.R * p.WS * 16), a_stride = (uint32_t)(p.a_stage_bytes + kAPad);
    const uint32_t ky_stride = (uint32_t)(p.WS * 16), g_stride = (uint32_t)(p.G * 16);
    int slot = 0, bst = 0, buf = 0; uint32_t aph = 0u, bph = 0u, cph = 1u; bool cwrapped = false;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      if (cwrapped) { tc::mbar_wait(&acc_empty[buf], cph); tc::tc_fence_after(); }
      const uint32_t dbase = tmem_base + (uint32_t)(buf * MT) * Ncols;
      uint32_t acc = 0u;
      int ridx = 0;                                     // resident weights: tile index inside the CTA tile
      for (int kb = 0; kb < nkb; ++kb) {
        tc::mbar_wait(&a_full[slot], aph);
        tc::tc_fence_after();
        uint32_t a_ky = a_base + (uint32_t)slot * a_stride;
        for (int ky = 0; ky < kk; ++ky, a_ky += ky_stride) {
          uint32_t a_g = a_ky;
          for (int g = 0; g < ng; ++g, a_g += g_stride) {
            int st;
            if (resident) { st = ridx++; tc::mbar_wait(&b_full[st], 0u); }
            else { st = bst; tc::mbar_wait(&b_full[st], bph); }
            tc::tc_fence_after();
            if (tc::elect_one()) {
              const uint32_t bs = b_base + (uint32_t)st * tb;
              const uint64_t bd_hi = tc::smem_desc_at(bdesc0, bs), bd_lo = tc::smem_desc_at(bdesc0, bs + lo_b);
              uint32_t a = a_g + (uint32_t)iw * mt_stride, d = dbase + (uint32_t)iw * Ncols;
              for (int mt = iw; mt < MT; mt += kIssuers, a += kIssuers * mt_stride, d += kIssuers * Ncols) {
                const uint64_t ad_hi = tc::smem_desc_at(adesc0, a);
                tc::umma_f16(d, ad_hi, bd_hi, idesc, acc);
                tc::umma_f16(d, tc::smem_desc_at(adesc0, a + lo_a), bd_hi, idesc, 1u);
                tc::umma_f16(d, ad_hi, bd_lo, idesc, 1u);
              }
              if (!resident) tc::umma_commit(&b_empty[st]);
            }
            __syncwarp();
            acc = 1u;
            if (!resident && ++bst == NB) { bst = 0; bph ^= 1u; }
          }
        }
        if (tc::elect_one()) tc::umma_commit(&a_empty[slot]);
        __syncwarp();
        if (++slot == NA) { slot = 0; aph ^= 1u; }
      }
      if (tc::elect_one()) tc::umma_commit(&acc_full[buf]);
      __syncwarp();
      if (++buf == nbuf) { buf = 0; if (cwrapped) cph ^= 1u; else { cwrapped = true; cph = 0u; } }
    }
  } else {
    // ============================ epilogue ============================
    // Two groups of 4 warps; group g handles the M-tiles with (tile * MT + mt) % 2 == g, so two M-tiles are drained
    // concurrently (one epilogue warp per scheduler is latency-bound: ~130 dependent instructions per 16-column chunk
    // with the BatchNorm statistics).
    const int q = warp & 3;                          // TMEM lane quarter of this warp
    const int grp = (warp - kEpiWarp0) >> 2;
    const int s = q * 32 + lane;                     // window slot of the M-tile held by this thread
    const int r = s / p.WS, sx = s - r * p.WS;
    const bool accum = (p.flags & 1) != 0, do_exp = (p.flags & 2) != 0;
    float* orow = obuf + (grp * 128 + s) * kOPitch;
    float* wstat = sstat + (size_t)(grp * 4 + q) * 2 * p.Cp;
    const int bar_id = 1 + grp;
    int ti = 0, buf = 0; uint32_t fph = 0u;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
      int t = tile;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; const int n = t / p.tiles_y;
      const int oy = ty * p.TR, ox = tx * p.VW;
      tc::mbar_wait(&acc_full[buf], fph);
      tc::tc_fence_after();
      for (int mt = 0; mt < p.MT; ++mt) {
        if (((ti * p.MT + mt) & (kEpiGroups - 1)) != grp) continue;
        const int yy = oy + mt * p.R + r, xx = ox + sx;
        const bool inside = sx < p.VW && yy < p.Hv && xx < p.Wv;
        float* yp = p.y + ((size_t)n * p.HW + (inside ? (size_t)yy * p.Wv + xx : 0)) * p.y_ct;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * p.MT + mt) * p.Ncols);
        for (int c16 = 0; c16 < p.Cp; c16 += 16) {
          float v[16];
          tc::tmem_ld16(taddr + (uint32_t)c16, v);                          // j = 0
          if (p.G > 1) {
            // out[x] = sum_j D[x + j][j]: slot s adds its j-th block into row s - j (rows are distinct per j)
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(orow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            for (int j = 1; j < p.G; ++j) {
              tc::tmem_ld16(taddr + (uint32_t)(j * p.Cp + c16), v);
              if (sx >= j) {
                float* tr = orow - j * kOPitch;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  float4 o = *reinterpret_cast<const float4*>(tr + i);
                  o.x += v[i]; o.y += v[i + 1]; o.z += v[i + 2]; o.w += v[i + 3];
                  *reinterpret_cast<float4*>(tr + i) = o;
                }
              }
              asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            }
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 o = *reinterpret_cast<const float4*>(orow + i);
              v[i] = o.x; v[i + 1] = o.y; v[i + 2] = o.z; v[i + 3] = o.w;
            }
          }
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 bq = *reinterpret_cast<const float4*>(sbias + c16 + i);     // broadcast read
              v[i] += bq.x; v[i + 1] += bq.y; v[i + 2] += bq.z; v[i + 3] += bq.w;
            }
          }
          if (p.st.scratch) bnepi::accumulate16(v, inside, lane, wstat, p.Cp, c16);
          if (!inside || c16 >= p.cout_valid) continue;
          if (do_exp) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = expf(v[i]);
          }
          float* dst = yp + view_phys(c16, p.y_c0, p.y_n0, p.y_gap);
          if (c16 + 16 <= p.cout_valid) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              if (accum) { const float4 old = *reinterpret_cast<const float4*>(dst + i); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
              *reinterpret_cast<float4*>(dst + i) = o;
            }
          } else {
            for (int i = 0; i < p.cout_valid - c16; ++i) dst[i] = accum ? dst[i] + v[i] : v[i];
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);   // accumulator buffer free again
      if (++buf == p.nbuf) { buf = 0; fph ^= 1u; }
    }
    if (p.st.scratch)
      bnepi::finalize<4 * kEpiGroups, 3>(p.st, sstat, p.Cp, p.cout_valid, threadIdx.x - 32 * kEpiWarp0, reinterpret_cast<volatile int*>(tmem_base_sh + 1));
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------ weight packing
// fp32 OIHW -> tiles [(kb, ky, g)] of [hi: N x 16 ch][lo: N x 16 ch], N = G*Cp columns n = j*Cp + co (kx = g*G + j), each
// in the UMMA SWIZZLE_NONE K-major core-matrix order blob[n/8][kk/8][n%8][kk%8] (LBO 128 B, SBO 256 B).
// flip: dgrad operand  W'[ci][co][k-1-ky][k-1-kx]  (GEMM K-channels = forward Cout, N-channels = forward Cin).
struct Pack2Desc { const float* w; uint8_t* out; int cin, cout, k, flip, G, ng; };

__global__ void pack2_kernel(const Pack2Desc* __restrict__ descs)
{
  const Pack2Desc d = descs[blockIdx.y];
  const int kc = d.flip ? d.cout : d.cin, nc = d.flip ? d.cin : d.cout;      // GEMM K-channels, N-channels
  const int kcp = (kc + 15) / 16 * 16, Cp = (nc + 15) / 16 * 16;
  const int Ncols = d.G * Cp;
  const int tile_bytes = 64 * Ncols;
  const long long total = (long long)kcp * d.k * d.ng * Ncols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % kcp);
    long long rest = i / kcp;
    const int nn = (int)(rest % Ncols); rest /= Ncols;
    const int g = (int)(rest % d.ng);
    const int ky = (int)(rest / d.ng);
    const int j = nn / Cp, co = nn - j * Cp;
    const int kx = g * d.G + j;
    float v = 0.f;
    if (kx < d.k && c < kc && co < nc) {
      if (!d.flip) v = d.w[(((size_t)co * d.cin + c) * d.k + ky) * d.k + kx];
      else         v = d.w[(((size_t)c * d.cin + co) * d.k + (d.k - 1 - ky)) * d.k + (d.k - 1 - kx)];
    }
    const int kb = c >> 4, kk = c & 15;
    const size_t tile = ((size_t)kb * d.k + ky) * d.ng + g;
    const size_t off = tile * tile_bytes + (size_t)(nn >> 3) * 256 + (size_t)(kk >> 3) * 128 + (size_t)(nn & 7) * 16 + (size_t)(kk & 7) * 2;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    *reinterpret_cast<__nv_bfloat16*>(d.out + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(d.out + off + 32 * Ncols) = __float2bfloat16_rn(v - __bfloat162float(h));
  }
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

// kx-group size: as many horizontal taps per MMA as fit N <= 256 columns, balanced over the groups
void tap_groups(int k, int Cp, int* G, int* ng)
{
  int gmax = 256 / Cp; if (gmax < 1) gmax = 1;
  if (const char* e = getenv("CVD2_GMAX")) { const int v = atoi(e); if (v >= 1 && v < gmax) gmax = v; }
  if (gmax > k) gmax = k;
  *ng = (k + gmax - 1) / gmax;
  *G = (k + *ng - 1) / *ng;
}

}  // namespace

extern "C" int cvd_conv2_tap_groups(int cout_gemm, int k, int* G, int* ng)
{
  CVD_CHECK_ARG(G && ng && cout_gemm > 0 && k >= 1, "cvd_conv2_tap_groups: bad arguments");
  tap_groups(k, round_up(cout_gemm, 16), G, ng);
  return 0;
}

extern "C" size_t cvd_conv2_packed_bytes(int cin_gemm, int cout_gemm, int k)
{
  int G, ng;
  tap_groups(k, round_up(cout_gemm, 16), &G, &ng);
  return (size_t)(round_up(cin_gemm, 16) / 16) * k * ng * 64 * G * round_up(cout_gemm, 16);
}

// descs_dev: device array of { const float* w_oihw; void* packed; int cin, cout, k, flip, G, ng } (40 bytes each);
// cin / cout are the forward conv's (OIHW extents); flip = 1 packs the dgrad operand.
extern "C" int cvd_conv2_pack_batch(const void* descs_dev, int n, void* stream)
{
  CVD_CHECK_ARG(descs_dev && n > 0, "cvd_conv2_pack_batch: bad arguments");
  pack2_kernel<<<dim3(16, n), 256, 0, (cudaStream_t)stream>>>((const Pack2Desc*)descs_dev);
  CVD_LAUNCH_OK("pack2_kernel");
  return 0;
}

// z: operand planes written by cvd_prep_operand ([2][N][zc8][H*W] x 16 B); the conv reads GEMM-K channels
// [8*zc8_off, 8*zc8_off + ceil16(cin)) of it.  cin / cout in GEMM terms (dgrad: cin = forward Cout).
extern "C" int cvd_conv2_fwd(const void* z, int zc8, int zc8_off, const void* packed_w, const float* bias,
                             const cvd_dst_t* dst, int N, int H, int W, int cin, int cout, int k,
                             int flags, const cvd_bn_t* bn, void* stream)
{
  CVD_CHECK_ARG(z && packed_w && dst && dst->y, "cvd_conv2_fwd: null pointer");
  CVD_CHECK_ARG(k >= 1 && k <= 11 && (k & 1), "cvd_conv2_fwd: k=%d unsupported (odd, <= 11)", k);
  CVD_CHECK_ARG(N > 0 && H > 0 && W > 0 && cin > 0 && cout > 0, "cvd_conv2_fwd: bad shape");
  CVD_CHECK_ARG((dst->c_total & 3) == 0 || dst->c_total == 1, "cvd_conv2_fwd: destination channel stride must be a multiple of 4 (or 1)");
  CVD_CHECK_ARG(((uintptr_t)z & 15) == 0 && ((uintptr_t)packed_w & 15) == 0, "cvd_conv2_fwd: operands must be 16-byte aligned");
  C2Args p{};
  p.wp = (const uint8_t*)packed_w; p.bias = bias;
  p.y = dst->y; p.y_ct = dst->c_total; p.y_c0 = dst->c_off; p.y_n0 = dst->n0 > 0 ? dst->n0 : (1 << 30); p.y_gap = dst->gap;
  p.cout_valid = cout; p.flags = flags;
  p.Cp = round_up(cout, 16); p.nkb = round_up(cin, 16) / 16; p.zc8_off = zc8_off;
  CVD_CHECK_ARG(p.Cp <= 256, "cvd_conv2_fwd: cout=%d > 256", cout);
  CVD_CHECK_ARG(zc8_off >= 0 && zc8_off + 2 * p.nkb <= zc8, "cvd_conv2_fwd: channel range exceeds the operand planes");
  p.N = N; p.k = k; p.pad = (k - 1) / 2; p.HW = (long long)H * W;
  tap_groups(k, p.Cp, &p.G, &p.ng);
  p.Ncols = p.G * p.Cp;
  if (bn) {
    CVD_CHECK_ARG(bn->scratch && bn->a && bn->b && bn->rstd && bn->mean, "cvd_conv2_fwd: bn: null pointer");
    CVD_CHECK_ARG(!(flags & 3), "cvd_conv2_fwd: statistics cannot be fused with accumulate / exp");
    p.st = bnepi::Stats{(double*)bn->scratch, bn->gamma, bn->beta, bn->running_mean, bn->running_var,
                        bn->a + dst->c_off, bn->b + dst->c_off, bn->rstd + dst->c_off, bn->mean + dst->c_off,
                        bn->eps, bn->momentum, (long long)N * H * W};
  }
  // ---- window geometry
  int flat_ws = 0;
  if (k == 1) {
    // flattened image: rows of WS slots (largest WS dividing H*W so that no row straddles two images)
    flat_ws = 128;
    while (flat_ws > 8 && (p.HW % flat_ws) != 0) flat_ws >>= 1;
    if (p.HW % flat_ws != 0) flat_ws = 0;            // tiny odd-sized maps: the general 2-D window path below
  }
  if (flat_ws) {
    p.WS = flat_ws; p.Wv = flat_ws; p.Hv = (int)(p.HW / flat_ws);
  } else {
    p.Wv = W; p.Hv = H;
    p.WS = (W + k - 1 <= 32) ? 32 : 64;
    if (const char* e = getenv("CVD2_WS")) { const int v = atoi(e); if (v == 32 || v == 64 || v == 128) p.WS = v; }
    CVD_CHECK_ARG(p.WS >= k, "cvd_conv2_fwd: window narrower than the filter");
  }
  p.R = 128 / p.WS;
  p.VW = p.WS - (k - 1);
  // ---- CTA tile: MT M-tiles stacked vertically; accumulators MT * Ncols <= 512 TMEM columns
  int mt = 512 / p.Ncols; if (mt > 4) mt = 4;
  // 1x1 convolutions are epilogue-bound (K = 128..256 only): keep two accumulator buffers so the epilogue of a tile
  // overlaps the MMAs of the next (measured: 128->128 at 112x192 0.076 -> 0.059 ms, 224->128 0.095 -> 0.068 ms)
  if (k == 1 && mt > 1) { mt = 256 / p.Ncols; if (mt < 1) mt = 1; }
  if (const char* e = getenv("CVD2_MT")) { const int v = atoi(e); if (v >= 1 && v <= mt) mt = v; }
  const int rows_needed = (p.Hv + p.R - 1) / p.R;
  if (mt > rows_needed) mt = rows_needed;
  // grid fill: prefer at least one tile per SM
  const int tiles_x = (p.Wv + p.VW - 1) / p.VW;
  while (mt > 1 && (long long)N * tiles_x * ((p.Hv + mt * p.R - 1) / (mt * p.R)) < cvd_num_sms()) --mt;
  const int smem_budget = 224 * 1024;
  bool found = false;
  for (; mt >= 1 && !found; --mt) {
    p.MT = mt; p.TR = mt * p.R; p.WR = p.TR + k - 1;
    p.plane_bytes = p.WR * p.WS * 16; p.a_stage_bytes = 4 * p.plane_bytes;
    p.b_tile_bytes = 64 * p.Ncols; p.b_tiles = p.nkb * k * p.ng;
    const size_t fixed = kEpiGroups * (128 * kOPitch * 4 + 8 * p.Cp * 4) + p.Cp * 4 + (2 * kMaxA + 2 * kMaxB + 4) * 8 + 64 + 1024;
    // weights resident in shared memory (1x1 convolutions): loaded once per CTA
    const size_t res_bytes = (size_t)p.b_tiles * p.b_tile_bytes;
    const bool allow_res = !(getenv("CVD2_NO_RESIDENT") && getenv("CVD2_NO_RESIDENT")[0] == '1');
    if (allow_res && p.b_tiles <= kMaxB && fixed + res_bytes + 2 * (size_t)(p.a_stage_bytes + kAPad) <= (size_t)smem_budget) {
      p.b_resident = 1; p.NB = p.b_tiles;
      int na = (int)((smem_budget - fixed - res_bytes) / (p.a_stage_bytes + kAPad));
      p.NA = na > kMaxA ? kMaxA : na;
      found = true; break;
    }
    // streamed weights: enough activation stages for ~96 KB of loads in flight (at least two), then the weight ring
    int na_want = (96 * 1024 + p.a_stage_bytes - 1) / p.a_stage_bytes;
    if (na_want < 2) na_want = 2;
    if (na_want > kMaxA) na_want = kMaxA;
    for (int na = na_want; na >= 1 && !found; --na) {
      const long long left = (long long)smem_budget - (long long)fixed - (long long)na * (p.a_stage_bytes + kAPad);
      int nb = (int)(left / p.b_tile_bytes);
      int nb_cap = 16;
      if (const char* e = getenv("CVD2_NB")) { const int v = atoi(e); if (v >= 2 && v <= kMaxB) nb_cap = v; }
      if (nb > nb_cap) nb = nb_cap;
      if (nb >= 2) { p.b_resident = 0; p.NA = na; p.NB = nb; found = true; }
    }
    if (found) break;
  }
  CVD_CHECK_ARG(found, "cvd_conv2_fwd: no tile fits shared memory (cin=%d cout=%d k=%d)", cin, cout, k);
  p.tiles_x = tiles_x; p.tiles_y = (p.Hv + p.TR - 1) / p.TR;
  p.ntiles = N * p.tiles_x * p.tiles_y;
  p.nbuf = (2 * p.MT * p.Ncols <= 512) ? 2 : 1;
  int cols = p.nbuf * p.MT * p.Ncols, pw = 32;
  while (pw < cols) pw <<= 1;
  p.tmem_cols = pw;
  CVD_CHECK_ARG(p.plane_bytes < (1 << 18), "cvd_conv2_fwd: descriptor offset overflow");

  // ---- tensor map over the operand planes: (x as u64 pairs, y, chunk, image, plane)
  EncodeTiledFn enc = encode_fn();
  CVD_CHECK_ARG(enc != nullptr, "cvd_conv2_fwd: cuTensorMapEncodeTiled not available from the driver");
  alignas(64) CUtensorMap map;
  const cuuint64_t gdim[5] = {(cuuint64_t)(2 * p.Wv), (cuuint64_t)p.Hv, (cuuint64_t)zc8, (cuuint64_t)N, 2};
  const cuuint64_t gstr[4] = {(cuuint64_t)p.Wv * 16, (cuuint64_t)p.HW * 16, (cuuint64_t)zc8 * p.HW * 16, (cuuint64_t)N * zc8 * p.HW * 16};
  p.a_split = (getenv("CVD2_ASPLIT") && getenv("CVD2_ASPLIT")[0] == '1') ? 1 : 0;
  const cuuint32_t box[5] = {(cuuint32_t)(2 * p.WS), (cuuint32_t)p.WR, p.a_split ? 1u : 2u, 1, p.a_split ? 1u : 2u};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, const_cast<void*>(z), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CVD_CHECK_ARG(cr == CUDA_SUCCESS, "cvd_conv2_fwd: cuTensorMapEncodeTiled failed (%d) [Wv=%d Hv=%d zc8=%d N=%d WS=%d WR=%d]",
                (int)cr, p.Wv, p.Hv, zc8, N, p.WS, p.WR);

  const size_t smem = (size_t)p.NA * (p.a_stage_bytes + kAPad) + (size_t)p.NB * p.b_tile_bytes + kEpiGroups * (128 * kOPitch * 4 + 8 * p.Cp * 4) + p.Cp * 4 +
                      (2 * kMaxA + 2 * kMaxB + 4) * 8 + 64;
  const long long grid = p.ntiles < cvd_num_sms() ? p.ntiles : cvd_num_sms();
  static bool cfg = false;
  if (!cfg) {
    const cudaError_t e = cudaFuncSetAttribute(conv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) return cvd_fail("cvd_conv2_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  conv2_kernel<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(map, p);
  CVD_LAUNCH_OK("conv2_kernel");
  return 0;
}
