// tcgen05 implicit-GEMM convolution (forward and input-gradient) for the depth CNNs.
//
// Replaces nn.Conv2d forward (monodepth/mannequin_challenge/models/hourglass.py:27,39,42,164,173)
// and the dgrad half of its autograd backward (depth_fine_tuning.py:282), with the
// BatchNorm2d(train)+ReLU that PRECEDES the conv (hourglass.py:28-29,40-41,43-44,165-166) applied
// while the activation tile is staged ("normalise on load"), and for dgrad the BatchNorm+ReLU
// BACKWARD of the conv's output applied the same way.
//
// Formulation (stride 1, "same" zero padding):   D[pixel, co] = sum_{tap, ci} A_tap[pixel, ci] * W[co, tap, ci]
//   M = 128 output pixels  (an 8-wide x 16-tall patch: 16 groups of 8 x-adjacent pixels)
//   N = Cout (16..256)     K = taps x Cin, walked in k-blocks of 16 channels of one tap
// HBM layout: NHWC fp32 activations, each conv reads/writes a channel VIEW (offset, gap) of a wider buffer
//   so torch.cat is free.
// SMEM layout of the activation halo tile (the key to the design): bf16, channel-chunk-major
//   [chunk of 8 channels][halo row][halo col][8 ch = 16 B]
// which is exactly the UMMA SWIZZLE_NONE K-major canonical layout: 8 x-adjacent pixels x 16 B form one
// contiguous 128-B core matrix; the next 8-pixel group (next patch row) is SBO = halo_row_pitch away, the
// next 8 channels are LBO = plane_stride away.  A filter tap (ky,kx) is then JUST A DIFFERENT START
// ADDRESS of the A descriptor: the halo tile is staged once and re-used by all k*k taps from shared
// memory; im2col never exists, and HBM/L2 sees each input element ~once per CTA.
// Precision: bf16 tensor cores with the operands split v = hi + lo (3 MMAs: hi*hi, lo*hi, hi*lo,
// fp32 accumulate in TMEM) => fp32-class results ("precision 3"), or plain bf16 ("precision 1").
//
// Persistent, warp-specialised pipeline (one CTA per SM, 480 threads, each CTA loops over output tiles):
//   warps 0-1   MMA issuers (one elected lane each issues tcgen05.mma / commit; warp 0 also owns TMEM)
//   warp 2      weight streamer: cp.async.bulk of pre-packed core-matrix blobs through an mbarrier ring
//   warps 3-10  activation producers: global -> transform -> bf16 hi/lo -> smem slot (ring of 1-2 slots)
//   warps 11-14 epilogue: tcgen05.ld -> +bias / exp -> NHWC global stores
// Two TMEM accumulator buffers and the slot ring let tile t's epilogue, tile t+1's MMAs and tile t+2's
// activation staging run concurrently; for large filters the 64 input channels are staged as two
// 32-channel slots so staging overlaps the MMAs even when one full-depth halo tile is all that fits.
#include "cvd_common.cuh"
#include "tc_common.cuh"
#include "fill.cuh"
#include "bn_epilogue.cuh"

namespace {

constexpr int kIssuers = 2;                // MMA issuer warps (M-tiles split between them)
constexpr int kThreads = 32 * (kIssuers + 1 + 8 + 4);   // issuers | weight streamer | 8 producer warps | 4 epilogue warps
constexpr int kProducerThreads = 256;
constexpr int kMaxStages = 8;
constexpr int kGroupCh = 64;          // channels per activation group resident in one A slot

struct ConvArgs {
  fillns::SrcView src;               // source view + transform
  // weights / bias
  const uint8_t* wp; const float* bias;
  // destination view
  float* y; int y_ct, y_c0, y_n0, y_gap; int cout_valid;
  // problem
  int N, H, W, cin, cout, k, pad;
  int flags, nsplit;                 // nsplit: 1 (bf16) or 3 (bf16x3)
  // tiling
  int mtx, mty, tiles_x, tiles_y;    // M-tiles per CTA along x / y
  int HP, WP, plane_bytes;           // halo dims, bytes of one 8-channel plane (padded)
  int slot_bytes, nslots, ngroups, gchunks;   // items per tile, chunks (8 ch) per item (max)
  int gsplit;                                 // 1: an item = one 64-channel pack group; 2: half of it (32 ch)
  int ntiles;
  int stage_bytes, nstages, kbs, kb_bytes;    // a weight stage = kbs k-blocks of kb_bytes each
  int tmem_cols;
  // fused BatchNorm(train) statistics of the conv output (NULL scratch: off)
  double* st_scratch; const float* st_gamma; const float* st_beta; float* st_rm; float* st_rv;
  float* st_a; float* st_b; float* st_rstd; float* st_mean; float st_eps, st_mom; long long st_count;
  // chunked launch: blockIdx.y = chunk j, an independent conv of the same shape whose source / destination views
  // (with bias and BatchNorm arrays) sit j * ch_src / j * ch_dst channels further, weights j * ch_wp bytes further and
  // statistics scratch j * ch_scratch doubles further (grouped-conv chunks, 256-column chunks of a wide Cout)
  int ch_src, ch_dst, ch_scratch; long long ch_wp;
};

__device__ __forceinline__ int view_phys(int c, int c0, int n0, int gap) { return c0 + c + (c >= n0 ? gap : 0); }

// ------------------------------------------------------------------ the kernel
template <int MT, int NSPLIT>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const ConvArgs p)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  // layout: [A slots][B stages][barriers]
  uint8_t* a_slots = smem;
  uint8_t* b_stages = a_slots + (size_t)p.nslots * p.slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_stages + (size_t)p.nstages * p.stage_bytes);
  uint64_t* b_full = bars;                       // [kMaxStages]
  uint64_t* b_empty = bars + kMaxStages;         // [kMaxStages]
  uint64_t* a_full = bars + 2 * kMaxStages;      // [2]
  uint64_t* a_empty = a_full + 2;                // [2]
  uint64_t* acc_full = a_empty + 2;              // [2]
  uint64_t* acc_empty = acc_full + 2;            // [2]
  uint32_t* tmem_base_sh = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sparams = reinterpret_cast<float*>(bars + 32);           // per-channel constants, 5 x cin floats
  const int chunk = (int)blockIdx.y;
  fillns::SrcView sv = p.src;
  sv.c0 += chunk * p.ch_src; sv.dy_c0 += chunk * p.ch_src;
  const int dshift = chunk * p.ch_dst;
  const uint8_t* const wp = p.wp + (size_t)chunk * p.ch_wp;
  const float* const bias = p.bias ? p.bias + dshift : nullptr;
  const int y_c0 = p.y_c0 + dshift;
  fillns::stage_params(sv, sparams, p.cin, threadIdx.x, kThreads);
  float* sstat = sparams + 5 * p.cin;                              // [4 epilogue warps][2][cout] column sums / sums of squares
  if (p.st_scratch) for (int i = threadIdx.x; i < 8 * p.cout; i += kThreads) sstat[i] = 0.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nstages; ++i) { tc::mbar_init(&b_full[i], 1); tc::mbar_init(&b_empty[i], kIssuers); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&a_full[i], kProducerThreads); tc::mbar_init(&a_empty[i], kIssuers);
      tc::mbar_init(&acc_full[i], kIssuers); tc::mbar_init(&acc_empty[i], 4);
    }
    tc::mbar_fence_init();
  }
  if (warp == 0) {
    tc::tmem_alloc_dyn(tmem_base_sh, (uint32_t)p.tmem_cols);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_sh;

  const int taps = p.k * p.k;

  // item i of a tile -> (pack group g64, sub-group h): channel range and k-block range inside the pack group
  auto item_cfirst = [&](int it) { return p.gsplit == 2 ? (it >> 1) * kGroupCh + (it & 1) * 32 : it * kGroupCh; };
  auto item_chunks = [&](int it) { return p.gsplit == 2 ? 4 : min(p.gchunks, (p.cin - it * kGroupCh) >> 3); };

  if (warp < kIssuers) {
    // ============================ MMA issuers ============================
    // Two issuer warps, each owning half of the CTA's M-tiles (a single thread cannot issue the small-N MMAs
    // fast enough; measured 1.5x on the wgrad kernel).  With one M-tile the second warp only keeps the protocol.
    constexpr int MTW = MT >= kIssuers ? MT / kIssuers : MT;        // M-tiles per issuer warp
    const int mt0 = MT >= kIssuers ? warp * MTW : 0;
    const bool active = MT >= kIssuers || warp == 0;
    const uint32_t idesc = tc::idesc_bf16(128, p.cout, 0, 0);
    const uint32_t a_base = tc::smem_u32(a_slots), b_base = tc::smem_u32(b_stages);
    const uint32_t lo_a = (uint32_t)p.gchunks * p.plane_bytes;      // hi -> lo plane offset inside a slot
    const uint32_t lo_b = (uint32_t)p.cout * 32;                     // hi -> lo blob offset inside a k-block
    const uint64_t adesc0 = tc::smem_desc_base((uint32_t)p.plane_bytes, (uint32_t)p.WP * 16);
    const uint64_t bdesc0 = tc::smem_desc_base(128, 256);
    uint32_t aoff[MTW];
#pragma unroll
    for (int i = 0; i < MTW; ++i) {
      const int mt = mt0 + i;
      const int my = mt / p.mtx, mx = mt - my * p.mtx;
      aoff[i] = (uint32_t)((my * 16 * p.WP + mx * 8) * 16);
    }
    int stage = 0; uint32_t bphase = 0;
    int item = 0, ti = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      if (ti >= 2) { tc::mbar_wait(&acc_empty[buf], (uint32_t)(((ti >> 1) - 1) & 1)); tc::tc_fence_after(); }
      const uint32_t dbase = tmem_base + (uint32_t)(buf * MT * p.cout);
      uint32_t first = 0u;
      for (int g = 0; g < p.ngroups; ++g, ++item) {
        const int slot = item % p.nslots;
        tc::mbar_wait(&a_full[slot], (uint32_t)((item / p.nslots) & 1));
        tc::tc_fence_after();
        const uint32_t slot_addr = a_base + (uint32_t)slot * p.slot_bytes;
        const int nks = (item_chunks(g) >> 1) / p.kbs;              // weight stages per tap for this item
        for (int ky = 0; ky < p.k; ++ky) {
          uint32_t a_row = slot_addr + (uint32_t)(ky * p.WP * 16);
          for (int kx = 0; kx < p.k; ++kx, a_row += 16) {
            for (int ks = 0; ks < nks; ++ks) {
              tc::mbar_wait(&b_full[stage], bphase);
              tc::tc_fence_after();
              if (tc::elect_one()) {
                uint32_t bs = b_base + (uint32_t)stage * p.stage_bytes;
                uint32_t a_kb = a_row + (uint32_t)(2 * ks * p.kbs) * p.plane_bytes;
                for (int j = 0; j < p.kbs && active; ++j, bs += p.kb_bytes, a_kb += 2 * p.plane_bytes) {
                  const uint64_t bd_hi = tc::smem_desc_at(bdesc0, bs), bd_lo = tc::smem_desc_at(bdesc0, bs + lo_b);
#pragma unroll
                  for (int i = 0; i < MTW; ++i) {
                    const uint32_t d = dbase + (uint32_t)((mt0 + i) * p.cout);
                    const uint64_t ad_hi = tc::smem_desc_at(adesc0, a_kb + aoff[i]);
                    tc::umma_f16(d, ad_hi, bd_hi, idesc, first);
                    if (NSPLIT == 3) {
                      tc::umma_f16(d, tc::smem_desc_at(adesc0, a_kb + aoff[i] + lo_a), bd_hi, idesc, 1u);
                      tc::umma_f16(d, ad_hi, bd_lo, idesc, 1u);
                    }
                  }
                  first = 1u;
                }
                tc::umma_commit(&b_empty[stage]);        // weight stage reusable once these MMAs retire
              }
              __syncwarp();
              first = 1u;
              if (++stage == p.nstages) { stage = 0; bphase ^= 1; }
            }
          }
        }
        if (tc::elect_one()) tc::umma_commit(&a_empty[slot]);   // activation slot reusable
        __syncwarp();
      }
      if (tc::elect_one()) tc::umma_commit(&acc_full[buf]);     // this tile's accumulators complete
      __syncwarp();
    }
  } else if (warp == kIssuers) {
    // ============================ weight streamer ============================
    if (lane == 0) {
      int stage = 0; uint32_t ephase = 0; long long issued = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        for (int g = 0; g < p.ngroups; ++g) {
          const int g64 = p.gsplit == 2 ? (g >> 1) : g;
          const int kbg = min(kGroupCh, p.cin - g64 * kGroupCh) >> 4;       // k-blocks of the pack group
          const int kb0 = p.gsplit == 2 ? (g & 1) * 2 : 0;                  // first k-block of this item
          const int nks = (item_chunks(g) >> 1) / p.kbs;
          const uint8_t* gsrc = wp + (size_t)g64 * taps * (kGroupCh / 16) * p.kb_bytes;
          for (int tap = 0; tap < taps; ++tap) {
            for (int ks = 0; ks < nks; ++ks, ++issued) {
              if (issued >= p.nstages) {
                tc::mbar_wait(&b_empty[stage], ephase);
              }
              const uint8_t* src = gsrc + (size_t)(tap * kbg + kb0 + ks * p.kbs) * p.kb_bytes;
              tc::mbar_arrive_expect_tx(&b_full[stage], (uint32_t)p.stage_bytes);
              tc::bulk_g2s(b_stages + (size_t)stage * p.stage_bytes, src, (uint32_t)p.stage_bytes, &b_full[stage]);
              if (++stage == p.nstages) { stage = 0; if (issued >= p.nstages) ephase ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < kIssuers + 9) {
    // ============================ activation producers ============================
    const int tid = threadIdx.x - 32 * (kIssuers + 1);
    int item = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      int t = tile;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; const int n = t / p.tiles_y;
      const int ox = tx * (8 * p.mtx), oy = ty * (16 * p.mty);
      for (int g = 0; g < p.ngroups; ++g, ++item) {
        const int slot = item % p.nslots;
        if (item >= p.nslots) tc::mbar_wait(&a_empty[slot], (uint32_t)(((item / p.nslots) - 1) & 1));
        fillns::fill_window<kProducerThreads>(sv, a_slots + (size_t)slot * p.slot_bytes, p.plane_bytes, p.gchunks * p.plane_bytes, p.nsplit,
                            n, p.H, p.W, oy - p.pad, ox - p.pad, p.HP, p.WP, item_cfirst(g), item_chunks(g), tid, sparams, p.cin);
        tc::fence_proxy_async_smem();
        tc::mbar_arrive(&a_full[slot]);
      }
    }
  } else {
    // ============================ epilogue ============================
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                   // accumulator row = pixel inside the M-tile
    const int py = row >> 3, px = row & 7;
    const bool accum = (p.flags & 1) != 0, do_exp = (p.flags & 2) != 0;
    int ti = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
      int t = tile;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y; const int n = t / p.tiles_y;
      const int ox = tx * (8 * p.mtx), oy = ty * (16 * p.mty);
      const int buf = ti & 1;
      tc::mbar_wait(&acc_full[buf], (uint32_t)((ti >> 1) & 1));
      tc::tc_fence_after();
      for (int mt = 0; mt < MT; ++mt) {
        const int my = mt / p.mtx, mx = mt - my * p.mtx;
        const int yy = oy + my * 16 + py, xx = ox + mx * 8 + px;
        const bool inside = yy < p.H && xx < p.W;
        float* yp = p.y + (((size_t)n * p.H + (inside ? yy : 0)) * p.W + (inside ? xx : 0)) * p.y_ct;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * MT + mt) * p.cout);
        for (int c16 = 0; c16 < p.cout; c16 += 16) {
          float v[16];
          tc::tmem_ld16(taddr + (uint32_t)c16, v);    // warp-collective: executed by all lanes
          if (bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (c16 + i < p.cout_valid) v[i] += __ldg(bias + c16 + i);
          }
          if (p.st_scratch) bnepi::accumulate16(v, inside, lane, sstat + (size_t)q * 2 * p.cout, p.cout, c16);
          if (!inside || c16 >= p.cout_valid) continue;
          if (do_exp) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = expf(v[i]);
          }
          float* dst = yp + view_phys(c16, y_c0, p.y_n0, p.y_gap);
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
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);   // accumulator buffer free for tile ti + 2
    }
    if (p.st_scratch) {
      const bnepi::Stats st{p.st_scratch + (size_t)chunk * p.ch_scratch, p.st_gamma ? p.st_gamma + dshift : nullptr,
                            p.st_beta ? p.st_beta + dshift : nullptr, p.st_rm ? p.st_rm + dshift : nullptr,
                            p.st_rv ? p.st_rv + dshift : nullptr, p.st_a + dshift, p.st_b + dshift, p.st_rstd + dshift,
                            p.st_mean + dshift, p.st_eps, p.st_mom, p.st_count};
      bnepi::finalize(st, sstat, p.cout, p.cout_valid, threadIdx.x - 32 * (kIssuers + 9),
                      reinterpret_cast<volatile int*>(tmem_base_sh + 1));   // (static smem would exceed the 227 KB opt-in)
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------ weight packing
// fp32 OIHW -> per (group, tap, k-block) blobs [hi: cout x 16 ch][lo: cout x 16 ch], each in the
// UMMA SWIZZLE_NONE K-major core-matrix order: blob[n/8][kk/8][n%8][kk%8] (LBO = 128 B, SBO = 256 B).
// transpose_flip builds the dgrad operand: W'[ci][co][k-1-ky][k-1-kx].
__device__ __forceinline__ void pack_one(const float* __restrict__ w, int cin_w, int cout_w, int k, int transpose_flip,
                                         int cin_pad, int cout_pad, int nsplit, uint8_t* __restrict__ out,
                                         long long first, long long step)
{
  // logical GEMM dims: K-channels = cin_pad (multiple of 16), N = cout_pad (multiple of 16)
  // GEMM N above 256 (the tcgen05 / TMEM limit of one launch) is packed as independent 256-column chunks, one
  // after the other: chunk j is exactly the blob of a conv with N = min(256, cout_pad - 256 j).
  // Work item = (column nn, 8 K-channels, tap), nn fastest: a thread writes one 16-B core-matrix row (hi, and lo),
  // a warp 4 x 128 contiguous bytes; the dgrad (flip) reads are fully coalesced, the forward reads sector-exact.
  const int taps = k * k;
  const int ns2 = nsplit == 3 ? 2 : 1;
  const int c8n = cin_pad >> 3;
  const long long total = (long long)cout_pad * c8n * taps;
  const int flip = transpose_flip & 1, gs = transpose_flip >> 8;   // gs > 0: grouped conv, `gs` channels per group
  for (long long i = first; i < total; i += step) {
    const int nn = (int)(i % cout_pad);
    const int c0 = (int)((i / cout_pad) % c8n) * 8;
    const int tap = (int)(i / ((long long)cout_pad * c8n));
    const int ky = tap / k, kx = tap - ky * k;
    const int ftap = (k - 1 - ky) * k + (k - 1 - kx);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      float x = 0.f;
      if (gs == 0) {
        if (!flip) {
          if (c < cin_w && nn < cout_w) x = w[((size_t)nn * cin_w + c) * taps + tap];
        } else {
          // GEMM "cin" = forward Cout (w dim 0), GEMM "cout" = forward Cin (w dim 1)
          if (c < cout_w && nn < cin_w) x = w[((size_t)c * cin_w + nn) * taps + ftap];
        }
      } else if (c < cin_w && nn < cout_w && c / gs == nn / gs) {
        // grouped weights (Cout, gs, k, k) expanded block-diagonally into a dense cin_w x cout_w chunk (cin_w == cout_w)
        if (!flip) x = w[((size_t)nn * gs + c % gs) * taps + tap];
        else       x = w[((size_t)c * gs + nn % gs) * taps + ftap];
      }
      v[j] = x;
    }
    const int nj = nn >> 8, nl = nn & 255;
    const int cp = min(256, cout_pad - (nj << 8));                 // columns of this chunk
    const int stage_bytes = cp * 32 * ns2;
    uint8_t* cout_base = out + (size_t)nj * 256 * cin_pad * taps * 2 * ns2;
    const int g = c0 / kGroupCh, cg = c0 - g * kGroupCh;
    const int gch = min(kGroupCh, cin_pad - g * kGroupCh);
    const int kbg = gch / 16;
    // stage index: groups before g contribute taps * (their kb count) stages
    const int stage = g * taps * (kGroupCh / 16) + tap * kbg + cg / 16;
    const size_t off = (size_t)stage * stage_bytes + (size_t)(nl >> 3) * 256 + (size_t)((cg & 15) >> 3) * 128 + (size_t)(nl & 7) * 16;
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    *reinterpret_cast<uint4*>(cout_base + off) = hi;
    if (nsplit == 3) *reinterpret_cast<uint4*>(cout_base + off + (size_t)cp * 32) = lo;
  }
}

__global__ void pack_weights_kernel(const float* __restrict__ w, int cin_w, int cout_w, int k, int transpose_flip,
                                    int cin_pad, int cout_pad, int nsplit, uint8_t* __restrict__ out)
{
  pack_one(w, cin_w, cout_w, k, transpose_flip, cin_pad, cout_pad, nsplit, out,
           (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

struct PackDesc { const float* w; uint8_t* out; int cin, cout, k, flip; };

// all convolutions of a network in ONE launch (blockIdx.y = conv): replaces ~180 tiny launches per step
__global__ void pack_weights_batch_kernel(const PackDesc* __restrict__ descs, int nsplit)
{
  const PackDesc d = descs[blockIdx.y];
  const int kc = (d.flip & 1) ? d.cout : d.cin, nc = (d.flip & 1) ? d.cin : d.cout;
  pack_one(d.w, d.cin, d.cout, d.k, d.flip, (kc + 15) / 16 * 16, (nc + 15) / 16 * 16, nsplit, d.out,
           (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

extern "C" size_t cvd_conv_packed_bytes(int cin, int cout, int k, int precision)
{
  const int cin_pad = round_up(cin, 16), cout_pad = round_up(cout, 16);
  return (size_t)cin_pad * cout_pad * k * k * 2 * (precision == 3 ? 2 : 1);
}

extern "C" int cvd_conv_pack_weights(const float* w_oihw, int cin, int cout, int k, int transpose_flip,
                                     int precision, void* packed, void* stream)
{
  CVD_CHECK_ARG(w_oihw && packed, "cvd_conv_pack_weights: null pointer");
  CVD_CHECK_ARG(precision == 1 || precision == 3, "cvd_conv_pack_weights: precision must be 1 or 3");
  // forward: GEMM K-channels = cin, N = cout; dgrad: K-channels = cout, N = cin
  CVD_CHECK_ARG((transpose_flip >> 8) == 0 || cin == cout, "cvd_conv_pack_weights: grouped chunks must be square");
  const int kc = (transpose_flip & 1) ? cout : cin, nc = (transpose_flip & 1) ? cin : cout;
  const int cin_pad = round_up(kc, 16), cout_pad = round_up(nc, 16);
  const long long total = (long long)cin_pad * cout_pad * k * k;
  long long blocks = (total / 8 + 255) / 256; if (blocks > 4096) blocks = 4096; if (blocks < 1) blocks = 1;
  pack_weights_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, cin, cout, k, transpose_flip,
                                                                          cin_pad, cout_pad, precision, (uint8_t*)packed);
  CVD_LAUNCH_OK("pack_weights_kernel");
  return 0;
}

extern "C" int cvd_conv_pack_batch(const void* descs_dev, int n, int precision, void* stream)
{
  CVD_CHECK_ARG(descs_dev && n > 0, "cvd_conv_pack_batch: bad arguments");
  CVD_CHECK_ARG(precision == 1 || precision == 3, "cvd_conv_pack_batch: precision must be 1 or 3");
  pack_weights_batch_kernel<<<dim3(32, n), 256, 0, (cudaStream_t)stream>>>((const PackDesc*)descs_dev, precision);
  CVD_LAUNCH_OK("pack_weights_batch_kernel");
  return 0;
}

static int conv_fwd_impl(const cvd_src_t* src, const void* packed_w, const float* bias,
                         const cvd_dst_t* dst, int N, int H, int W, int cin, int cout, int k,
                         int precision, int flags, const cvd_bn_t* bn, void* stream,
                         int nchunks = 1, int ch_src = 0, int ch_dst = 0, long long ch_wp = 0)
{
  if (cout > 256) {
    // more output channels than one launch's TMEM accumulator holds: 256-column chunks (see pack_one) -- one chunked
    // launch (blockIdx.y = chunk) when the chunks are uniform, else one launch per chunk
    CVD_CHECK_ARG(src && dst && packed_w, "cvd_conv_fwd: null pointer");
    CVD_CHECK_ARG(nchunks == 1, "cvd_conv_fwd_chunks: cout must be <= 256");
    if (cout % 256 == 0 && dst->gap == 0 && !getenv("CVD_CONV_NO_CHUNKS"))
      return conv_fwd_impl(src, packed_w, bias, dst, N, H, W, cin, 256, k, precision, flags, bn, stream, cout / 256, 0, 256,
                           (long long)cvd_conv_packed_bytes(cin, 256, k, precision));
    for (int c0 = 0; c0 < cout; c0 += 256) {
      const int cc = cout - c0 < 256 ? cout - c0 : 256;
      cvd_dst_t d = *dst;
      if (d.gap != 0 && c0 >= d.n0) { d.c_off += c0 + d.gap; d.n0 = 0; d.gap = 0; }
      else { d.c_off += c0; if (d.gap != 0) d.n0 -= c0; }
      cvd_bn_t b2;
      if (bn) {
        b2 = *bn;
        if (b2.gamma) b2.gamma += c0;
        if (b2.beta) b2.beta += c0;
        if (b2.running_mean) { b2.running_mean += c0; b2.running_var += c0; }
      }
      const int rc = conv_fwd_impl(src, (const uint8_t*)packed_w + cvd_conv_packed_bytes(cin, c0, k, precision),
                                   bias ? bias + c0 : nullptr, &d, N, H, W, cin, cc, k, precision, flags,
                                   bn ? &b2 : nullptr, stream);
      if (rc) return rc;
    }
    return 0;
  }
  CVD_CHECK_ARG(src && dst && packed_w && src->x && dst->y, "cvd_conv_fwd: null pointer");
  CVD_CHECK_ARG(precision == 1 || precision == 3, "cvd_conv_fwd: precision must be 1 (bf16) or 3 (bf16x3)");
  CVD_CHECK_ARG(k >= 1 && k <= 11 && (k & 1), "cvd_conv_fwd: k=%d unsupported (odd, <= 11)", k);
  CVD_CHECK_ARG(N > 0 && H > 0 && W > 0, "cvd_conv_fwd: bad shape");
  CVD_CHECK_ARG(src->mode == CVD_XF_AFFINE || (src->mode == CVD_XF_BNBWD && src->dy && src->bw && src->a && src->b),
                "cvd_conv_fwd: bad source transform");
  CVD_CHECK_ARG((src->c_total & 3) == 0 && (src->c_off & 3) == 0 && (src->n0 & 7) == 0 && (src->gap & 3) == 0,
                "cvd_conv_fwd: source view must be 4-channel aligned");
  CVD_CHECK_ARG((dst->c_total & 3) == 0 || dst->c_total == 1, "cvd_conv_fwd: destination channel stride must be a multiple of 4 (or 1)");
  ConvArgs p{};
  fillns::SrcView& v = p.src;
  v.x = src->x; v.dy = src->dy; v.a = src->a; v.b = src->b; v.bw = reinterpret_cast<const float4*>(src->bw);
  v.ct = src->c_total; v.c0 = src->c_off; v.n0 = src->n0 > 0 ? src->n0 : (1 << 30); v.gap = src->gap;
  v.dy_ct = src->dy_ctotal; v.dy_c0 = src->dy_coff; v.dy_n0 = src->dy_n0 > 0 ? src->dy_n0 : (1 << 30); v.dy_gap = src->dy_gap;
  v.relu = src->relu; v.mode = src->mode;
  v.cvalid = round_up(cin, 4);
  p.wp = (const uint8_t*)packed_w; p.bias = bias;
  p.y = dst->y; p.y_ct = dst->c_total; p.y_c0 = dst->c_off; p.y_n0 = dst->n0 > 0 ? dst->n0 : (1 << 30); p.y_gap = dst->gap;
  p.cout_valid = cout;
  if (bn) {
    CVD_CHECK_ARG(bn->scratch && bn->a && bn->b && bn->rstd && bn->mean, "cvd_conv_fwd_bn: null pointer");
    CVD_CHECK_ARG(!(flags & 3), "cvd_conv_fwd_bn: statistics cannot be fused with accumulate / exp");
    p.st_scratch = (double*)bn->scratch; p.st_gamma = bn->gamma; p.st_beta = bn->beta; p.st_rm = bn->running_mean; p.st_rv = bn->running_var;
    p.st_a = bn->a + dst->c_off; p.st_b = bn->b + dst->c_off; p.st_rstd = bn->rstd + dst->c_off; p.st_mean = bn->mean + dst->c_off;
    p.st_eps = bn->eps; p.st_mom = bn->momentum; p.st_count = (long long)N * H * W;
  }
  CVD_CHECK_ARG(nchunks >= 1 && nchunks <= 65535, "cvd_conv_fwd_chunks: nchunks=%d", nchunks);
  CVD_CHECK_ARG(nchunks == 1 || (src->gap == 0 && src->dy_gap == 0 && dst->gap == 0 && (ch_src & 3) == 0 && (ch_dst & 3) == 0),
                "cvd_conv_fwd_chunks: chunked launches need gap-free, 4-channel aligned views");
  p.ch_src = ch_src; p.ch_dst = ch_dst; p.ch_wp = ch_wp; p.ch_scratch = 2 * 256 + 1;   // = cvd_bn_scratch_bytes(256) / 8
  p.N = N; p.H = H; p.W = W; p.k = k; p.pad = (k - 1) / 2;
  p.cin = round_up(cin, 16); p.cout = round_up(cout, 16);
  CVD_CHECK_ARG(p.cout <= 256, "cvd_conv_fwd: cout=%d > 256", cout);
  p.flags = flags; p.nsplit = precision;
  const int ng64 = (p.cin + kGroupCh - 1) / kGroupCh;
  p.kb_bytes = p.cout * 32 * (precision == 3 ? 2 : 1);

  // Choose the CTA tile (M-tiles of 8x16 px), the channels per staging slot (64, or 32 = "gsplit 2") and the
  // number of slots.  Two TMEM accumulator buffers: 2 * MT * cout <= 512 columns.  Preference: two slots
  // (staging overlaps the MMAs) with the largest tile that fits; else one slot.
  const int smem_budget = 212 * 1024;
  const int cand[6][2] = {{4, 2}, {4, 1}, {2, 2}, {2, 1}, {1, 2}, {1, 1}};   // (mtx, mty), largest first
  bool found = false;
  // Grid fill: a persistent CTA owns whole tiles, so on small feature maps the largest tile leaves most SMs idle
  // (28x48x8 frames in 32x32-pixel tiles = 16 CTAs).  Prefer the largest tile that still yields >= one tile per SM,
  // then relax the requirement step by step (CVD_TILE_FILL=0 restores "largest tile that fits").
  static const bool fill_grid = !(getenv("CVD_TILE_FILL") && getenv("CVD_TILE_FILL")[0] == '0');
  for (int min_tiles = fill_grid ? cvd_num_sms() : 0; !found; min_tiles = min_tiles > 8 ? min_tiles / 2 : 0) {
  for (int want_slots = 2; want_slots >= 1 && !found; --want_slots) {
    for (int ci = 0; ci < 6 && !found; ++ci) {
      const int mtx = cand[ci][0], mty = cand[ci][1];
      if (2 * mtx * mty * p.cout > 512) continue;
      if ((long long)nchunks * N * ((W + 8 * mtx - 1) / (8 * mtx)) * ((H + 16 * mty - 1) / (16 * mty)) < min_tiles) continue;
      if (8 * mtx > round_up(W, 8) && mtx > 1) continue;
      if (16 * mty > round_up(H, 16) && mty > 1) continue;
      const int HP = 16 * mty + k - 1, WP = 8 * mtx + k - 1;
      for (int gsplit = 1; gsplit <= 2 && !found; ++gsplit) {
        if (gsplit == 2 && (p.cin % kGroupCh != 0)) continue;
        const int gchunks = gsplit == 2 ? 4 : (ng64 == 1 ? p.cin : kGroupCh) / 8;
        const int items = gsplit == 2 ? 2 * ng64 : ng64;
        int plane = HP * WP * 16;
        // producer store bank spreading: plane stride = 16*q (mod 128) with q = pixels per quarter-warp
        const int q = gchunks >= 8 ? 1 : 8 / gchunks;
        plane = (plane + 127) / 128 * 128 + 16 * q;
        const int slot = plane * gchunks * (precision == 3 ? 2 : 1);
        // k-blocks per weight stage: divides every item's k-block count, stage <= 16 KB
        const int last_kb = gsplit == 2 ? 2 : ((p.cin - (ng64 - 1) * kGroupCh) >> 4);
        const int full_kb = gsplit == 2 ? 2 : ((ng64 > 1 ? kGroupCh : p.cin) >> 4);
        int kbs = 4;
        while (kbs > 1 && (last_kb % kbs != 0 || full_kb % kbs != 0 || kbs * p.kb_bytes > 16384)) kbs >>= 1;
        const int stage_bytes = kbs * p.kb_bytes;
        const int stages_per_tile = k * k * (p.cin >> 4) / kbs;
        for (int nst = kMaxStages; nst >= 2 && !found; --nst) {
          if ((size_t)want_slots * slot + (size_t)nst * stage_bytes + 1024 + fillns::param_bytes(p.cin) + 32 * p.cout > (size_t)smem_budget) continue;
          p.mtx = mtx; p.mty = mty; p.nslots = want_slots; p.gsplit = gsplit; p.gchunks = gchunks; p.ngroups = items;
          p.HP = HP; p.WP = WP; p.plane_bytes = plane; p.slot_bytes = slot;
          p.kbs = kbs; p.stage_bytes = stage_bytes;
          p.nstages = nst > stages_per_tile && stages_per_tile >= 1 ? (stages_per_tile < 2 ? 2 : stages_per_tile) : nst;
          if (p.nstages > kMaxStages) p.nstages = kMaxStages;
          found = true;
        }
      }
    }
  }
    if (min_tiles == 0) break;
  }
  CVD_CHECK_ARG(found, "cvd_conv_fwd: no tile fits shared memory (cin=%d cout=%d k=%d)", cin, cout, k);
  p.tiles_x = (W + 8 * p.mtx - 1) / (8 * p.mtx);
  p.tiles_y = (H + 16 * p.mty - 1) / (16 * p.mty);
  p.ntiles = N * p.tiles_x * p.tiles_y;
  int cols = 2 * p.mtx * p.mty * p.cout, pw = 32;
  while (pw < cols) pw <<= 1;
  p.tmem_cols = pw;
  CVD_CHECK_ARG(p.plane_bytes < (1 << 18) && p.WP * 16 < (1 << 18), "cvd_conv_fwd: descriptor offset overflow");

  const size_t smem = (size_t)p.nslots * p.slot_bytes + (size_t)p.nstages * p.stage_bytes + 1024 + fillns::param_bytes(p.cin) + 32 * p.cout;
  // persistent: one CTA per SM (chunked launch: the SMs are divided between the chunks)
  long long grid = (cvd_num_sms() + nchunks - 1) / nchunks;
  if (grid > p.ntiles) grid = p.ntiles;
  const int MT = p.mtx * p.mty;
  cudaError_t e = cudaSuccess;
#define CVD_CONV_LAUNCH(MTV, NS)                                                                             \
  do {                                                                                                       \
    static bool cfg = false;                                                                                 \
    if (!cfg) {                                                                                              \
      e = cudaFuncSetAttribute(conv_tc_kernel<MTV, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)); \
      cfg = true;                                                                                            \
    }                                                                                                        \
    if (e == cudaSuccess) conv_tc_kernel<MTV, NS><<<dim3((unsigned)grid, (unsigned)nchunks), kThreads, smem, (cudaStream_t)stream>>>(p); \
  } while (0)
#define CVD_CONV_MT(NS)                                                                                      \
  do {                                                                                                       \
    if (MT == 8) CVD_CONV_LAUNCH(8, NS); else if (MT == 4) CVD_CONV_LAUNCH(4, NS);                           \
    else if (MT == 2) CVD_CONV_LAUNCH(2, NS); else CVD_CONV_LAUNCH(1, NS);                                   \
  } while (0)
  if (precision == 3) CVD_CONV_MT(3); else CVD_CONV_MT(1);
  if (e != cudaSuccess) return cvd_fail("cvd_conv_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  CVD_LAUNCH_OK("conv_tc_kernel");
  return 0;
}

extern "C" int cvd_conv_fwd(const cvd_src_t* src, const void* packed_w, const float* bias,
                            const cvd_dst_t* dst, int N, int H, int W, int cin, int cout, int k,
                            int precision, int flags, void* stream)
{
  return conv_fwd_impl(src, packed_w, bias, dst, N, H, W, cin, cout, k, precision, flags, nullptr, stream);
}

extern "C" int cvd_conv_fwd_bn(const cvd_src_t* src, const void* packed_w, const float* bias,
                               const cvd_dst_t* dst, int N, int H, int W, int cin, int cout, int k,
                               int precision, int flags, const cvd_bn_t* bn, void* stream)
{
  CVD_CHECK_ARG(bn != nullptr, "cvd_conv_fwd_bn: bn is NULL");
  return conv_fwd_impl(src, packed_w, bias, dst, N, H, W, cin, cout, k, precision, flags, bn, stream);
}

extern "C" int cvd_conv_fwd_chunks(const cvd_src_t* src, const void* packed_w, const float* bias,
                                   const cvd_dst_t* dst, int N, int H, int W, int cin, int cout, int k,
                                   int precision, int flags, const cvd_bn_t* bn,
                                   int nchunks, int src_shift, int dst_shift, long long packed_stride, void* stream)
{
  CVD_CHECK_ARG(cout <= 256, "cvd_conv_fwd_chunks: cout=%d > 256", cout);
  return conv_fwd_impl(src, packed_w, bias, dst, N, H, W, cin, cout, k, precision, flags, bn, stream,
                       nchunks, src_shift, dst_shift, packed_stride);
}
