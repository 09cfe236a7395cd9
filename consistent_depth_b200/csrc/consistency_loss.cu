// Fused geometric-consistency loss, forward + backward, one HBM-bound kernel.
//
// Replaces (reference file:line, /root/reference):
//   loss/consistency_loss.py:210-253  ConsistencyLoss.__call__
//   loss/consistency_loss.py:98-208   geometry_consistency_loss
//   loss/consistency_loss.py:73-89    weighted_mean_loss
//   utils/geometry.py:9-19,38-61,86-100  pixel_grid / pixels_to_rays / pixels_to_points
//   utils/geometry.py:103-128         reproject_points (baddbmm + bmm, K=3)
//   utils/geometry.py:64-83           project
//   utils/geometry.py:201-208         sample (grid_sample bilinear, border, align_corners=False)
// and the autograd backward of that whole sub-graph (depth_fine_tuning.py:282).
//
// Algorithmic traffic (fp32): per pixel per frame pair, read depth 2x4 + flow 2x8 +
// mask 2x4 = 32 B, write grad 2x4 = 8 B  => 40 B/px/pair (fwd only: 32 B).
// Everything else (pixel grid, rays, points, reprojected points, projected
// pixels, sampled z, per-pixel distances) lives in registers.
//
// Work decomposition: grid = (pixels / 1024, B).  A block owns 1024 consecutive
// pixels; a thread owns 4 of them strided by the block size (so a warp's lanes
// are x-adjacent pixels: coalesced 128 B loads AND spatially coherent scatter
// REDs that merge into few L2 sector requests) and handles BOTH directions, so
// 32 independent loads are in flight per thread before any math.  Per
// pixel-direction the math is ~80 instructions (rcp/rsqrt via MUFU approx).
// The bilinear gather reads the other frame's depth plane through the
// read-only path (it is L1/L2 resident: flow is spatially coherent); the
// matching backward scatter uses fire-and-forget fp32 REDs that resolve in L2.
#include "cvd_common.cuh"
#include <cstdlib>

namespace {

// Per pair-direction constants, computed once per block (in double) and broadcast from smem.
// Q = d * (M ray) + c with M = R_t^T R_r, c = R_t^T (t_r - t_t): the reference's
// baddbmm+bmm chain (geometry.py:121-127) collapsed to 9 FMAs per pixel.
struct DirConst {
  float M[9], c[3];
  float ifx, ify, cx, cy;     // reference-frame intrinsics (1/fx, 1/fy)
  float fxt, fyt, cxt, cyt;   // target-frame intrinsics
  float cr, cd;               // gradient coefficients incl. 1/(2 B_global) and 1/mask_sum
};

struct PairConst { DirConst d[2]; };

__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

#ifndef CVD_LOSS_PIX
#define CVD_LOSS_PIX 4
#endif
#ifndef CVD_LOSS_MINB
#define CVD_LOSS_MINB 4
#endif
constexpr int PIX = CVD_LOSS_PIX;   // pixels per thread, strided by the block size (lane <-> adjacent pixels)
constexpr int LOSS_THREADS = 256;

// fire-and-forget vector REDs (sm_90+: REDG.ADD.F32x2 / F32x4): one L1TEX/L2 request for two x-adjacent bilinear taps
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// the two x-adjacent taps (i, i+1) of one row: aligned pair -> v2; pair inside an aligned float4 -> v4 (zeros elsewhere);
// straddling a 16-byte boundary or clamped at the image border -> scalars
__device__ __forceinline__ void scatter_row(float* __restrict__ g, int i0, int i1, float v0, float v1) {
  if (v0 == 0.f && v1 == 0.f) return;                  // masked-out pixel / zero-weight border taps
  if (i1 == i0 + 1) {
    const int a = i0 & 3;
    if (!(a & 1)) { red_add_v2(g + i0, v0, v1); return; }
    if (a == 1) { red_add_v4(g + i0 - 1, 0.f, v0, v1, 0.f); return; }
  }
  if (v0 != 0.f) atomicAdd(g + i0, v0);
  if (v1 != 0.f) atomicAdd(g + i1, v1);
}

template <bool want_grad, bool do_r, bool do_d>
__device__ __forceinline__ float pixel_term(
    const DirConst& c, const float* __restrict__ depth_t, float* __restrict__ grad_t, int plane_off,
    float x, float y, float d, float fu, float fv, float mk,
    int W, int Wm1i, int Hm1i, float sxs, float sys,
    float& acc_r, float& acc_d)
{
  // pixels_to_rays (geometry.py:38-61): ((x-cx)/fx, -(y-cy)/fy, -1) ; m = M ray
  const float rx = (x - c.cx) * c.ifx;
  const float ry = (c.cy - y) * c.ify;
  const float mx = fmaf(c.M[0], rx, fmaf(c.M[1], ry, -c.M[2]));
  const float my = fmaf(c.M[3], rx, fmaf(c.M[4], ry, -c.M[5]));
  const float mz = fmaf(c.M[6], rx, fmaf(c.M[7], ry, -c.M[8]));
  // pixels_to_points + reproject_points (:86-128)
  const float Qx = fmaf(d, mx, c.c[0]), Qy = fmaf(d, my, c.c[1]), Qz = fmaf(d, mz, c.c[2]);
  const float rq = rcp_fast(Qz);                       // 1/Qz
  const float a = Qx * rq, b = Qy * rq;
  const float tx = x + fu, ty = y + fv;                // matched_pixels_tgt (:165)
  float g = 0.f;

  if (do_r) {
    // project (:64-83): u = -fx' Qx/Qz + cx' ; v = +fy' Qy/Qz + cy'
    const float ex = fmaf(-c.fxt, a, c.cxt) - tx;
    const float ey = fmaf(c.fyt, b, c.cyt) - ty;
    const float d2 = fmaf(ex, ex, ey * ey);
    const float rs = rsqrt_fast(d2);
    const float dist = d2 > 0.f ? d2 * rs : d2;       // torch.norm (:170); d2 itself carries NaN/0
    acc_r = fmaf(mk, dist, acc_r);
    if (want_grad) {
      const float du = -c.fxt * (mx - a * mz);         // * rq below
      const float dv = c.fyt * (my - b * mz);
      const float t = (ex * du + ey * dv) * rq * rs;   // d dist / d d
      g = d2 > 0.f ? c.cr * mk * t : 0.f;              // norm subgradient 0 at 0
    }
  }
  if (do_d) {
    // sample (:201-208): grid_sample(bilinear, border, align_corners=False) at
    // ix = tx W/(W-1) - 0.5 (the (2uv/(W-1)-1 -> ((g+1)W-1)/2) chain folded to one FMA)
    float ix = fmaf(tx, sxs, -0.5f), iy = fmaf(ty, sys, -0.5f);
    ix = fminf(fmaxf(ix, 0.f), (float)Wm1i);
    iy = fminf(fmaxf(iy, 0.f), (float)Hm1i);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float wx1 = ix - x0f, wy1 = iy - y0f;
    const float wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
    const int x0 = (int)x0f, y0 = (int)y0f;
    // border taps: the +1 neighbour is clamped; its weight is exactly 0 there
    const int i00 = y0 * W + x0 + plane_off;     // 32-bit element index into the whole tensor
    const int i10 = i00 + (x0 < Wm1i ? 1 : 0);
    const int i01 = i00 + (y0 < Hm1i ? W : 0);
    const int i11 = i01 + (i10 - i00);
    const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
    const float d00 = __ldg(depth_t + i00), d10 = __ldg(depth_t + i10);
    const float d01 = __ldg(depth_t + i01), d11 = __ldg(depth_t + i11);
    const float zs = fmaf(w00, d00, fmaf(w10, d10, fmaf(w01, d01, w11 * d11)));   // = -z_w
    const float rz = rcp_fast(zs);                     // -1/z_w
    const float s = rq + rz;                           // 1/Qz - 1/z_w (:188-189)
    acc_d = fmaf(mk, fabsf(s), acc_d);
    if (want_grad) {
      const float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
      const float k0 = c.cd * mk * sg;
      g = fmaf(k0, -mz * rq * rq, g);
      const float k1 = -k0 * rz * rz;                  // d|s|/d depth_tgt[tap] = -sg wt / z_w^2
      const float v00 = k1 * w00, v10 = k1 * w10, v01 = k1 * w01, v11 = k1 * w11;
      scatter_row(grad_t, i00, i10, v00, v10);          // vector REDs where the tap pair is 8- / 16-byte aligned
      scatter_row(grad_t, i01, i11, v01, v11);
    }
  }
  return g;
}

// One thread per (pair, direction): collapse the poses to M = R_t^T R_r, c = R_t^T (t_r - t_t) in
// double, fold 1/(2 B_global), 1/max(sum(mask),1e-6) and the batch-mean focal length into the
// gradient coefficients.  Runs once per call (B*2 threads), so the main kernel has no serial prologue.
__global__ void consistency_setup(const float* __restrict__ extr, const float* __restrict__ intr,
                                  const float* __restrict__ msum, float f0, float f1, int f_given,
                                  const float* __restrict__ f_dev,
                                  float lam_r, float lam_b, int B, int B_global, PairConst* __restrict__ consts)
{
  if (f_dev) { f0 = f_dev[0]; f1 = f_dev[1]; f_given = 1; }   // device-resident global focal length (graph-safe)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * B) return;
  const int b = i >> 1, k = i & 1, t = 1 - k;
  DirConst c;
  const float* Er = extr + ((size_t)b * 2 + k) * 12;
  const float* Et = extr + ((size_t)b * 2 + t) * 12;
  double dt[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) dt[a] = (double)Er[a * 4 + 3] - (double)Et[a * 4 + 3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {          // row a of R_t^T = column a of R_t
    double ca = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double m = 0.0;
#pragma unroll
      for (int l = 0; l < 3; ++l) m += (double)Et[l * 4 + a] * (double)Er[l * 4 + j];
      c.M[a * 3 + j] = (float)m;
      ca += (double)Et[j * 4 + a] * dt[j];
    }
    c.c[a] = (float)ca;
  }
  const float* Ir = intr + ((size_t)b * 2 + k) * 4;
  const float* It = intr + ((size_t)b * 2 + t) * 4;
  c.ifx = 1.0f / Ir[0]; c.ify = 1.0f / Ir[1]; c.cx = Ir[2]; c.cy = Ir[3];
  c.fxt = It[0]; c.fyt = It[1]; c.cxt = It[2]; c.cyt = It[3];
  float f = k ? f1 : f0;
  if (!f_given) {          // f = mean(focal_length(intrinsics_ref)) over the batch (:178)
    float sacc = 0.f;
    for (int bb = 0; bb < B; ++bb) { const float* I = intr + ((size_t)bb * 2 + k) * 4; sacc += I[0]; sacc += I[1]; }
    f = sacc / (float)(2 * B);
  }
  const float inv = 1.0f / fmaxf(msum[b * 2 + k], 1e-6f);   // weighted_mean_loss eps (:73)
  const float half_over_B = 0.5f / (float)B_global;
  c.cr = lam_r * half_over_B * inv;
  c.cd = lam_b * f * half_over_B * inv;
  consts[b].d[k] = c;
}

template <bool GRAD, bool DO_R, bool DO_D, bool FULL>
__global__ void __launch_bounds__(LOSS_THREADS, CVD_LOSS_MINB)
consistency_kernel(const float* __restrict__ depth,
                   const float* __restrict__ flow0, const float* __restrict__ flow1,
                   const float* __restrict__ mask0, const float* __restrict__ mask1,
                   const PairConst* __restrict__ consts, int H, int W,
                   double* __restrict__ acc, float* __restrict__ grad)
{
  __shared__ PairConst pc;
  __shared__ float red[LOSS_THREADS / 32][4];
  const int b = blockIdx.y;
  const int HW = H * W;

  {   // per-pair constants were prepared by consistency_setup; stage them in smem (broadcast reads)
    const float* src = reinterpret_cast<const float*>(consts + b);
    float* dst = reinterpret_cast<float*>(&pc);
    for (int i = threadIdx.x; i < (int)(sizeof(PairConst) / 4); i += LOSS_THREADS) dst[i] = __ldg(src + i);
  }
  __syncthreads();

  const float sxs = (float)W / (float)(W - 1), sys = (float)H / (float)(H - 1);
  const float invW = 1.0f / (float)W;
  float a_r[2] = {0.f, 0.f}, a_d[2] = {0.f, 0.f};

  const int base = blockIdx.x * (LOSS_THREADS * PIX) + threadIdx.x;
  // per-array pointers for this thread's first pixel; the other PIX-1 are at constant strides
  const float* pd0 = depth + (size_t)b * 2 * HW + base;
  const float* pd1 = pd0 + HW;
  const float* pf0u = flow0 + (size_t)b * 2 * HW + base;
  const float* pf0v = pf0u + HW;
  const float* pf1u = flow1 + (size_t)b * 2 * HW + base;
  const float* pf1v = pf1u + HW;
  const float* pm0 = mask0 + (size_t)b * HW + base;
  const float* pm1 = mask1 + (size_t)b * HW + base;
  float dv[2][PIX], fu[2][PIX], fv[2][PIX], mk[2][PIX];
#pragma unroll
  for (int j = 0; j < PIX; ++j) {                      // 32 independent coalesced loads in flight
    const int o = j * LOSS_THREADS;
    const bool ok = FULL || (base + o < HW);
    dv[0][j] = ok ? __ldcs(pd0 + o) : 1.f;   dv[1][j] = ok ? __ldcs(pd1 + o) : 1.f;
    fu[0][j] = ok ? __ldcs(pf0u + o) : 0.f;  fv[0][j] = ok ? __ldcs(pf0v + o) : 0.f;
    fu[1][j] = ok ? __ldcs(pf1u + o) : 0.f;  fv[1][j] = ok ? __ldcs(pf1v + o) : 0.f;
    mk[0][j] = ok ? __ldcs(pm0 + o) : 0.f;   mk[1][j] = ok ? __ldcs(pm1 + o) : 0.f;
  }
  float* gpl = GRAD ? grad + (size_t)b * 2 * HW : nullptr;
  float* gd0 = GRAD ? gpl + base : nullptr;            // direct-term targets, constant strides per j
  float* gd1 = GRAD ? gd0 + HW : nullptr;
#pragma unroll
  for (int j = 0; j < PIX; ++j) {
    const int p = base + j * LOSS_THREADS;
    if (FULL || p < HW) {
      // (y, x) of the flattened index without an integer divide (exact for HW < 2^22, checked on host)
      int y = __float2int_rd(((float)p + 0.5f) * invW);
      int x = p - y * W;
      if (x < 0) { x += W; --y; } else if (x >= W) { x -= W; ++y; }
      const float xf = (float)x, yf = (float)y;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float g = pixel_term<GRAD, DO_R, DO_D>(pc.d[k], depth, grad, (b * 2 + 1 - k) * HW, xf, yf,
                                                     dv[k][j], fu[k][j], fv[k][j], mk[k][j],
                                                     W, W - 1, H - 1, sxs, sys, a_r[k], a_d[k]);
        if (GRAD) atomicAdd((k ? gd1 : gd0) + j * LOSS_THREADS, g);
      }
    }
  }

  // block reduction of the four masked sums -> f64 accumulators acc[b][k][{r,d}]
  float v0 = warp_sum(a_r[0]), v1 = warp_sum(a_d[0]), v2 = warp_sum(a_r[1]), v3 = warp_sum(a_d[1]);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[wid][0] = v0; red[wid][1] = v1; red[wid][2] = v2; red[wid][3] = v3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < LOSS_THREADS / 32; ++w) s += (double)red[w][threadIdx.x];
    atomicAdd(acc + (size_t)b * 4 + threadIdx.x, s);
  }
}

// Variant for W % 4 == 0 (every configuration of BASELINE.json): a thread owns 4 x-ADJACENT pixels, so its eight input
// streams are 128-bit loads (8 LDG.128 instead of 32 LDG.32) and the direct gradient term is ONE vector RED per direction
// instead of four scalar ones; the bilinear scatter uses the pair REDs above.  L1TEX requests per pixel pair drop from
// ~26 to ~15 (the kernel is bound by L1TEX / RED throughput, profiles/r01_ncu_full_loss_v4_grad.csv).
template <bool GRAD, bool DO_R, bool DO_D>
__global__ void __launch_bounds__(LOSS_THREADS, 3)
consistency_kernel_x4(const float* __restrict__ depth,
                      const float* __restrict__ flow0, const float* __restrict__ flow1,
                      const float* __restrict__ mask0, const float* __restrict__ mask1,
                      const PairConst* __restrict__ consts, int H, int W,
                      double* __restrict__ acc, float* __restrict__ grad)
{
  __shared__ PairConst pc;
  __shared__ float red[LOSS_THREADS / 32][4];
  const int b = blockIdx.y;
  const int HW = H * W;
  {
    const float* src = reinterpret_cast<const float*>(consts + b);
    float* dst = reinterpret_cast<float*>(&pc);
    for (int i = threadIdx.x; i < (int)(sizeof(PairConst) / 4); i += LOSS_THREADS) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const float sxs = (float)W / (float)(W - 1), sys = (float)H / (float)(H - 1);
  float a_r[2] = {0.f, 0.f}, a_d[2] = {0.f, 0.f};
  const int p0 = (blockIdx.x * LOSS_THREADS + threadIdx.x) * 4;           // first of this thread's 4 pixels (same row: W % 4 == 0)
  if (p0 < HW) {
    const size_t o2 = (size_t)b * 2 * HW + p0, o1 = (size_t)b * HW + p0;
    float4 dq[2], fuq[2], fvq[2], mq[2];
    dq[0] = __ldcs(reinterpret_cast<const float4*>(depth + o2));   dq[1] = __ldcs(reinterpret_cast<const float4*>(depth + o2 + HW));
    fuq[0] = __ldcs(reinterpret_cast<const float4*>(flow0 + o2));  fvq[0] = __ldcs(reinterpret_cast<const float4*>(flow0 + o2 + HW));
    fuq[1] = __ldcs(reinterpret_cast<const float4*>(flow1 + o2));  fvq[1] = __ldcs(reinterpret_cast<const float4*>(flow1 + o2 + HW));
    mq[0] = __ldcs(reinterpret_cast<const float4*>(mask0 + o1));   mq[1] = __ldcs(reinterpret_cast<const float4*>(mask1 + o1));
    const int y = p0 / W, x0 = p0 - y * W;
    const float yf = (float)y;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float dv[4] = {dq[k].x, dq[k].y, dq[k].z, dq[k].w}, fu[4] = {fuq[k].x, fuq[k].y, fuq[k].z, fuq[k].w};
      const float fv[4] = {fvq[k].x, fvq[k].y, fvq[k].z, fvq[k].w}, mk[4] = {mq[k].x, mq[k].y, mq[k].z, mq[k].w};
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        g[j] = pixel_term<GRAD, DO_R, DO_D>(pc.d[k], depth, grad, (b * 2 + 1 - k) * HW, (float)(x0 + j), yf,
                                            dv[j], fu[j], fv[j], mk[j], W, W - 1, H - 1, sxs, sys, a_r[k], a_d[k]);
      if (GRAD && (g[0] != 0.f || g[1] != 0.f || g[2] != 0.f || g[3] != 0.f))
        red_add_v4(grad + (size_t)(b * 2 + k) * HW + p0, g[0], g[1], g[2], g[3]);
    }
  }
  float v0 = warp_sum(a_r[0]), v1 = warp_sum(a_d[0]), v2 = warp_sum(a_r[1]), v3 = warp_sum(a_d[1]);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[wid][0] = v0; red[wid][1] = v1; red[wid][2] = v2; red[wid][3] = v3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < LOSS_THREADS / 32; ++w) s += (double)red[w][threadIdx.x];
    atomicAdd(acc + (size_t)b * 4 + threadIdx.x, s);
  }
}

// acc[b][k][{r,d}] -> out_pair[0][b] = lam_r * mean_k(reproj_k), out_pair[1][b] = lam_b * mean_k(f_k disp_k)
__global__ void consistency_finalize(const double* __restrict__ acc, const float* __restrict__ msum,
                                     const float* __restrict__ intr, float f0, float f1, int f_given,
                                     const float* __restrict__ f_dev,
                                     float lam_r, float lam_b, int B, int B_global,
                                     float* __restrict__ out_pair, float* __restrict__ out_loss)
{
  if (f_dev) { f0 = f_dev[0]; f1 = f_dev[1]; f_given = 1; }
  __shared__ float fsh[2];
  __shared__ double tot[32];
  if (threadIdx.x < 2) {
    float f = threadIdx.x ? f1 : f0;
    if (!f_given) {
      float sacc = 0.f;
      for (int bb = 0; bb < B; ++bb) { const float* I = intr + ((size_t)bb * 2 + threadIdx.x) * 4; sacc += I[0]; sacc += I[1]; }
      f = sacc / (float)(2 * B);
    }
    fsh[threadIdx.x] = f;
  }
  __syncthreads();
  double local = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float r = 0.f, d = 0.f;
    for (int k = 0; k < 2; ++k) {
      const float inv = 1.0f / fmaxf(msum[b * 2 + k], 1e-6f);
      r += (float)(acc[(size_t)b * 4 + k * 2 + 0]) * inv;
      d += fsh[k] * ((float)(acc[(size_t)b * 4 + k * 2 + 1]) * inv);
    }
    r = lam_r > 0.f ? lam_r * (0.5f * r) : 0.f;
    d = lam_b > 0.f ? lam_b * (0.5f * d) : 0.f;
    out_pair[b] = r; out_pair[B + b] = d;
    local += (double)r + (double)d;
  }
  local = warp_sum_d(local);
  if ((threadIdx.x & 31) == 0) tot[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += tot[w];
    out_loss[0] = (float)(s / (double)B_global);
  }
}

__global__ void __launch_bounds__(256)
mask_sum_kernel(const float* __restrict__ mask0, const float* __restrict__ mask1, int HW,
                float* __restrict__ msum)
{
  __shared__ float red[8][2];
  const int b = blockIdx.y;
  const float* m0 = mask0 + (size_t)b * HW;
  const float* m1 = mask1 + (size_t)b * HW;
  float s0 = 0.f, s1 = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((HW & 3) == 0) {
    for (long long i = t; i < HW / 4; i += stride) {
      float4 a = ldg_stream4(m0 + 4 * i), c = ldg_stream4(m1 + 4 * i);
      s0 += (a.x + a.y) + (a.z + a.w); s1 += (c.x + c.y) + (c.z + c.w);
    }
  } else {
    for (long long i = t; i < HW; i += stride) { s0 += m0[i]; s1 += m1[i]; }
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s0; red[threadIdx.x >> 5][1] = s1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(msum + b * 2 + threadIdx.x, s);   // 0/1 masks: integer-valued partials, exact below 2^24
  }
}

}  // namespace

extern "C" int cvd_mask_sums(const float* mask0, const float* mask1, int B, int H, int W,
                             float* msum, void* stream)
{
  CVD_CHECK_ARG(mask0 && mask1 && msum, "cvd_mask_sums: null pointer");
  CVD_CHECK_ARG(B > 0 && H > 0 && W > 0, "cvd_mask_sums: bad shape B=%d H=%d W=%d", B, H, W);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(msum, 0, sizeof(float) * 2 * B, st);
  if (e != cudaSuccess) return cvd_fail("cvd_mask_sums: memset: %s", cudaGetErrorString(e));
  const int HW = H * W;
  int gx = (HW / 4 + 255) / 256; if (gx < 1) gx = 1;
  const int cap = (cvd_num_sms() * 8 + B - 1) / B; if (gx > cap) gx = cap < 1 ? 1 : cap;
  mask_sum_kernel<<<dim3(gx, B), 256, 0, st>>>(mask0, mask1, HW, msum);
  CVD_LAUNCH_OK("mask_sum_kernel");
  return 0;
}

extern "C" size_t cvd_consistency_workspace_bytes(int B)
{
  return B > 0 ? (size_t)B * (4 * sizeof(double) + sizeof(PairConst)) : 0;
}

extern "C" int cvd_consistency_fwd_bwd(const float* depth,
                                       const float* flow0, const float* flow1,
                                       const float* mask0, const float* mask1,
                                       const float* extr, const float* intr,
                                       const float* msum, const float* f_dir_host, const float* f_dir_dev,
                                       float lam_r, float lam_b,
                                       int B, int B_global, int H, int W,
                                       void* workspace, float* out_pair, float* out_loss,
                                       float* grad_depth, void* stream)
{
  CVD_CHECK_ARG(depth && flow0 && flow1 && mask0 && mask1 && extr && intr && msum && workspace && out_pair && out_loss,
                "cvd_consistency_fwd_bwd: null pointer");
  CVD_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "cvd_consistency_fwd_bwd: workspace must be 16-byte aligned");
  double* acc = reinterpret_cast<double*>(workspace);
  CVD_CHECK_ARG(B > 0 && B <= 65535 && H > 1 && W > 1 && B_global >= B,
                "cvd_consistency_fwd_bwd: bad shape B=%d B_global=%d H=%d W=%d", B, B_global, H, W);
  CVD_CHECK_ARG((long long)H * W < (1ll << 22), "cvd_consistency_fwd_bwd: image larger than 4 Mpx not supported");
  CVD_CHECK_ARG(2ll * B * H * W < (1ll << 31), "cvd_consistency_fwd_bwd: B*2*H*W must stay below 2^31 elements per call");
  cudaStream_t st = (cudaStream_t)stream;
  const long long HW = (long long)H * W;
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double) * 4 * B, st);
  if (e == cudaSuccess && grad_depth) e = cudaMemsetAsync(grad_depth, 0, sizeof(float) * 2 * HW * B, st);
  if (e != cudaSuccess) return cvd_fail("cvd_consistency_fwd_bwd: memset: %s", cudaGetErrorString(e));
  const int fg = f_dir_host != nullptr;
  const float f0 = fg ? f_dir_host[0] : 0.f, f1 = fg ? f_dir_host[1] : 0.f;
  PairConst* consts = reinterpret_cast<PairConst*>(acc + 4 * (size_t)B);
  consistency_setup<<<(2 * B + 127) / 128, 128, 0, st>>>(extr, intr, msum, f0, f1, fg, f_dir_dev, lam_r, lam_b, B, B_global, consts);
  CVD_LAUNCH_OK("consistency_setup");
  dim3 grid((unsigned)((HW + LOSS_THREADS * PIX - 1) / (LOSS_THREADS * PIX)), B);
  const bool full = (HW % (LOSS_THREADS * PIX)) == 0;
  const bool gr = grad_depth != nullptr, dr = lam_r > 0.f, dd = lam_b > 0.f;   // (:169,176)
#define CVD_LOSS_LAUNCH(G, R, D, F)                                                                    \
  consistency_kernel<G, R, D, F><<<grid, LOSS_THREADS, 0, st>>>(depth, flow0, flow1, mask0, mask1, consts, \
      H, W, acc, grad_depth)
#define CVD_LOSS_F(G, R, D) do { if (full) CVD_LOSS_LAUNCH(G, R, D, true); else CVD_LOSS_LAUNCH(G, R, D, false); } while (0)
#define CVD_LOSS_D(G, R) do { if (dd) CVD_LOSS_F(G, R, true); else CVD_LOSS_F(G, R, false); } while (0)
#define CVD_LOSS_R(G) do { if (dr) CVD_LOSS_D(G, true); else CVD_LOSS_D(G, false); } while (0)
  static const bool no_x4 = getenv("CVD_LOSS_X4") && getenv("CVD_LOSS_X4")[0] == '0';
  // measured (1080x1920, B = 16, B200): forward-only 2.83 TB/s with the x4 variant vs 2.40 TB/s lane-per-pixel; forward +
  // backward 1.55 vs 1.77 TB/s -- with a pixel per lane the scatter REDs of a warp fall into a few 32-byte sectors and are
  // coalesced by the LSU, with four pixels per lane they are not.  So: x4 for the loss-only call, lane-per-pixel with grads.
  const bool x4 = !no_x4 && !grad_depth && (W & 3) == 0 && ((uintptr_t)depth & 15) == 0 && ((uintptr_t)flow0 & 15) == 0 && ((uintptr_t)flow1 & 15) == 0 &&
                  ((uintptr_t)mask0 & 15) == 0 && ((uintptr_t)mask1 & 15) == 0 && (!grad_depth || ((uintptr_t)grad_depth & 15) == 0);
  if (x4) {
    dim3 grid4((unsigned)((HW / 4 + LOSS_THREADS - 1) / LOSS_THREADS), B);
#define CVD_LOSS4_LAUNCH(G, R, D)                                                                      \
  consistency_kernel_x4<G, R, D><<<grid4, LOSS_THREADS, 0, st>>>(depth, flow0, flow1, mask0, mask1, consts, H, W, acc, grad_depth)
#define CVD_LOSS4_D(G, R) do { if (dd) CVD_LOSS4_LAUNCH(G, R, true); else CVD_LOSS4_LAUNCH(G, R, false); } while (0)
#define CVD_LOSS4_R(G) do { if (dr) CVD_LOSS4_D(G, true); else CVD_LOSS4_D(G, false); } while (0)
    if (gr) CVD_LOSS4_R(true); else CVD_LOSS4_R(false);
  } else {
    if (gr) CVD_LOSS_R(true); else CVD_LOSS_R(false);
  }
  CVD_LAUNCH_OK("consistency_kernel");
  consistency_finalize<<<1, 128, 0, st>>>(acc, msum, intr, f0, f1, fg, f_dir_dev, lam_r, lam_b, B, B_global, out_pair, out_loss);
  CVD_LAUNCH_OK("consistency_finalize");
  return 0;
}
