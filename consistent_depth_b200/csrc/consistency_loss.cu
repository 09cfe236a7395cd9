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
// Work decomposition: grid = (pixel-quads / 256, B).  A thread owns 4
// x-adjacent pixels and handles BOTH directions (ref=0->tgt=1 and ref=1->tgt=0)
// so the eight 128-bit streaming loads are all in flight before any math.
// The bilinear gather reads the other frame's depth plane through the
// read-only path (it is L1/L2 resident: flow is spatially coherent); the
// matching backward scatter uses fire-and-forget fp32 REDs that resolve in L2.
#include "cvd_common.cuh"

namespace {

struct DirConst {
  float Rr[9], tr[3], tt[3], Rt[9];
  float fx, fy, cx, cy;       // reference-frame intrinsics
  float fxt, fyt, cxt, cyt;   // target-frame intrinsics
  float cr, cd;               // gradient coefficients incl. 1/(2 B_global) and 1/mask_sum
};

struct PairConst { DirConst d[2]; };

__device__ __forceinline__ void pixel_term(
    const DirConst& c, const float* __restrict__ depth_t, float* __restrict__ grad_t,
    float x, float y, float d, float fu, float fv, float mk,
    int H, int W, float Wm1, float Hm1, bool do_r, bool do_d, bool want_grad,
    float& acc_r, float& acc_d, float& g_direct)
{
  // pixels_to_rays (geometry.py:38-61): ((x-cx)/fx, -(y-cy)/fy, -1)
  const float rx = (x - c.cx) / c.fx;
  const float ry = -(y - c.cy) / c.fy;
  // pixels_to_points (:86-100)
  const float px = rx * d, py = ry * d, pz = -d;
  // reproject_points (:103-128): world = t_r + R_r p ; cam_tgt = R_t^T (world - t_t)
  const float wx = c.tr[0] + (c.Rr[0] * px + c.Rr[1] * py + c.Rr[2] * pz);
  const float wy = c.tr[1] + (c.Rr[3] * px + c.Rr[4] * py + c.Rr[5] * pz);
  const float wz = c.tr[2] + (c.Rr[6] * px + c.Rr[7] * py + c.Rr[8] * pz);
  const float dx = wx - c.tt[0], dy = wy - c.tt[1], dz = wz - c.tt[2];
  const float Qx = c.Rt[0] * dx + c.Rt[3] * dy + c.Rt[6] * dz;
  const float Qy = c.Rt[1] * dx + c.Rt[4] * dy + c.Rt[7] * dz;
  const float Qz = c.Rt[2] * dx + c.Rt[5] * dy + c.Rt[8] * dz;
  // dQ/dd = R_t^T R_r ray
  const float ax = c.Rr[0] * rx + c.Rr[1] * ry - c.Rr[2];
  const float ay = c.Rr[3] * rx + c.Rr[4] * ry - c.Rr[5];
  const float az = c.Rr[6] * rx + c.Rr[7] * ry - c.Rr[8];
  const float mx = c.Rt[0] * ax + c.Rt[3] * ay + c.Rt[6] * az;
  const float my = c.Rt[1] * ax + c.Rt[4] * ay + c.Rt[7] * az;
  const float mz = c.Rt[2] * ax + c.Rt[5] * ay + c.Rt[8] * az;

  const float tx = x + fu, ty = y + fv;   // matched_pixels_tgt = pixels_ref + flow (:165)
  const float inv_nz = 1.0f / (-Qz);
  float g = 0.f;

  if (do_r) {
    // project (:64-83): u = (Qx/-Qz) fx' + cx' ; v = -(Qy/-Qz) fy' + cy'
    const float u = (Qx * inv_nz) * c.fxt + c.cxt;
    const float v = -((Qy * inv_nz) * c.fyt) + c.cyt;
    const float ex = u - tx, ey = v - ty;
    const float dist = sqrtf(ex * ex + ey * ey);      // torch.norm(dim=1) (:170)
    acc_r += mk * dist;
    if (want_grad && dist > 0.f) {                    // norm subgradient 0 at 0
      const float iq2 = inv_nz * inv_nz;              // 1/Qz^2
      const float du = -c.fxt * (mx * Qz - Qx * mz) * iq2;
      const float dv =  c.fyt * (my * Qz - Qy * mz) * iq2;
      g += c.cr * mk * ((ex * du + ey * dv) / dist);
    }
  }
  if (do_d) {
    // sample (:201-208): grid = 2 uv/(W-1,H-1) - 1 ; grid_sample unnormalise
    // (align_corners=False): ((g+1)*size-1)/2 ; border clamp ; bilinear.
    float ix = ((2.0f * tx / Wm1 - 1.0f) + 1.0f) * (float)W;
    ix = (ix - 1.0f) * 0.5f;
    float iy = ((2.0f * ty / Hm1 - 1.0f) + 1.0f) * (float)H;
    iy = (iy - 1.0f) * 0.5f;
    ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
    const float x0f = floorf(ix), y0f = floorf(iy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const int x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - x0f, wx0 = (x0f + 1.0f) - ix;
    const float wy1 = iy - y0f, wy0 = (y0f + 1.0f) - iy;
    const bool xin = x1 < W, yin = y1 < H;           // x0,y0 always in bounds after clamp
    const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
    const int i00 = y0 * W + x0;
    float zs = w00 * __ldg(depth_t + i00);
    if (xin) zs += w10 * __ldg(depth_t + i00 + 1);
    if (yin) zs += w01 * __ldg(depth_t + i00 + W);
    if (xin && yin) zs += w11 * __ldg(depth_t + i00 + W + 1);
    const float zw = -zs;                             // target points' z = -depth
    const float s = 1.0f / Qz - 1.0f / zw;            // disp_diff (:188-189)
    acc_d += mk * fabsf(s);
    if (want_grad) {
      const float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
      const float k0 = c.cd * mk * sg;
      g += k0 * (-mz * (inv_nz * inv_nz));
      const float k1 = -k0 / (zw * zw);               // d|s|/d depth_tgt[tap] = -sg * wt / zw^2
      if (k1 != 0.f) {
        atomicAdd(grad_t + i00, k1 * w00);
        if (xin) atomicAdd(grad_t + i00 + 1, k1 * w10);
        if (yin) atomicAdd(grad_t + i00 + W, k1 * w01);
        if (xin && yin) atomicAdd(grad_t + i00 + W + 1, k1 * w11);
      }
    }
  }
  g_direct = g;
}

template <int VEC>
__global__ void __launch_bounds__(256)
consistency_kernel(const float* __restrict__ depth,
                   const float* __restrict__ flow0, const float* __restrict__ flow1,
                   const float* __restrict__ mask0, const float* __restrict__ mask1,
                   const float* __restrict__ extr, const float* __restrict__ intr,
                   const float* __restrict__ msum,
                   float f0, float f1, int f_given,
                   float lam_r, float lam_b, int B, int B_global, int H, int W,
                   double* __restrict__ acc, float* __restrict__ grad)
{
  __shared__ PairConst pc;
  __shared__ float red[8][4];
  const int b = blockIdx.y;
  const int HW = H * W;
  const bool want_grad = grad != nullptr;

  if (threadIdx.x < 2) {
    const int k = threadIdx.x, t = 1 - k;
    DirConst& c = pc.d[k];
    const float* Er = extr + ((size_t)b * 2 + k) * 12;
    const float* Et = extr + ((size_t)b * 2 + t) * 12;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { c.Rr[i * 3 + j] = Er[i * 4 + j]; c.Rt[i * 3 + j] = Et[i * 4 + j]; }
      c.tr[i] = Er[i * 4 + 3]; c.tt[i] = Et[i * 4 + 3];
    }
    const float* Ir = intr + ((size_t)b * 2 + k) * 4;
    const float* It = intr + ((size_t)b * 2 + t) * 4;
    c.fx = Ir[0]; c.fy = Ir[1]; c.cx = Ir[2]; c.cy = Ir[3];
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
  }
  __syncthreads();

  const bool do_r = lam_r > 0.f, do_d = lam_b > 0.f;   // (:169,176)
  const float Wm1 = (float)(W - 1), Hm1 = (float)(H - 1);
  float a_r[2] = {0.f, 0.f}, a_d[2] = {0.f, 0.f};

  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long p0 = q * VEC;
  if (p0 < HW) {
    const size_t pl = (size_t)b * 2 * HW;              // plane base of pair b (2 planes of HW)
    float dv[2][VEC], fu[2][VEC], fv[2][VEC], mk[2][VEC];
    if constexpr (VEC == 4) {
      auto ld4 = [](const float* p, float (&a)[VEC]) { float4 t = ldg_stream4(p); a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w; };
      ld4(depth + pl + p0, dv[0]);       ld4(depth + pl + HW + p0, dv[1]);
      ld4(flow0 + pl + p0, fu[0]);       ld4(flow0 + pl + HW + p0, fv[0]);
      ld4(flow1 + pl + p0, fu[1]);       ld4(flow1 + pl + HW + p0, fv[1]);
      ld4(mask0 + (size_t)b * HW + p0, mk[0]);
      ld4(mask1 + (size_t)b * HW + p0, mk[1]);
    } else {
      dv[0][0] = depth[pl + p0]; dv[1][0] = depth[pl + HW + p0];
      fu[0][0] = flow0[pl + p0]; fv[0][0] = flow0[pl + HW + p0];
      fu[1][0] = flow1[pl + p0]; fv[1][0] = flow1[pl + HW + p0];
      mk[0][0] = mask0[(size_t)b * HW + p0]; mk[1][0] = mask1[(size_t)b * HW + p0];
    }
    const int y = (int)(p0 / W), x0 = (int)(p0 - (long long)y * W);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float* depth_t = depth + pl + (size_t)(1 - k) * HW;
      float* grad_t = want_grad ? grad + pl + (size_t)(1 - k) * HW : nullptr;
      float gd[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        pixel_term(pc.d[k], depth_t, grad_t, (float)(x0 + j), (float)y,
                   dv[k][j], fu[k][j], fv[k][j], mk[k][j], H, W, Wm1, Hm1,
                   do_r, do_d, want_grad, a_r[k], a_d[k], gd[j]);
      }
      if (want_grad) {
        float* gp = grad + pl + (size_t)k * HW + p0;
        if constexpr (VEC == 4) {
          atomicAdd(reinterpret_cast<float4*>(gp), make_float4(gd[0], gd[1], gd[2], gd[3]));
        } else {
          atomicAdd(gp, gd[0]);
        }
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
    for (int w = 0; w < 8; ++w) s += (double)red[w][threadIdx.x];
    atomicAdd(acc + (size_t)b * 4 + threadIdx.x, s);
  }
}

// acc[b][k][{r,d}] -> out_pair[0][b] = lam_r * mean_k(reproj_k), out_pair[1][b] = lam_b * mean_k(f_k disp_k)
__global__ void consistency_finalize(const double* __restrict__ acc, const float* __restrict__ msum,
                                     const float* __restrict__ intr, float f0, float f1, int f_given,
                                     float lam_r, float lam_b, int B, int B_global,
                                     float* __restrict__ out_pair, float* __restrict__ out_loss)
{
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

extern "C" int cvd_consistency_fwd_bwd(const float* depth,
                                       const float* flow0, const float* flow1,
                                       const float* mask0, const float* mask1,
                                       const float* extr, const float* intr,
                                       const float* msum, const float* f_dir_host,
                                       float lam_r, float lam_b,
                                       int B, int B_global, int H, int W,
                                       double* acc, float* out_pair, float* out_loss,
                                       float* grad_depth, void* stream)
{
  CVD_CHECK_ARG(depth && flow0 && flow1 && mask0 && mask1 && extr && intr && msum && acc && out_pair && out_loss,
                "cvd_consistency_fwd_bwd: null pointer");
  CVD_CHECK_ARG(B > 0 && B <= 65535 && H > 1 && W > 1 && B_global >= B,
                "cvd_consistency_fwd_bwd: bad shape B=%d B_global=%d H=%d W=%d", B, B_global, H, W);
  CVD_CHECK_ARG((long long)H * W < (1ll << 30), "cvd_consistency_fwd_bwd: image too large");
  cudaStream_t st = (cudaStream_t)stream;
  const long long HW = (long long)H * W;
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double) * 4 * B, st);
  if (e == cudaSuccess && grad_depth) e = cudaMemsetAsync(grad_depth, 0, sizeof(float) * 2 * HW * B, st);
  if (e != cudaSuccess) return cvd_fail("cvd_consistency_fwd_bwd: memset: %s", cudaGetErrorString(e));
  const int fg = f_dir_host != nullptr;
  const float f0 = fg ? f_dir_host[0] : 0.f, f1 = fg ? f_dir_host[1] : 0.f;
  const bool vec = (W % 4 == 0) && ((((uintptr_t)depth | (uintptr_t)flow0 | (uintptr_t)flow1 | (uintptr_t)mask0 |
                                      (uintptr_t)mask1 | (uintptr_t)grad_depth) & 15) == 0);
  if (vec) {
    const long long quads = HW / 4;
    dim3 grid((unsigned)((quads + 255) / 256), B);
    consistency_kernel<4><<<grid, 256, 0, st>>>(depth, flow0, flow1, mask0, mask1, extr, intr, msum,
                                                 f0, f1, fg, lam_r, lam_b, B, B_global, H, W, acc, grad_depth);
  } else {
    dim3 grid((unsigned)((HW + 255) / 256), B);
    consistency_kernel<1><<<grid, 256, 0, st>>>(depth, flow0, flow1, mask0, mask1, extr, intr, msum,
                                                 f0, f1, fg, lam_r, lam_b, B, B_global, H, W, acc, grad_depth);
  }
  CVD_LAUNCH_OK("consistency_kernel");
  consistency_finalize<<<1, 128, 0, st>>>(acc, msum, intr, f0, f1, fg, lam_r, lam_b, B, B_global, out_pair, out_loss);
  CVD_LAUNCH_OK("consistency_finalize");
  return 0;
}
