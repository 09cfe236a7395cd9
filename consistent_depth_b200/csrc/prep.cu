// Operand preparation for the TMA-fed conv kernels (conv2.cu, wgrad2.cu).
//
// The tensor cores consume bf16; parity with the reference's fp32 convolutions (hourglass.py:27,39,42) needs the
// split v = hi + lo (tc_common.cuh).  The first-generation kernels redo "per-channel transform + split" in every
// consumer's producer warps (a forward activation is read by its conv, by that conv's wgrad, ...).  Here each tensor
// is transformed and split ONCE into two bf16 planes in a chunk-planar layout
//        Z[plane = hi|lo][n][c/8][y][x][c%8]          (16 B per pixel per 8-channel chunk)
// which is what a TMA box load drops into shared memory as the UMMA SWIZZLE_NONE canonical layout
// [chunk][row][col][16 B] without any thread touching it, and with "same" zero padding for free (out-of-bounds
// box elements are zero-filled by the TMA unit).
//
//   cvd_prep_act : v = relu(a[c]*x + b[c])                  BatchNorm2d(train)+ReLU of the producer (hourglass.py:28-29,40-41)
//   cvd_prep_grad: y = a*x + b; g = (!relu || y > 0) ? dy : 0; v = c0*g - c1 - c2*y
//                  backward of that BatchNorm+ReLU (autograd, depth_fine_tuning.py:282)
// Logical channels of the source VIEW become dense channels of Z (the view's gap disappears).
// HBM-bound elementwise passes: 4 (8) B read + 4 B written per element.
#include "cvd_common.cuh"
#include "tc_common.cuh"

namespace {

struct PrepArgs {
  const float* x; const float* dy; const float* a; const float* b; const float4* bw;
  int ct, c0, n0, gap, dy_ct, dy_c0, dy_n0, dy_gap, relu, cvalid;
  uint4* zhi; uint4* zlo;           // planes; lo == nullptr: precision 1 (hi only)
  int zc8, zc8_off;                 // chunks per image in Z, first chunk written
  long long npix_img;               // H*W
  long long npix;                   // N*H*W
  int nchunks;                      // chunks to produce (ceil(cvalid/8) rounded up to the conv's 16-channel k-blocks)
};

__device__ __forceinline__ int vphys(int c, int c0, int n0, int gap) { return c0 + c + (c >= n0 ? gap : 0); }

// block = 256 threads = 32 pixels x 8 chunks (warp w: chunk blockIdx.y*8 + w, lane: pixel)
template <int MODE>
__global__ void __launch_bounds__(256)
prep_kernel(const PrepArgs p)
{
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int chunk = blockIdx.y * 8 + w;
  if (chunk >= p.nchunks) return;
  const int cl = chunk * 8;
  float av[8], bv[8], c0v[8], c1v[8], c2v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    av[i] = 1.f; bv[i] = 0.f; c0v[i] = 0.f; c1v[i] = 0.f; c2v[i] = 0.f;
    if (cl + i < p.cvalid) {
      const int pc = vphys(cl + i, p.c0, p.n0, p.gap);
      if (p.a) { av[i] = __ldg(p.a + pc); bv[i] = __ldg(p.b + pc); }
      if (MODE == CVD_XF_BNBWD) { const float4 q = __ldg(p.bw + pc); c0v[i] = q.x; c1v[i] = q.y; c2v[i] = q.z; }
    }
  }
  const bool any = cl < p.cvalid, sec = cl + 4 < p.cvalid;
  const int pcx = vphys(cl, p.c0, p.n0, p.gap);
  const int pcd = MODE == CVD_XF_BNBWD ? vphys(cl, p.dy_c0, p.dy_n0, p.dy_gap) : 0;
  const long long stride = (long long)gridDim.x * 32;
  for (long long pix = (long long)blockIdx.x * 32 + lane; pix < p.npix; pix += stride) {
    float xv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (any) {
      const float* xp = p.x + pix * p.ct + pcx;
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(xp));
      xv[0] = q0.x; xv[1] = q0.y; xv[2] = q0.z; xv[3] = q0.w;
      if (sec) { const float4 q1 = __ldg(reinterpret_cast<const float4*>(xp + 4)); xv[4] = q1.x; xv[5] = q1.y; xv[6] = q1.z; xv[7] = q1.w; }
      if (MODE == CVD_XF_BNBWD) {
        const float* dp = p.dy + pix * p.dy_ct + pcd;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(dp));
        dv[0] = g0.x; dv[1] = g0.y; dv[2] = g0.z; dv[3] = g0.w;
        if (sec) { const float4 g1 = __ldg(reinterpret_cast<const float4*>(dp + 4)); dv[4] = g1.x; dv[5] = g1.y; dv[6] = g1.z; dv[7] = g1.w; }
      }
    }
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float y = fmaf(av[i], xv[i], bv[i]);
      if (MODE == CVD_XF_AFFINE) {
        v[i] = p.relu ? fmaxf(y, 0.f) : y;
      } else {
        const float g = (!p.relu || y > 0.f) ? dv[i] : 0.f;
        v[i] = c0v[i] * g - c1v[i] - c2v[i] * y;
      }
      if (cl + i >= p.cvalid || (i >= 4 && !sec)) v[i] = 0.f;
    }
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    const long long n = pix / p.npix_img, q = pix - n * p.npix_img;
    const long long o = (n * p.zc8 + p.zc8_off + chunk) * p.npix_img + q;
    p.zhi[o] = hi;
    if (p.zlo) p.zlo[o] = lo;
  }
}

}  // namespace

// z: [2 planes][N][zc8][H*W] x 16 B (bf16 hi plane then lo plane); writes chunks [zc8_off, zc8_off + ceil16(C)/8)
extern "C" int cvd_prep_operand(const cvd_src_t* src, int C, long long N, long long HW, void* z, int zc8, int zc8_off,
                                int precision, void* stream)
{
  CVD_CHECK_ARG(src && src->x && z, "cvd_prep_operand: null pointer");
  CVD_CHECK_ARG(C > 0 && N > 0 && HW > 0 && zc8 > 0 && zc8_off >= 0, "cvd_prep_operand: bad shape");
  CVD_CHECK_ARG(precision == 1 || precision == 3, "cvd_prep_operand: precision must be 1 or 3");
  CVD_CHECK_ARG(src->mode == CVD_XF_AFFINE || (src->mode == CVD_XF_BNBWD && src->dy && src->bw && src->a && src->b),
                "cvd_prep_operand: bad source transform");
  CVD_CHECK_ARG((src->c_total & 3) == 0 && (src->c_off & 3) == 0 && (src->n0 & 7) == 0 && (src->gap & 3) == 0,
                "cvd_prep_operand: source view must be 4-channel aligned");
  PrepArgs p{};
  p.x = src->x; p.dy = src->dy; p.a = src->a; p.b = src->b; p.bw = reinterpret_cast<const float4*>(src->bw);
  p.ct = src->c_total; p.c0 = src->c_off; p.n0 = src->n0 > 0 ? src->n0 : (1 << 30); p.gap = src->gap;
  p.dy_ct = src->dy_ctotal; p.dy_c0 = src->dy_coff; p.dy_n0 = src->dy_n0 > 0 ? src->dy_n0 : (1 << 30); p.dy_gap = src->dy_gap;
  p.relu = src->relu; p.cvalid = (C + 3) & ~3;
  p.nchunks = ((C + 15) / 16) * 2;
  CVD_CHECK_ARG(zc8_off + p.nchunks <= zc8, "cvd_prep_operand: chunks [%d, %d) exceed the Z buffer's %d", zc8_off, zc8_off + p.nchunks, zc8);
  p.zc8 = zc8; p.zc8_off = zc8_off; p.npix_img = HW; p.npix = N * HW;
  p.zhi = reinterpret_cast<uint4*>(z);
  p.zlo = precision == 3 ? p.zhi + (size_t)N * zc8 * HW : nullptr;
  long long bx = (p.npix + 31) / 32;
  const long long cap = (long long)cvd_num_sms() * 8;
  if (bx > cap) bx = cap;
  dim3 grid((unsigned)bx, (unsigned)((p.nchunks + 7) / 8));
  if (src->mode == CVD_XF_AFFINE) prep_kernel<CVD_XF_AFFINE><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  else prep_kernel<CVD_XF_BNBWD><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  CVD_LAUNCH_OK("prep_kernel");
  return 0;
}
