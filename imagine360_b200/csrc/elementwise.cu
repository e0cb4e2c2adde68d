// HBM-bound helper kernels around the tensor-core path (channels-last, 16-byte vectors):
//   * nearest x2 upsample with the pano circular halo folded in   (Upsample3D, resnet.py:86-114 +
//     pad_pano(1)/unpad_pano(2) of MVGenModel.py:449-456)
//   * im2col for the three stride-2 downsample convs               (Downsample3D, resnet.py:117-140 +
//     pad_pano(2)/unpad_pano(1) of MVGenModel.py:305-314)
//   * fused classifier-free guidance + DDIM v-prediction update    (pipeline...dual.py:789-800,
//     scheduling_ddim.py:319-346) reproducing the reference's bf16 rounding after every tensor op
//   * a*x + b*y, avg-pool over frames (resampler.py:251,:264), equirect<->perspective resampling
//     (kornia.remap == grid_sample, e2p.py:77 / p2e.py:70) as a vectorised gather.
#include "common.cuh"
#include "tmap.h"

namespace i360 {

__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16(x)); }

// out [B, 2H, 2*(W + 2*pad_in), C]; out col w' reads in col ((w'/2) - pad_in) mod W
__global__ void upsample2x_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int B, int H, int W, int C,
                                  int pad_in) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  const int nvec = C >> 3, Wo = 2 * (W + 2 * pad_in), Ho = 2 * H;
  const long long total = static_cast<long long>(B) * Ho * Wo * nvec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    long long pix = i / nvec;
    const int wo = static_cast<int>(pix % Wo); pix /= Wo;
    const int ho = static_cast<int>(pix % Ho);
    const int b = static_cast<int>(pix / Ho);
    const int wi = ((wo >> 1) - pad_in + W) % W, hi = ho >> 1;
    reinterpret_cast<uint4*>(out)[i] =
        *reinterpret_cast<const uint4*>(x + ((static_cast<long long>(b) * H + hi) * W + wi) * C + v * 8);
  }
}

// out [B*Ho*Wo, 9*C] for a 3x3 / stride 2 conv with `pad_lo` zero rows/cols before and (2 - pad_lo - (H&1)) after
// (pad_lo = 1: symmetric pad 1 of Downsample3D; pad_lo = 0: the VAE's asymmetric F.pad(0,1,0,1),
// diffusers/models/resnet.py:184); circular != 0: columns wrap instead (pano halo), rows zero-pad
__global__ void im2col_s2_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int B, int H, int W, int C,
                                 int circular, int pad_lo) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  const int nvec = C >> 3, Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * 9 * nvec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    long long r = i / nvec;
    const int tap = static_cast<int>(r % 9); r /= 9;
    const int wo = static_cast<int>(r % Wo); r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const int b = static_cast<int>(r / Ho);
    const int hi = 2 * ho + tap / 3 - pad_lo;
    int wi = 2 * wo + tap % 3 - pad_lo;
    bool ok = hi >= 0 && hi < H;
    if (circular) wi = (wi + W) % W; else ok = ok && wi >= 0 && wi < W;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (ok) val = *reinterpret_cast<const uint4*>(x + ((static_cast<long long>(b) * H + hi) * W + wi) * C + v * 8);
    reinterpret_cast<uint4*>(out)[i] = val;
  }
}

__global__ void axpby_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y, bf16* __restrict__ out, float a,
                             float b, long long n) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // reference order: (y * b) rounded to bf16, then x + that (add_noise_to_condition, MVGenModel.py:11-14)
    const float t = y ? rbf(__bfloat162float(y[i]) * b) : 0.f;
    out[i] = __float2bfloat16(a * __bfloat162float(x[i]) + t);
  }
}

// latent, pred_uncond, pred_cond: same shape bf16.  One fused pass instead of ~12 elementwise launches.
__global__ void cfg_ddim_kernel(const bf16* __restrict__ x, const bf16* __restrict__ vu, const bf16* __restrict__ vc,
                                bf16* __restrict__ out, float guidance, float sa, float sb, float sap, float sbp,
                                long long n) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float xs = __bfloat162float(x[i]), u = __bfloat162float(vu[i]), c = __bfloat162float(vc[i]);
    const float v = rbf(u + rbf(guidance * rbf(c - u)));           // uncond + s * (text - uncond)
    const float x0 = rbf(rbf(sa * xs) - rbf(sb * v));              // pred_original_sample
    const float eps = rbf(rbf(sa * v) + rbf(sb * xs));             // v-prediction -> epsilon
    out[i] = __float2bfloat16(rbf(sap * x0) + rbf(sbp * eps));     // eta = 0
  }
}

// x [B, F, D, C] -> out [B, F/4, D, C], mean over groups of 4 frames
__global__ void avgpool_frames4_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int B, int F, long long DC) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  const int Fo = F / 4;
  const long long total = static_cast<long long>(B) * Fo * DC;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = i % DC; const long long r = i / DC;
    const int fo = static_cast<int>(r % Fo); const int b = static_cast<int>(r / Fo);
    const bf16* src = x + (static_cast<long long>(b) * F + fo * 4) * DC + e;
    const float s = __bfloat162float(src[0]) + __bfloat162float(src[DC]) + __bfloat162float(src[2 * DC]) +
                    __bfloat162float(src[3 * DC]);
    out[i] = __float2bfloat16(s * 0.25f);
  }
}

// grid_sample(align_corners=True, padding zeros) on channels-first fp32 images:
//   img [N, C, Hi, Wi], grid [N, Ho, Wo, 2] NORMALISED coordinates (x, y) in fp32, out [N, C, Ho, Wo].
// One thread per output pixel, looping channels (the 4 taps and weights are reused by every channel).
__global__ void grid_sample_kernel(const float* __restrict__ img, const float* __restrict__ grid, float* __restrict__ out,
                                   int N, int C, int Hi, int Wi, int Ho, int Wo, int nearest) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  const long long total = static_cast<long long>(N) * Ho * Wo;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / (static_cast<long long>(Ho) * Wo));
    const long long po = i % (static_cast<long long>(Ho) * Wo);
    const float gx = grid[i * 2], gy = grid[i * 2 + 1];
    const float ix = (gx + 1.f) * 0.5f * (Wi - 1), iy = (gy + 1.f) * 0.5f * (Hi - 1);
    const float* src = img + static_cast<long long>(n) * C * Hi * Wi;
    float* dst = out + static_cast<long long>(n) * C * Ho * Wo + po;
    if (nearest) {
      const int xn = static_cast<int>(nearbyintf(ix)), yn = static_cast<int>(nearbyintf(iy));
      const bool ok = xn >= 0 && xn < Wi && yn >= 0 && yn < Hi;
      for (int c = 0; c < C; ++c)
        dst[static_cast<long long>(c) * Ho * Wo] = ok ? src[(static_cast<long long>(c) * Hi + yn) * Wi + xn] : 0.f;
    } else {
      const float fx = floorf(ix), fy = floorf(iy);
      const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy), x1 = x0 + 1, y1 = y0 + 1;
      const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
      const bool okx0 = x0 >= 0 && x0 < Wi, okx1 = x1 >= 0 && x1 < Wi, oky0 = y0 >= 0 && y0 < Hi, oky1 = y1 >= 0 && y1 < Hi;
      for (int c = 0; c < C; ++c) {
        const float* pc = src + static_cast<long long>(c) * Hi * Wi;
        float acc = 0.f;
        if (oky0 && okx0) acc += pc[static_cast<long long>(y0) * Wi + x0] * (wx0 * wy0);
        if (oky0 && okx1) acc += pc[static_cast<long long>(y0) * Wi + x1] * (wx1 * wy0);
        if (oky1 && okx0) acc += pc[static_cast<long long>(y1) * Wi + x0] * (wx0 * wy1);
        if (oky1 && okx1) acc += pc[static_cast<long long>(y1) * Wi + x1] * (wx1 * wy1);
        dst[static_cast<long long>(c) * Ho * Wo] = acc;
      }
    }
  }
}

static inline unsigned grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms()) * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

}  // namespace i360

using namespace i360;

extern "C" int i360_upsample2x_nhwc(const void* x, void* out, int B, int H, int W, int C, int pad_in, void* stream) {
  if (!x || !out || (C % 8) || B <= 0 || pad_in < 0 || pad_in > W) return I360_ERR_ARG;
  const long long total = static_cast<long long>(B) * 2 * H * 2 * (W + 2 * pad_in) * (C / 8);
  launch_k(upsample2x_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const bf16*>(x), static_cast<bf16*>(out), B, H, W, C, pad_in);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}

extern "C" int i360_im2col3x3_s2_nhwc(const void* x, void* out, int B, int H, int W, int C, int circular, int pad_lo,
                                      void* stream) {
  if (!x || !out || (C % 8) || (H % 2) || (W % 2) || B <= 0 || pad_lo < 0 || pad_lo > 1) return I360_ERR_ARG;
  const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * 9 * (C / 8);
  launch_k(im2col_s2_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const bf16*>(x), static_cast<bf16*>(out), B, H, W, C, circular, pad_lo);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}

extern "C" int i360_axpby_bf16(const void* x, const void* y, void* out, float a, float b, long long n, void* stream) {
  if (!x || !out || n <= 0) return I360_ERR_ARG;
  launch_k(axpby_kernel, dim3(grid_for(n, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const bf16*>(x), static_cast<const bf16*>(y), static_cast<bf16*>(out), a, b, n);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}

extern "C" int i360_cfg_ddim_step_bf16(const void* latent, const void* pred_uncond, const void* pred_cond, void* out,
                                       float guidance, float sqrt_a_t, float sqrt_1m_a_t, float sqrt_a_prev,
                                       float sqrt_1m_a_prev, long long n, void* stream) {
  if (!latent || !pred_uncond || !pred_cond || !out || n <= 0) return I360_ERR_ARG;
  launch_k(cfg_ddim_kernel, dim3(grid_for(n, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const bf16*>(latent), static_cast<const bf16*>(pred_uncond), static_cast<const bf16*>(pred_cond),
      static_cast<bf16*>(out), guidance, sqrt_a_t, sqrt_1m_a_t, sqrt_a_prev, sqrt_1m_a_prev, n);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}

extern "C" int i360_avgpool_frames4_bf16(const void* x, void* out, int B, int F, long long DC, void* stream) {
  if (!x || !out || B <= 0 || F < 4 || DC <= 0) return I360_ERR_ARG;
  const long long total = static_cast<long long>(B) * (F / 4) * DC;
  launch_k(avgpool_frames4_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const bf16*>(x), static_cast<bf16*>(out), B, F, DC);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}

extern "C" int i360_grid_sample_f32(const float* img, const float* grid, float* out, int N, int C, int Hi, int Wi,
                                    int Ho, int Wo, int nearest, void* stream) {
  if (!img || !grid || !out || N <= 0 || C <= 0) return I360_ERR_ARG;
  const long long total = static_cast<long long>(N) * Ho * Wo;
  launch_k(grid_sample_kernel, dim3(grid_for(total, 128)), dim3(128), 0, static_cast<cudaStream_t>(stream), img, grid, out, N, C, Hi, Wi,
                                                                                           Ho, Wo, nearest);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}
