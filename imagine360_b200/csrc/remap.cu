// Bicubic equirectangular <-> perspective resampling of 8-bit frames on the GPU, bit-exact with
// cv2.remap(src, mapx, mapy, INTER_CUBIC, borderMode=BORDER_WRAP) -- the call the reference makes on the CPU for every
// (frame, view) pair in src/utils/pano_utils/Equirec2Perspec.py:61 (process_equi, inference_dual_p2e.py:113-144;
// get_anchor_target, animatediff/utils/video_mask.py:158-217) and Perspec2Equirec.py:74 (pers2pano_vid,
// inference_dual_p2e.py:293-301).
//
// OpenCV evaluates 8-bit remaps in fixed point: the float maps are quantised to 1/32 pixel with round-half-even,
// the 4x4 taps come from a 32x32 table of int16 weights (scale 2^15, A = -0.75 kernels evaluated in float32, sum
// forced to 2^15), and the result is saturate_u8((sum + 2^14) >> 15).  All of that is integer work and is reproduced
// exactly; the table itself is built on the host by i360_remap_cubic_table_i16 with OpenCV's own float32 sequence.
//
// One thread per output pixel, x fastest: map reads and CHW float stores are coalesced, the 48 source bytes of a pixel
// are a 2-D local gather served by L1/L2 (the whole source video is read from HBM once).  HBM bound: algorithmic bytes =
// source frames once + output once (+ maps once).
#include "common.cuh"
#include "tmap.h"
#include <math.h>

namespace i360 {

struct RemapParams {
  const uint8_t* src; int n_img, H, W;          // [n_img, H, W, 3], base 4-byte aligned
  long long src_bytes;
  const float* mapx; const float* mapy;          // [n_map, h, w]
  const uint8_t* keep;                           // optional [n_map, h, w]: output multiplied by 0/1 (P2E's mask)
  int n_map, h, w;
  int paired;                                    // 0: every image x every map; 1: image i with map i
  const short* tab;                              // [1024][16]
  uint8_t* out_u8;                               // [n_img, n_out, h, w, 3] or null
  float* out_f32;                                // mode 1: [n_img, n_out, 3, h, w] = u8 / 127.5 - 1; mode 2: [n_img, n_out, 1, h, w] = any(u8 > 0)
  int f32_mode;
};

__device__ __forceinline__ int wrap_index(int p, int n) {   // borderInterpolate(..., BORDER_WRAP)
  if (p < 0) p -= ((p - n + 1) / n) * n;
  if (p >= n) p %= n;
  return p;
}

__global__ void __launch_bounds__(256)
remap_cubic_wrap_kernel(const RemapParams p) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  const long long hw = static_cast<long long>(p.h) * p.w;
  const int n_out = p.paired ? 1 : p.n_map;
  const long long total = hw * n_out * p.n_img;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = idx % hw;
    const int v = static_cast<int>((idx / hw) % n_out);
    const int img = static_cast<int>(idx / (hw * n_out));
    const int mi = p.paired ? img : v;
    const float mx = __ldg(p.mapx + mi * hw + pix), my = __ldg(p.mapy + mi * hw + pix);
    const int sx = __float2int_rn(mx * 32.0f), sy = __float2int_rn(my * 32.0f);      // cvRound: half to even
    const int fidx = (sy & 31) * 32 + (sx & 31);
    const int ix = max(-32768, min(32767, sx >> 5)) - 1, iy = max(-32768, min(32767, sy >> 5)) - 1;
    short wt[16];
    {
      const uint4* t4 = reinterpret_cast<const uint4*>(p.tab + fidx * 16);
      const uint4 a = __ldg(t4), b = __ldg(t4 + 1);
      *reinterpret_cast<uint4*>(wt) = a; *reinterpret_cast<uint4*>(wt + 8) = b;
    }
    const long long img_off = static_cast<long long>(img) * p.H * p.W * 3;
    const uint8_t* S = p.src + img_off;
    int acc0 = 0, acc1 = 0, acc2 = 0;
    const bool inside = (ix >= 0) && (ix + 3 < p.W) && (iy >= 0) && (iy + 3 < p.H);
    if (inside) {
      // 4 taps x 3 channels of one row are 12 contiguous bytes at an arbitrary byte offset: fetch the 4 aligned
      // 32-bit words that cover them and funnel-shift, instead of 12 byte loads (the kernel is LSU bound otherwise)
      const long long row0 = img_off + (static_cast<long long>(iy) * p.W + ix) * 3;
      const bool words_ok = row0 + 3LL * p.W * 3 + 16 <= p.src_bytes;
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const long long b = row0 + static_cast<long long>(ky) * p.W * 3;
        uint32_t q0, q1, q2;
        if (words_ok) {
          const uint32_t* a = reinterpret_cast<const uint32_t*>(p.src) + (b >> 2);
          const uint32_t w0 = __ldg(a), w1 = __ldg(a + 1), w2 = __ldg(a + 2), w3 = __ldg(a + 3);
          const uint32_t sh = static_cast<uint32_t>(b & 3) * 8;
          q0 = __funnelshift_r(w0, w1, sh); q1 = __funnelshift_r(w1, w2, sh); q2 = __funnelshift_r(w2, w3, sh);
        } else {
          const uint8_t* r = p.src + b;
          q0 = r[0] | (r[1] << 8) | (r[2] << 16) | (static_cast<uint32_t>(r[3]) << 24);
          q1 = r[4] | (r[5] << 8) | (r[6] << 16) | (static_cast<uint32_t>(r[7]) << 24);
          q2 = r[8] | (r[9] << 8) | (r[10] << 16) | (static_cast<uint32_t>(r[11]) << 24);
        }
        // bytes: q0 = r0 g0 b0 r1, q1 = g1 b1 r2 g2, q2 = b2 r3 g3 b3
        const int w0 = wt[ky * 4 + 0], w1 = wt[ky * 4 + 1], w2 = wt[ky * 4 + 2], w3 = wt[ky * 4 + 3];
        acc0 += static_cast<int>(q0 & 255) * w0 + static_cast<int>(q0 >> 24) * w1 + static_cast<int>((q1 >> 16) & 255) * w2 +
                static_cast<int>((q2 >> 8) & 255) * w3;
        acc1 += static_cast<int>((q0 >> 8) & 255) * w0 + static_cast<int>(q1 & 255) * w1 + static_cast<int>(q1 >> 24) * w2 +
                static_cast<int>((q2 >> 16) & 255) * w3;
        acc2 += static_cast<int>((q0 >> 16) & 255) * w0 + static_cast<int>((q1 >> 8) & 255) * w1 + static_cast<int>(q2 & 255) * w2 +
                static_cast<int>(q2 >> 24) * w3;
      }
    } else {
      int xs[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) xs[k] = wrap_index(ix + k, p.W) * 3;
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const uint8_t* r = S + static_cast<long long>(wrap_index(iy + ky, p.H)) * p.W * 3;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          const int wgt = wt[ky * 4 + kx];
          acc0 += __ldg(r + xs[kx] + 0) * wgt; acc1 += __ldg(r + xs[kx] + 1) * wgt; acc2 += __ldg(r + xs[kx] + 2) * wgt;
        }
      }
    }
    int o0 = min(255, max(0, (acc0 + 16384) >> 15));
    int o1 = min(255, max(0, (acc1 + 16384) >> 15));
    int o2 = min(255, max(0, (acc2 + 16384) >> 15));
    if (p.keep && !__ldg(p.keep + mi * hw + pix)) { o0 = 0; o1 = 0; o2 = 0; }
    const long long oi = static_cast<long long>(img) * n_out + v;
    if (p.out_u8) {
      uint8_t* d = p.out_u8 + (oi * hw + pix) * 3;
      d[0] = static_cast<uint8_t>(o0); d[1] = static_cast<uint8_t>(o1); d[2] = static_cast<uint8_t>(o2);
    }
    if (p.out_f32) {
      if (p.f32_mode == 2) {
        p.out_f32[oi * hw + pix] = (o0 > 0 || o1 > 0 || o2 > 0) ? 1.0f : 0.0f;
      } else {
        float* d = p.out_f32 + oi * 3 * hw + pix;
        d[0] = __fdiv_rn(static_cast<float>(o0), 127.5f) - 1.0f;       // numpy: (img.astype(float32) / 127.5) - 1
        d[hw] = __fdiv_rn(static_cast<float>(o1), 127.5f) - 1.0f;
        d[2 * hw] = __fdiv_rn(static_cast<float>(o2), 127.5f) - 1.0f;
      }
    }
  }
}

// float32 frames [n, 3, H, W] in (-1, 1) (or (0, 1)) -> uint8 [n, H, W, 3], exactly like
// ((x + 1) * 127.5).permute(1, 2, 0).numpy().astype(np.uint8)  /  (x * 255)...  (inference_dual_p2e.py:122-129) and like
// save_videos_grid's (x * 255).numpy().astype(np.uint8) after the optional (x + 1) / 2 (animatediff/utils/util.py:55-72):
// float32 arithmetic, then C truncation toward zero.  Frame n of the input starts at x + n * frame_stride.
__global__ void __launch_bounds__(256)
frames_to_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, long long n, long long hw, int back_norm,
                    long long frame_stride, long long chan_stride) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  const long long total = n * hw;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long img = idx / hw, pix = idx % hw;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = x[img * frame_stride + c * chan_stride + pix];
      // mode 0: x * 255; 1: (x + 1) * 127.5; 2: ((x + 1) / 2) * 255 (save_videos_grid with rescale) -- fp32 op by op
      const float s = back_norm == 1 ? __fmul_rn(__fadd_rn(v, 1.0f), 127.5f)
                    : back_norm == 2 ? __fmul_rn(__fdiv_rn(__fadd_rn(v, 1.0f), 2.0f), 255.0f) : __fmul_rn(v, 255.0f);
      out[idx * 3 + c] = static_cast<uint8_t>(static_cast<int>(s));     // truncation; inputs are in range by contract
    }
  }
}

}  // namespace i360

using namespace i360;

// OpenCV's fixed-point bicubic table (imgwarp.cpp: interpolateCubic, initInterTab2D with fixpt = true):
// out[(fy * 32 + fx) * 16 + ky * 4 + kx].  Pure host function (no GPU needed).
extern "C" int i360_remap_cubic_table_i16(short* out) {
  if (!out) return I360_ERR_ARG;
  float tab1[32][4];
  const float A = -0.75f, scale = 1.f / 32;
  for (int i = 0; i < 32; ++i) {
    const float x = i * scale;
    float* c = tab1[i];
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
  }
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) {
      short* it = out + (i * 32 + j) * 16;
      int isum = 0;
      for (int k1 = 0; k1 < 4; ++k1)
        for (int k2 = 0; k2 < 4; ++k2) {
          const float v = tab1[i][k1] * tab1[j][k2];
          long r = lrintf(v * 32768.f);                       // saturate_cast<short>(float): round half to even
          r = r < -32768 ? -32768 : (r > 32767 ? 32767 : r);
          it[k1 * 4 + k2] = static_cast<short>(r);
          isum += static_cast<int>(r);
        }
      if (isum != 32768) {
        const int diff = isum - 32768;
        int mk = 2 * 4 + 2, Mk = 2 * 4 + 2;
        for (int k1 = 2; k1 < 4; ++k1)
          for (int k2 = 2; k2 < 4; ++k2) {
            if (it[k1 * 4 + k2] < it[mk]) mk = k1 * 4 + k2;
            else if (it[k1 * 4 + k2] > it[Mk]) Mk = k1 * 4 + k2;
          }
        if (diff < 0) it[Mk] = static_cast<short>(it[Mk] - diff);
        else it[mk] = static_cast<short>(it[mk] - diff);
      }
    }
  return I360_OK;
}

extern "C" int i360_remap_cubic_wrap_u8(const void* src, int n_img, int H, int W, const float* mapx, const float* mapy,
                                        const void* keep, int n_map, int h, int w, int paired, const short* table,
                                        void* out_u8, float* out_f32, int f32_mode, void* stream) {
  if (!src || !mapx || !mapy || !table || (!out_u8 && !out_f32)) return I360_ERR_ARG;
  if (n_img <= 0 || n_map <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0) return I360_ERR_ARG;
  if (paired && n_map != n_img) return I360_ERR_ARG;
  if (out_f32 && f32_mode != 1 && f32_mode != 2) return I360_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(table) & 15) != 0) return I360_ERR_ARG;
  RemapParams p;
  if ((reinterpret_cast<uintptr_t>(src) & 3) != 0) return I360_ERR_ARG;
  p.src = static_cast<const uint8_t*>(src); p.n_img = n_img; p.H = H; p.W = W;
  p.src_bytes = static_cast<long long>(n_img) * H * W * 3;
  p.mapx = mapx; p.mapy = mapy; p.keep = static_cast<const uint8_t*>(keep);
  p.n_map = n_map; p.h = h; p.w = w; p.paired = paired; p.tab = table;
  p.out_u8 = static_cast<uint8_t*>(out_u8); p.out_f32 = out_f32; p.f32_mode = f32_mode;
  const long long total = static_cast<long long>(h) * w * (paired ? 1 : n_map) * n_img;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  launch_k(remap_cubic_wrap_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), p);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}

extern "C" int i360_frames_to_u8_nhwc(const float* x, long long frame_stride, long long chan_stride, void* out, long long n,
                                      int H, int W, int mode, void* stream) {
  if (!x || !out || n <= 0 || H <= 0 || W <= 0 || mode < 0 || mode > 2) return I360_ERR_ARG;
  const int back_norm = mode;
  const long long total = n * H * W;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  launch_k(frames_to_u8_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, static_cast<uint8_t*>(out), n, static_cast<long long>(H) * W, back_norm, frame_stride, chan_stride);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}
