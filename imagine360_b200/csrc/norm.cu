// Channels-last normalisation kernels (HBM-bound: one coalesced 16-byte-vector read per element,
// fp32 statistics, bf16 out).
//
//  * GroupNorm (per-image statistics = InflatedGroupNorm, animatediff/models/resnet.py:9-17, and the
//    plain nn.GroupNorm of Transformer3DModel / TemporalTransformer3DModel / the VAE), split into
//      gn_stats : sum / sum-of-squares per (image, group), accumulated in fp64 atomics
//      gn_apply : (x - mean) * rstd * gamma + beta  [-> SiLU], written as a dense NHWC tensor.
//    Both read a VIRTUAL tensor: the channel concatenation of up to two sources (UNet skip concat,
//    MVGenModel.py:399) circularly padded by `pad` columns (pad_pano, src/utils/pano.py:75-92), so the
//    statistics are those of the padded tensor exactly as in the reference (SURVEY.md trap 1) while the
//    concat / pad copies are never materialised on their own.
//  * LayerNorm over the last dim with optional additive tables before (WarpAttn query PE,
//    src/modules/transformer.py:155-159) and after (temporal sinusoidal PE, motion_module.py:350).
#include "common.cuh"
#include "tmap.h"
#include <stdlib.h>

namespace i360 {

struct GnSrc {
  const bf16* x1; const bf16* x2;
  int C1, C2;       // channels of each source (C2 = 0 when unused); both multiples of 8
  int H, Wsrc, pad; // virtual width = Wsrc + 2*pad, column wv reads (wv - pad) mod Wsrc
};

// pix is the pixel index inside one image of the virtual (padded) tensor; 32-bit on purpose (a 64-bit divide
// per 16-byte load used to cost more issue slots than the arithmetic of the whole vector)
__device__ __forceinline__ uint4 gn_load(const GnSrc& s, int b, int pix, int v, int Wv) {
  long long p;
  if (s.pad == 0) {
    p = static_cast<long long>(b) * s.H * s.Wsrc + pix;
  } else {
    const int h = static_cast<int>(static_cast<unsigned>(pix) / static_cast<unsigned>(Wv));
    int w = pix - h * Wv - s.pad;
    if (w < 0) w += s.Wsrc; else if (w >= s.Wsrc) w -= s.Wsrc;          // pad <= Wsrc
    p = (static_cast<long long>(b) * s.H + h) * s.Wsrc + w;
  }
  const int nv1 = s.C1 >> 3;
  const bf16* ptr = (v < nv1) ? s.x1 + p * s.C1 + v * 8 : s.x2 + p * s.C2 + (v - nv1) * 8;
  return *reinterpret_cast<const uint4*>(ptr);
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// Loads in flight per thread.  4 (not 8) keeps both kernels under 56 registers so three 384-thread CTAs stay resident
// per SM: the ncu capture of the 8-deep version showed ONE resident CTA (88 regs) and 18 % warp occupancy.
constexpr int kGnDepth = 4;

// grid (chunks, B); block = R * nvec threads; dynamic smem = R*C*2 floats
__global__ void __launch_bounds__(384, 3) gn_stats_kernel(GnSrc s, int groups, int chunk_pixels, double* __restrict__ stats) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  extern __shared__ float sm[];
  const int C = s.C1 + s.C2, nvec = C >> 3, R = blockDim.x / nvec;
  const int Wv = s.Wsrc + 2 * s.pad;
  const int npix = s.H * Wv;
  const int b = blockIdx.y;
  const int r = threadIdx.x / nvec, v = threadIdx.x % nvec;
  const int p0 = blockIdx.x * chunk_pixels;
  int p1 = p0 + chunk_pixels; if (p1 > npix) p1 = npix;
  float2 s1[4], s2[4];                            // packed pairs: FADD2 / FFMA2
#pragma unroll
  for (int j = 0; j < 4; ++j) { s1[j] = make_float2(0.f, 0.f); s2[j] = make_float2(0.f, 0.f); }
  auto acc = [&](const uint4& u) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      s1[j] = fadd2(s1[j], f); s2[j] = ffma2(f, f, s2[j]);
    }
  };
  if (r < R) {
    int p = p0 + r;
    for (; p + (kGnDepth - 1) * R < p1; p += kGnDepth * R) {   // kGnDepth independent 16-byte loads in flight per thread
      uint4 u[kGnDepth];
#pragma unroll
      for (int k = 0; k < kGnDepth; ++k) u[k] = gn_load(s, b, p + k * R, v, Wv);
#pragma unroll
      for (int k = 0; k < kGnDepth; ++k) acc(u[k]);
    }
    for (; p < p1; p += R) acc(gn_load(s, b, p, v, Wv));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sm[(r * C + v * 8 + 2 * j) * 2] = s1[j].x;     sm[(r * C + v * 8 + 2 * j) * 2 + 1] = s2[j].x;
      sm[(r * C + v * 8 + 2 * j + 1) * 2] = s1[j].y; sm[(r * C + v * 8 + 2 * j + 1) * 2 + 1] = s2[j].y;
    }
  }
  __syncthreads();
  // fold the R pixel lanes per channel (conflict-free: consecutive threads read consecutive channels) ...
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int rr = 0; rr < R; ++rr) { a += sm[(rr * C + c) * 2]; q += sm[(rr * C + c) * 2 + 1]; }
    sm[c * 2] = a; sm[c * 2 + 1] = q;             // row 0 of the table: only this thread touches channel c
  }
  __syncthreads();
  // ... then the channels of each group, in fp64, into the global accumulators
  const int gs = C / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double a = 0.0, q = 0.0;
    for (int c = g * gs; c < (g + 1) * gs; ++c) { a += sm[c * 2]; q += sm[c * 2 + 1]; }
    atomicAdd(&stats[(static_cast<long long>(b) * groups + g) * 2], a);
    atomicAdd(&stats[(static_cast<long long>(b) * groups + g) * 2 + 1], q);
  }
}

// grid (chunks, B); block = R * nvec; dynamic smem = 2*C floats (scale, shift)
template <bool CHAN>
__global__ void __launch_bounds__(384, 3) gn_apply_kernel(GnSrc s, int groups, int chunk_pixels, const double* __restrict__ stats,
                                const bf16* __restrict__ gamma, const bf16* __restrict__ beta, float eps,
                                int do_silu, double count, bf16* __restrict__ out, const double* __restrict__ chan) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  extern __shared__ __align__(16) float sm[];
  const int C = s.C1 + s.C2, nvec = C >> 3, R = blockDim.x / nvec;
  const int Wv = s.Wsrc + 2 * s.pad;
  const int npix = s.H * Wv;
  const int b = blockIdx.y;
  const int gs = C / groups;
  const double n = count > 0.0 ? count : static_cast<double>(npix) * gs;
  // chan != nullptr: per-CHANNEL (sum, sumsq) written by the producing conv's epilogue (fp64) are folded into groups here
  __shared__ double sgd[CHAN ? 128 : 2];         // [groups <= 64][2]
  if (CHAN) {
    // coalesced copy of this image's channel sums into sm (reused for scale / shift below), then one thread per group adds its
    // gs consecutive channels in fp64 (no atomics: contended fp64 shared-memory atomics cost more than the statistics pass saved)
    const double2* src = reinterpret_cast<const double2*>(chan) + static_cast<long long>(b) * C;
    double2* stage = reinterpret_cast<double2*>(sm);
    for (int c = threadIdx.x; c < C; c += blockDim.x) stage[c] = src[c];
    __syncthreads();
    if (threadIdx.x < groups) {
      double a0 = 0.0, a1 = 0.0;
      const double2* q = stage + threadIdx.x * gs;
#pragma unroll 4
      for (int i = 0; i < gs; ++i) { const double2 v = q[i]; a0 += v.x; a1 += v.y; }
      sgd[threadIdx.x * 2] = a0; sgd[threadIdx.x * 2 + 1] = a1;
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / gs;
    const double s0 = CHAN ? sgd[g * 2] : stats[(static_cast<long long>(b) * groups + g) * 2];
    const double s1 = CHAN ? sgd[g * 2 + 1] : stats[(static_cast<long long>(b) * groups + g) * 2 + 1];
    const double m = s0 / n;
    double var = s1 / n - m * m;
    if (var < 0.0) var = 0.0;
    const float rstd = rsqrtf(static_cast<float>(var) + eps);   // mean / variance in fp64, rstd in fp32 (2 ulp)
    const float a = rstd * __bfloat162float(gamma[c]);
    sm[c] = a;
    sm[C + c] = __bfloat162float(beta[c]) - static_cast<float>(m) * a;
  }
  __syncthreads();
  const int r = threadIdx.x / nvec, v = threadIdx.x % nvec;
  if (r >= R) return;
  const int p0 = blockIdx.x * chunk_pixels;
  int p1 = p0 + chunk_pixels; if (p1 > npix) p1 = npix;
  float2 a[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    a[j] = make_float2(sm[v * 8 + 2 * j], sm[v * 8 + 2 * j + 1]);
    sh[j] = make_float2(sm[C + v * 8 + 2 * j], sm[C + v * 8 + 2 * j + 1]);
  }
  bf16* const ob = out + static_cast<long long>(b) * npix * C + v * 8;
  auto emit = [&](const uint4& u, int pp) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = ffma2(unpack_bf16x2(w[j]), a[j], sh[j]);
      if (do_silu) f = silu2(f);
      o[j] = pack_bf16x2(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(ob + static_cast<long long>(pp) * C) = make_uint4(o[0], o[1], o[2], o[3]);
  };
  int p = p0 + r;
  for (; p + (kGnDepth - 1) * R < p1; p += kGnDepth * R) {
    uint4 u[kGnDepth];
#pragma unroll
    for (int k = 0; k < kGnDepth; ++k) u[k] = gn_load(s, b, p + k * R, v, Wv);
#pragma unroll
    for (int k = 0; k < kGnDepth; ++k) emit(u[k], p + k * R);
  }
  for (; p < p1; p += R) emit(gn_load(s, b, p, v, Wv), p);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: LPR lanes per row (8 / 16 / 32 -> 4 / 2 / 1 rows per warp), up to 5 16-byte vectors per lane,
// so C = 320 / 640 / 1280 keep every lane busy; the row lives in registers, gamma / beta in shared memory.
// ---------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(256, 2)
layernorm_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ y, long long ldy,
                 long long M, int C, const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                 float eps, const bf16* __restrict__ pre_add, int pre_div_a, int pre_mod_a,
                 int pre_mul_a, int pre_mod_b,
                 const float* __restrict__ post_add, int post_div, int post_mod) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  constexpr int VPL = 5, RPW = 32 / LPR, PD = 2;
  extern __shared__ uint4 ln_gb[];                 // [nvec] gamma vectors, then [nvec] beta vectors
  const int nvec = C >> 3;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    ln_gb[v] = *reinterpret_cast<const uint4*>(gamma + v * 8);
    ln_gb[nvec + v] = *reinterpret_cast<const uint4*>(beta + v * 8);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane / LPR, l = lane % LPR;
  const long long gstride = static_cast<long long>(gridDim.x) * (blockDim.x >> 5) * RPW;
  long long base = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
  const float invC = 1.0f / C;
  // PD row groups in flight per warp (register ring) to cover HBM latency
  uint4 ring[PD][VPL];
#pragma unroll
  for (int d = 0; d < PD; ++d) {
    const long long rr = base + d * gstride + sub;
    if (rr < M) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = l + i * LPR;
        if (v < nvec) ring[d][i] = *reinterpret_cast<const uint4*>(x + rr * ldx + v * 8);
      }
    }
  }
  for (; base < M; base += gstride) {              // warp-uniform loop: the shuffles below need every lane
    const long long row = base + sub;
    const bool valid = row < M;
    float2 f[VPL][4];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      f[i][0] = unpack_bf16x2(ring[0][i].x); f[i][1] = unpack_bf16x2(ring[0][i].y);
      f[i][2] = unpack_bf16x2(ring[0][i].z); f[i][3] = unpack_bf16x2(ring[0][i].w);
    }
#pragma unroll
    for (int d = 0; d + 1 < PD; ++d)
#pragma unroll
      for (int i = 0; i < VPL; ++i) ring[d][i] = ring[d + 1][i];
    const long long rn = base + static_cast<long long>(PD) * gstride + sub;   // refill PD groups ahead
    if (rn < M) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = l + i * LPR;
        if (v < nvec) ring[PD - 1][i] = *reinterpret_cast<const uint4*>(x + rn * ldx + v * 8);
      }
    }
    // pre-add table row = ((row / div_a) % mod_a) * mul_a + row % mod_b   (view-major PE of WarpAttn's pers tokens)
    float2 sum2 = make_float2(0.f, 0.f);
    if (pre_add != nullptr && valid) {
      const bf16* pr = pre_add + (((row / pre_div_a) % pre_mod_a) * pre_mul_a + row % pre_mod_b) * C;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = l + i * LPR;
        if (v < nvec) {
          const uint4 u = *reinterpret_cast<const uint4*>(pr + v * 8);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {            // bf16 add like the reference: round the sum to bf16
            const float2 t = fadd2(f[i][j], unpack_bf16x2(w[j]));
            f[i][j] = unpack_bf16x2(pack_bf16x2(t.x, t.y));
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      if (l + i * LPR < nvec) sum2 = fadd2(fadd2(f[i][0], f[i][1]), fadd2(fadd2(f[i][2], f[i][3]), sum2));
    float sum = sum2.x + sum2.y;
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * invC;
    const float2 nmean = make_float2(-mean, -mean);
    float2 var2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      if (l + i * LPR < nvec) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { f[i][j] = fadd2(f[i][j], nmean); var2 = ffma2(f[i][j], f[i][j], var2); }
      }
    float var = var2.x + var2.y;
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    if (!valid) continue;
    const float rstd = rsqrtf(var * invC + eps);
    const float2 rstd2 = make_float2(rstd, rstd);
    const float* po = post_add ? post_add + static_cast<long long>((row / post_div) % post_mod) * C : nullptr;
    bf16* yr = y + row * ldy;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = l + i * LPR;
      if (v < nvec) {
        const uint4 gq = ln_gb[v], bq = ln_gb[nvec + v];
        const uint32_t gw[4] = {gq.x, gq.y, gq.z, gq.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w};
        uint32_t outw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 o = ffma2(f[i][j], fmul2(unpack_bf16x2(gw[j]), rstd2), unpack_bf16x2(bw[j]));
          if (po) {
            const float2 r = unpack_bf16x2(pack_bf16x2(o.x, o.y));
            o = fadd2(r, make_float2(po[v * 8 + 2 * j], po[v * 8 + 2 * j + 1]));
          }
          outw[j] = pack_bf16x2(o.x, o.y);
        }
        *reinterpret_cast<uint4*>(yr + v * 8) = make_uint4(outw[0], outw[1], outw[2], outw[3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// row softmax (fp32 math on bf16 logits) -- the VAE's single-head AttentionBlock computes
// softmax(scores.float()).type(bf16) on materialised scores (diffusers/models/attention.py:353-367)
// ---------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long ld, int N) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  __shared__ float red[32];
  const bf16* xr = x + static_cast<long long>(blockIdx.x) * ld;
  bf16* yr = y + static_cast<long long>(blockIdx.x) * ld;
  const int nvec = N >> 3;
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float f[8]; unpack8(*reinterpret_cast<const uint4*>(xr + v * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) mx = fmaxf(mx, f[j]);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float f[8]; unpack8(*reinterpret_cast<const uint4*>(xr + v * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += __expf(f[j] - mx);
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) sum += red[w];
  const float inv = 1.0f / sum;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float f[8]; unpack8(*reinterpret_cast<const uint4*>(xr + v * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __expf(f[j] - mx) * inv;
    *reinterpret_cast<uint4*>(yr + v * 8) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                       pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
}

}  // namespace i360

using namespace i360;

extern "C" int i360_softmax_rows_bf16(const void* x, void* y, long long ld, long long M, int N, void* stream) {
  if (!x || !y || M <= 0 || N <= 0 || (N % 8) || (ld % 8)) return I360_ERR_ARG;
  launch_k(softmax_rows_kernel, dim3(static_cast<unsigned>(M)), dim3(256), 0, static_cast<cudaStream_t>(stream),
           static_cast<const bf16*>(x), static_cast<bf16*>(y), ld, N);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}

static int gn_geometry(int C, int B, long long npix, int* block, int* chunk, int* chunks) {
  const int nvec = C / 8;
  if (nvec <= 0 || nvec > 1024) return I360_ERR_ARG;
  int R = 384 / nvec; if (R < 1) R = 1;
  *block = R * nvec;
  // CTAs per image: ~8 CTAs per SM in total, so the 3 resident CTAs per SM turn over a few times and the last wave's
  // imbalance stays small (one CTA per image left 640 CTAs on 148 SMs: 5 vs 4 per SM, 14 % idle).  Swept on one box
  // (I360_GN_CTAS_PER_SM, GroupNorm+SiLU stats + apply): 12 -> 8 = 0.259 -> 0.251 ms (640x32x32x320), 0.126 -> 0.120
  // (32x64x128x320), 0.156 -> 0.142 (640x16x16x640), 0.116 -> 0.093 (640x8x8x1280); 3-4 and >= 18 are slower.
  static int per_sm = 0;
  if (per_sm == 0) { const char* e = getenv("I360_GN_CTAS_PER_SM"); per_sm = e ? atoi(e) : 8; if (per_sm < 1) per_sm = 8; }
  long long target = (static_cast<long long>(num_sms()) * per_sm + B - 1) / B;
  if (target < 1) target = 1;
  long long cp = (npix + target - 1) / target;
  const long long minp = static_cast<long long>(R) * kGnDepth * 2;
  if (cp < minp) cp = minp;
  *chunk = static_cast<int>(cp);
  *chunks = static_cast<int>((npix + cp - 1) / cp);
  return I360_OK;
}

// stats: [B, groups, 2] fp64, zeroed here.  x2/C2 optional second concat source.
extern "C" int i360_groupnorm_stats(const void* x1, int C1, const void* x2, int C2, int B, int H, int Wsrc,
                                    int pad, int groups, double* stats, void* stream) {
  const int C = C1 + C2;
  if (!x1 || !stats || (C1 % 8) || (C2 % 8) || C <= 0 || (C % groups) || B <= 0 || pad < 0 || pad > Wsrc)
    return I360_ERR_ARG;
  if (C2 > 0 && !x2) return I360_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * groups, st) != cudaSuccess) return I360_ERR_CUDA;
  GnSrc s{static_cast<const bf16*>(x1), static_cast<const bf16*>(x2), C1, C2, H, Wsrc, pad};
  int block, chunk, chunks;
  const long long npix = static_cast<long long>(H) * (Wsrc + 2 * pad);
  int r = gn_geometry(C, B, npix, &block, &chunk, &chunks); if (r) return r;
  const size_t smem = static_cast<size_t>(block) * 8 * 2 * sizeof(float);
  if (smem > 48 * 1024) return I360_ERR_ARG;
  launch_k(gn_stats_kernel, dim3(chunks, B), dim3(block), smem, st, s, groups, chunk, stats);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}

// out: dense [B, H, Wsrc+2*pad, C1+C2].  stats_count: elements per (image, group) the statistics were taken
// over; pass 0 when they were taken over this same virtual tensor (conv_norm_out normalises BEFORE the
// circular pad, MVGenModel.py:472-475, so there stats use pad 0 while the output is written with pad 1).
extern "C" int i360_groupnorm_apply(const void* x1, int C1, const void* x2, int C2, int B, int H, int Wsrc,
                                    int pad, int groups, const double* stats, double stats_count, const void* gamma,
                                    const void* beta, float eps, int do_silu, void* out, void* stream) {
  const int C = C1 + C2;
  if (!x1 || !stats || !gamma || !beta || !out || (C1 % 8) || (C2 % 8) || (C % groups)) return I360_ERR_ARG;
  GnSrc s{static_cast<const bf16*>(x1), static_cast<const bf16*>(x2), C1, C2, H, Wsrc, pad};
  int block, chunk, chunks;
  const long long npix = static_cast<long long>(H) * (Wsrc + 2 * pad);
  int r = gn_geometry(C, B, npix, &block, &chunk, &chunks); if (r) return r;
  const size_t smem = static_cast<size_t>(C) * 2 * sizeof(float);
  launch_k(gn_apply_kernel<false>, dim3(chunks, B), dim3(block), smem, static_cast<cudaStream_t>(stream),
      s, groups, chunk, stats, static_cast<const bf16*>(gamma), static_cast<const bf16*>(beta), eps, do_silu,
      stats_count, static_cast<bf16*>(out), static_cast<const double*>(nullptr));
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}

// GroupNorm (+SiLU) of a dense NHWC tensor x [B, H, W, C] whose PER-CHANNEL statistics the producing conv's epilogue already
// wrote (i360_conv3x3_chanstats_bf16): chan_stats [B, C, 2] fp64 (sum, sum of squares over the image); channels are folded
// into the `groups` groups here, so no statistics pass over x runs.  out: dense [B, H, W, C].
extern "C" int i360_groupnorm_apply_chanstats(const void* x, int C, int B, int H, int W, int groups, const double* chan_stats,
                                              const void* gamma, const void* beta, float eps, int do_silu, void* out,
                                              void* stream) {
  if (!x || !chan_stats || !gamma || !beta || !out || (C % 8) || groups <= 0 || groups > 64 || (C % groups)) return I360_ERR_ARG;
  GnSrc s{static_cast<const bf16*>(x), nullptr, C, 0, H, W, 0};
  int block, chunk, chunks;
  const long long npix = static_cast<long long>(H) * W;
  int r = gn_geometry(C, B, npix, &block, &chunk, &chunks); if (r) return r;
  const size_t smem = static_cast<size_t>(C) * 2 * sizeof(double);     // the channel sums are staged where scale / shift go later
  launch_k(gn_apply_kernel<true>, dim3(chunks, B), dim3(block), smem, static_cast<cudaStream_t>(stream),
           s, groups, chunk, static_cast<const double*>(nullptr), static_cast<const bf16*>(gamma), static_cast<const bf16*>(beta), eps,
           do_silu, 0.0, static_cast<bf16*>(out), chan_stats);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}

extern "C" int i360_layernorm(const void* x, long long ldx, void* y, long long ldy, long long M, int C,
                              const void* gamma, const void* beta, float eps, const void* pre_add, int pre_div_a,
                              int pre_mod_a, int pre_mul_a, int pre_mod_b, const float* post_add, int post_div,
                              int post_mod, void* stream) {
  if (!x || !y || !gamma || !beta || M <= 0 || C <= 0 || (C % 8) || (ldx % 8) || (ldy % 8)) return I360_ERR_ARG;
  if (pre_add && (pre_div_a <= 0 || pre_mod_a <= 0 || pre_mod_b <= 0)) return I360_ERR_ARG;
  if (post_add && (post_div <= 0 || post_mod <= 0)) return I360_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nvec = C / 8, warps = 8;
  const int lpr = nvec <= 8 * 5 ? 8 : (nvec <= 16 * 5 ? 16 : 32);
  if (nvec > 32 * 5) return I360_ERR_UNSUPPORTED;
  const int rows_per_block = warps * (32 / lpr);
  long long blocks = (M + rows_per_block - 1) / rows_per_block;
  const long long cap = static_cast<long long>(num_sms()) * 2;     // persistent: 2 resident blocks per SM loop over rows
  if (blocks > cap) blocks = cap;
  const unsigned grid = static_cast<unsigned>(blocks);
  const bf16 *xx = static_cast<const bf16*>(x), *g = static_cast<const bf16*>(gamma), *b = static_cast<const bf16*>(beta);
  const bf16* pa = static_cast<const bf16*>(pre_add);
  bf16* yy = static_cast<bf16*>(y);
  const size_t smem = static_cast<size_t>(nvec) * 2 * sizeof(uint4);
#define I360_LN_LAUNCH(L) launch_k(layernorm_kernel<L>, dim3(grid), dim3(warps * 32), smem, st, xx, ldx, yy, ldy, M, C, g, b, eps, pa, pre_div_a, pre_mod_a, pre_mul_a, pre_mod_b, post_add, post_div, post_mod)
  if (lpr == 8) I360_LN_LAUNCH(8);
  else if (lpr == 16) I360_LN_LAUNCH(16);
  else I360_LN_LAUNCH(32);
#undef I360_LN_LAUNCH
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}
