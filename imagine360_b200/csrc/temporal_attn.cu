// Temporal self-attention over the frame axis (sequence length F = 8..32) for every (clip, pixel, head).
//
// Replaces VersatileAttention's core (animatediff/models/motion_module.py:343-429: rearrange
// "(b f) d c -> (b d) f c", baddbmm -> softmax -> bmm) and the legacy CrossAttention core inside
// TemporalProjection (animatediff/models/resampler.py:231-267).  The work is tiny per problem
// (F x F x head_dim) and there are ~10^5..10^6 problems per call, so this is an HBM-bound CUDA-core
// kernel: q/k/v are read IN PLACE from the token-major [(b f d), 3C] projection output (no transpose
// copies), one warp per (pixel, head), fp32 math, one coalesced pass in and out.
#include "common.cuh"
#include "tmap.h"

namespace i360 {

struct TAParams {
  const bf16* q; const bf16* k; const bf16* v; bf16* o;
  long long ldq, ldk, ldv, ldo;   // row strides (elements)
  int B, F, D, heads, hd;
  float scale;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// MODE 0: F <= 16, two lanes per query row (8 keys each).  MODE 1: F <= 32, one lane per row.
template <int MODE>
__global__ void temporal_attn_kernel(const TAParams p) {
  extern __shared__ __align__(16) uint8_t smem_ta[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int tile_elems = p.F * p.hd;                 // per tensor
  bf16* sQ = reinterpret_cast<bf16*>(smem_ta) + static_cast<size_t>(warp) * 3 * tile_elems;
  bf16* sK = sQ + tile_elems;
  bf16* sV = sK + tile_elems;
  const long long total = static_cast<long long>(p.B) * p.D * p.heads;
  const int chunks_per_row = p.hd >> 3;
  const int nchunks = p.F * chunks_per_row;
  constexpr int KEYS = MODE == 0 ? 8 : 32;

  for (long long task = static_cast<long long>(blockIdx.x) * nwarps + warp; task < total;
       task += static_cast<long long>(gridDim.x) * nwarps) {
    const int head = static_cast<int>(task % p.heads);
    const long long pix = task / p.heads;
    const int d = static_cast<int>(pix % p.D);
    const int b = static_cast<int>(pix / p.D);
    const long long row0 = (static_cast<long long>(b) * p.F) * p.D + d;   // frame f -> row0 + f*D
    // ---- stage q, k, v [F, hd] ----
    // all 3*F*hd/8 16-byte copies of this task go out as asynchronous copies before a single wait
    for (int c = lane; c < nchunks; c += 32) {
      const int f = c / chunks_per_row, ch = c % chunks_per_row;
      const long long r = row0 + static_cast<long long>(f) * p.D;
      const int col = head * p.hd + ch * 8;
      cp_async16(reinterpret_cast<uint4*>(sQ) + c, p.q + r * p.ldq + col);
      cp_async16(reinterpret_cast<uint4*>(sK) + c, p.k + r * p.ldk + col);
      cp_async16(reinterpret_cast<uint4*>(sV) + c, p.v + r * p.ldv + col);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    const int i = MODE == 0 ? (lane >> 1) : lane;          // my query frame
    const int j0 = MODE == 0 ? (lane & 1) * 8 : 0;         // my first key
    const bool row_ok = i < p.F;
    float s[KEYS];
#pragma unroll
    for (int t = 0; t < KEYS; ++t) s[t] = 0.f;
    if (row_ok) {
      for (int ch = 0; ch < chunks_per_row; ++ch) {
        float qf[8];
        {
          const uint4 u = reinterpret_cast<const uint4*>(sQ)[i * chunks_per_row + ch];
          float2 a = unpack_bf16x2(u.x), bq = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), dd = unpack_bf16x2(u.w);
          qf[0] = a.x; qf[1] = a.y; qf[2] = bq.x; qf[3] = bq.y; qf[4] = c.x; qf[5] = c.y; qf[6] = dd.x; qf[7] = dd.y;
        }
#pragma unroll
        for (int t = 0; t < KEYS; ++t) {
          const int j = j0 + t;
          if (j < p.F) {
            const uint4 u = reinterpret_cast<const uint4*>(sK)[j * chunks_per_row + ch];
            float2 a = unpack_bf16x2(u.x), bk = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), dd = unpack_bf16x2(u.w);
            s[t] += qf[0] * a.x + qf[1] * a.y + qf[2] * bk.x + qf[3] * bk.y + qf[4] * c.x + qf[5] * c.y +
                    qf[6] * dd.x + qf[7] * dd.y;
          }
        }
      }
    }
    // ---- softmax over keys ----
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < KEYS; ++t) {
      s[t] = (j0 + t < p.F) ? s[t] * p.scale : -INFINITY;
      mx = fmaxf(mx, s[t]);
    }
    if (MODE == 0) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < KEYS; ++t) { s[t] = (j0 + t < p.F) ? __expf(s[t] - mx) : 0.f; sum += s[t]; }
    if (MODE == 0) sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    const float inv = row_ok ? 1.0f / sum : 0.f;
    __syncwarp();   // everyone finished reading sQ -> reuse it for the output tile
    // ---- O[i, :] = sum_j p[i,j] V[j, :] ----
    for (int ch = 0; ch < chunks_per_row; ++ch) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = 0.f;
      if (row_ok) {
#pragma unroll
        for (int t = 0; t < KEYS; ++t) {
          const int j = j0 + t;
          if (j < p.F) {
            const uint4 u = reinterpret_cast<const uint4*>(sV)[j * chunks_per_row + ch];
            float2 a = unpack_bf16x2(u.x), bv = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), dd = unpack_bf16x2(u.w);
            const float w = s[t];
            o[0] += w * a.x; o[1] += w * a.y; o[2] += w * bv.x; o[3] += w * bv.y;
            o[4] += w * c.x; o[5] += w * c.y; o[6] += w * dd.x; o[7] += w * dd.y;
          }
        }
      }
      if (MODE == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] += __shfl_xor_sync(0xffffffffu, o[e], 1);
      }
      if (row_ok && (MODE == 1 || (lane & 1) == 0)) {
        reinterpret_cast<uint4*>(sQ)[i * chunks_per_row + ch] =
            make_uint4(pack_bf16x2(o[0] * inv, o[1] * inv), pack_bf16x2(o[2] * inv, o[3] * inv),
                       pack_bf16x2(o[4] * inv, o[5] * inv), pack_bf16x2(o[6] * inv, o[7] * inv));
      }
    }
    __syncwarp();
    for (int c = lane; c < nchunks; c += 32) {
      const int f = c / chunks_per_row, ch = c % chunks_per_row;
      const long long r = row0 + static_cast<long long>(f) * p.D;
      *reinterpret_cast<uint4*>(p.o + r * p.ldo + head * p.hd + ch * 8) = reinterpret_cast<const uint4*>(sQ)[c];
    }
    __syncwarp();
  }
}

}  // namespace i360

using namespace i360;

// q/k/v/o rows are tokens ordered (b, f, d); head h occupies columns [h*hd, (h+1)*hd) of each pointer.
extern "C" int i360_temporal_attention_bf16(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                            long long ldv, void* o, long long ldo, int B, int F, int D, int heads,
                                            int head_dim, float scale, void* stream) {
  if (!q || !k || !v || !o || B <= 0 || F <= 0 || D <= 0 || heads <= 0) return I360_ERR_ARG;
  if (F > 32) return I360_ERR_UNSUPPORTED;
  if ((head_dim % 8) || (ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return I360_ERR_ARG;
  TAParams p{static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v),
             static_cast<bf16*>(o), ldq, ldk, ldv, ldo, B, F, D, heads, head_dim, scale};
  const size_t per_warp = static_cast<size_t>(3) * F * head_dim * sizeof(bf16);
  int warps = 8;
  while (warps > 1 && per_warp * warps > 100 * 1024) warps >>= 1;
  const size_t smem = per_warp * warps;
  if (smem > 200 * 1024) return I360_ERR_UNSUPPORTED;
  const long long total = static_cast<long long>(B) * D * heads;
  long long blocks = (total + warps - 1) / warps;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (F <= 16) {
    static bool set0 = false;
    if (!set0) { cudaFuncSetAttribute(temporal_attn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set0 = true; }
    temporal_attn_kernel<0><<<static_cast<unsigned>(blocks), warps * 32, smem, st>>>(p);
  } else {
    static bool set1 = false;
    if (!set1) { cudaFuncSetAttribute(temporal_attn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set1 = true; }
    temporal_attn_kernel<1><<<static_cast<unsigned>(blocks), warps * 32, smem, st>>>(p);
  }
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}
