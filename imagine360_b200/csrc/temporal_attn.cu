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
#include <stdlib.h>

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
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
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

// ------------------------------------------------------------------------------------------------------
// F <= 16: tensor-core version.  One warp per (pixel, head): S = Q K^T and O = P V are m16n8k16 bf16 MMAs
// (mma.sync -- the problem is a 16x16 tile, far below a tcgen05 tile); the softmaxed S accumulators are re-used
// in place as the A fragments of P.V, V is read through ldmatrix.trans, and the next task's q/k/v are prefetched
// with cp.async into the second smem stage while the current one is computed.  ~130 instructions per task
// instead of ~3000 scalar ones, which turns this kernel from issue-bound into HBM-bound.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_u32(smem_row)));
}

// STAGES smem stages per warp: the tasks of the next STAGES-1 iterations are in flight while one is computed.  The kernel
// is a pure HBM stream with ~4 KB per task, so bytes in flight per SM (Little's law against ~2 us of loaded-HBM
// latency) set its bandwidth: 16 warps x 1 task ahead = 61 KB gave 3.9 TB/s; 3 stages put 2 tasks per warp in flight.
template <int STAGES>
__global__ void temporal_attn_mma_kernel(const TAParams p, int hd_pad, int pitch) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  extern __shared__ __align__(16) uint8_t smem_ta[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int stage_elems = 3 * 16 * pitch;
  bf16* base = reinterpret_cast<bf16*>(smem_ta) + static_cast<size_t>(warp) * STAGES * stage_elems;
  for (int i = lane; i < STAGES * stage_elems / 8; i += 32) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  const long long total = static_cast<long long>(p.B) * p.D * p.heads;
  const long long tstride = static_cast<long long>(gridDim.x) * nwarps;
  const int cpr = p.hd >> 3, nchunks = p.F * cpr;
  const int g = lane >> 2, q4 = lane & 3;

  // the (frame, 16-byte chunk) pairs this lane copies are the same for every task: the runtime divisions by the
  // chunks-per-row count are done once here instead of ~6 times per task
  constexpr int kMaxCopies = 10;                         // 16 frames x head_dim 160 / 8 / 32 lanes
  int cp_smem[kMaxCopies]; int cp_frame[kMaxCopies]; int cp_col[kMaxCopies];
  int n_copies = 0;
#pragma unroll
  for (int k = 0; k < kMaxCopies; ++k) {
    const int c = lane + 32 * k;
    const int f = c / cpr, ch = c - f * cpr;
    cp_frame[k] = f; cp_col[k] = ch * 8; cp_smem[k] = f * pitch + ch * 8;
    if (c < nchunks) n_copies = k + 1;
  }
  auto issue = [&](long long task, int stage) {
    const int head = static_cast<int>(task % p.heads);
    const long long pix = task / p.heads;
    const long long row0 = (static_cast<long long>(pix / p.D) * p.F) * p.D + pix % p.D;
    bf16* sq = base + stage * stage_elems;
#pragma unroll
    for (int k = 0; k < kMaxCopies; ++k) {
      if (k < n_copies) {
        const long long r = row0 + static_cast<long long>(cp_frame[k]) * p.D;
        const int col = head * p.hd + cp_col[k];
        cp_async16(sq + cp_smem[k], p.q + r * p.ldq + col);
        cp_async16(sq + 16 * pitch + cp_smem[k], p.k + r * p.ldk + col);
        cp_async16(sq + 32 * pitch + cp_smem[k], p.v + r * p.ldv + col);
      }
    }
  };

  long long task = static_cast<long long>(blockIdx.x) * nwarps + warp;
  int stage = 0;
  // prologue: STAGES-1 tasks in flight; every iteration commits exactly one (possibly empty) group, so that
  // "all but the newest STAGES-1 groups are complete" always means "the current task's tile has landed"
#pragma unroll
  for (int k = 0; k < STAGES - 1; ++k) {
    if (task + k * tstride < total) issue(task + k * tstride, k);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (; task < total; task += tstride, stage = (stage + 1 == STAGES) ? 0 : stage + 1) {
    const long long nxt = task + (STAGES - 1) * tstride;
    if (nxt < total) issue(nxt, (stage + STAGES - 1) % STAGES);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
    __syncwarp();
    bf16* sQ = base + stage * stage_elems;
    const bf16* sK = sQ + 16 * pitch;
    const bf16* sV = sQ + 32 * pitch;
    // ---- S = Q K^T (16 x 16), two 8-key n-tiles ----
    float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int k0 = 0; k0 < hd_pad; k0 += 16) {
      uint32_t a[4];
      a[0] = *reinterpret_cast<const uint32_t*>(sQ + g * pitch + k0 + q4 * 2);
      a[1] = *reinterpret_cast<const uint32_t*>(sQ + (g + 8) * pitch + k0 + q4 * 2);
      a[2] = *reinterpret_cast<const uint32_t*>(sQ + g * pitch + k0 + 8 + q4 * 2);
      a[3] = *reinterpret_cast<const uint32_t*>(sQ + (g + 8) * pitch + k0 + 8 + q4 * 2);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(sK + (nt * 8 + g) * pitch + k0 + q4 * 2);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(sK + (nt * 8 + g) * pitch + k0 + 8 + q4 * 2);
        mma_bf16_16816(sc[nt], a, b0, b1);
      }
    }
    // ---- softmax over keys (row g: sc[.][0..1], row g+8: sc[.][2..3]; a row is spread over the 4 lanes of a quad) ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = (nt * 8 + q4 * 2 + e) < p.F;
        sc[nt][e] = ok ? sc[nt][e] * p.scale : -INFINITY;
        sc[nt][2 + e] = ok ? sc[nt][2 + e] * p.scale : -INFINITY;
        mx0 = fmaxf(mx0, sc[nt][e]); mx1 = fmaxf(mx1, sc[nt][2 + e]);
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sc[nt][e] = __expf(sc[nt][e] - mx0); sum0 += sc[nt][e];
        sc[nt][2 + e] = __expf(sc[nt][2 + e] - mx1); sum1 += sc[nt][2 + e];
      }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    uint32_t pa[4] = {pack_bf16x2(sc[0][0], sc[0][1]), pack_bf16x2(sc[0][2], sc[0][3]),
                      pack_bf16x2(sc[1][0], sc[1][1]), pack_bf16x2(sc[1][2], sc[1][3])};
    __syncwarp();      // all lanes are done reading Q: its tile becomes the output staging area
    // ---- O = P V, 8 head-dim columns per MMA ----
    for (int n0 = 0; n0 < p.hd; n0 += 8) {
      uint32_t b0, b1;
      ldmatrix_x2_trans(b0, b1, sV + (lane & 15) * pitch + n0);
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16_16816(o, pa, b0, b1);
      *reinterpret_cast<uint32_t*>(sQ + g * pitch + n0 + q4 * 2) = pack_bf16x2(o[0] * inv0, o[1] * inv0);
      *reinterpret_cast<uint32_t*>(sQ + (g + 8) * pitch + n0 + q4 * 2) = pack_bf16x2(o[2] * inv1, o[3] * inv1);
    }
    __syncwarp();
    {
      const int head = static_cast<int>(task % p.heads);
      const long long pix = task / p.heads;
      const long long row0 = (static_cast<long long>(pix / p.D) * p.F) * p.D + pix % p.D;
#pragma unroll
      for (int k = 0; k < kMaxCopies; ++k) {
        if (k < n_copies)
          *reinterpret_cast<uint4*>(p.o + (row0 + static_cast<long long>(cp_frame[k]) * p.D) * p.ldo + head * p.hd + cp_col[k]) =
              *reinterpret_cast<const uint4*>(sQ + cp_smem[k]);
      }
    }
    // rows >= F of the staging tile were written with garbage-free zeros only if F == 16; restore the zero padding
    if (p.F < 16) {
      for (int c = lane; c < (16 - p.F) * (pitch / 8); c += 32)
        reinterpret_cast<uint4*>(sQ + p.F * pitch)[c] = make_uint4(0, 0, 0, 0);
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
    const int hd_pad = (head_dim + 15) / 16 * 16;
    const int pitch = hd_pad + 8;                       // +16 bytes per row: conflict-free fragment / ldmatrix reads
    // Warps per SM are what buys bandwidth here (each warp keeps one ~4 KB task in flight while it computes another;
    // measured: 16 -> 20 warps per SM = -7.5 % at head_dim 40, 8 -> 10 = -15 % at head_dim 80), so the launch takes as
    // many warps as ~216 KB of smem and 80 registers per thread allow, in one or two CTAs per SM.  A third stage at the
    // cost of warps (12 per SM) was slower (I360_TA_STAGES3 keeps it for experiments).
    const size_t per_stage = static_cast<size_t>(3) * 16 * pitch * sizeof(bf16);
    const bool three = getenv("I360_TA_STAGES3") != nullptr && per_stage * 3 * 6 <= 110 * 1024;
    const size_t pw = per_stage * (three ? 3 : 2);
    int total_w = static_cast<int>((216 * 1024) / pw);
    if (total_w > 24) total_w = 24;
    if (total_w < 1) return I360_ERR_UNSUPPORTED;
    int ctas = total_w > 12 ? 2 : 1;
    int w2 = total_w / ctas;
    if (const char* e = getenv("I360_TA_WARPS")) { w2 = atoi(e); ctas = (pw * w2 * 2 <= 220 * 1024) ? 2 : 1; }
    if (three) { w2 = 6; ctas = 2; }
    const size_t sm2 = pw * w2;
    if (sm2 > 227 * 1024 || w2 < 1 || w2 > 32) return I360_ERR_UNSUPPORTED;
    long long bl = (total + w2 - 1) / w2;
    const long long cap2 = static_cast<long long>(num_sms()) * ctas;
    if (bl > cap2) bl = cap2;
    static bool setm = false;
    if (!setm) {
      cudaFuncSetAttribute(temporal_attn_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(temporal_attn_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      setm = true;
    }
    if (three) launch_k(temporal_attn_mma_kernel<3>, dim3(static_cast<unsigned>(bl)), dim3(w2 * 32), sm2, st, p, hd_pad, pitch);
    else       launch_k(temporal_attn_mma_kernel<2>, dim3(static_cast<unsigned>(bl)), dim3(w2 * 32), sm2, st, p, hd_pad, pitch);
    I360_CUDA_CHECK_LAUNCH();
    return I360_OK;
  }
  if (F <= 16) {
    static bool set0 = false;
    if (!set0) { cudaFuncSetAttribute(temporal_attn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set0 = true; }
    launch_k(temporal_attn_kernel<0>, dim3(static_cast<unsigned>(blocks)), dim3(warps * 32), smem, st, p);
  } else {
    static bool set1 = false;
    if (!set1) { cudaFuncSetAttribute(temporal_attn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set1 = true; }
    launch_k(temporal_attn_kernel<1>, dim3(static_cast<unsigned>(blocks)), dim3(warps * 32), smem, st, p);
  }
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}
