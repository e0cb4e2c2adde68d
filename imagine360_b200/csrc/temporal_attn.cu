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
// MT = 16-row tiles along the frame axis: 1 for F <= 16, 2 for F <= 32 (the 24-frame configuration; before, F > 16 took the
// scalar kernel at 0.8 TB/s -- 7.1 ms instead of ~1.4 ms for the 40 x 24 x 2304 x 320 level of the 24x768x1536 step).
template <int STAGES, int MT>
__global__ void __launch_bounds__(384) temporal_attn_mma_kernel(const TAParams p, int hd_pad, int pitch) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  constexpr int R = 16 * MT;                       // rows (frames, zero padded) of the q / k / v tiles
  extern __shared__ __align__(16) uint8_t smem_ta[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int stage_elems = 3 * R * pitch;
  bf16* base = reinterpret_cast<bf16*>(smem_ta) + static_cast<size_t>(warp) * STAGES * stage_elems;
  for (int i = lane; i < STAGES * stage_elems / 8; i += 32) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  const long long total = static_cast<long long>(p.B) * p.D * p.heads;
  const long long tstride = static_cast<long long>(gridDim.x) * nwarps;
  const int cpr = p.hd >> 3, nchunks = p.F * cpr;
  const int g = lane >> 2, q4 = lane & 3;

  // the (frame, 16-byte chunk) pairs this lane copies are the same for every task: the runtime divisions by the
  // chunks-per-row count are done once here instead of ~6 times per task
  constexpr int kMaxCopies = 10 * MT;                    // 16 * MT frames x head_dim 160 / 8 / 32 lanes
  // MT = 1 keeps the three indices in separate registers (80 registers in all); MT = 2 packs them into one word per copy
  // -- frame (8 bits) | column (8 bits) | smem element offset (16 bits) -- to stay within 170 registers at 12 warps per CTA
  constexpr int kIdxWords = (MT == 1) ? 3 : 1;
  int cp_idx[kMaxCopies][kIdxWords];
  int n_copies = 0;
#pragma unroll
  for (int k = 0; k < kMaxCopies; ++k) {
    const int c = lane + 32 * k;
    const int f = c / cpr, ch = c - f * cpr;
    if (MT == 1) { cp_idx[k][0] = f; cp_idx[k][kIdxWords > 1 ? 1 : 0] = ch * 8; cp_idx[k][kIdxWords > 2 ? 2 : 0] = f * pitch + ch * 8; }
    else cp_idx[k][0] = (f << 24) | ((ch * 8) << 16) | (f * pitch + ch * 8);
    if (c < nchunks) n_copies = k + 1;
  }
#define CP_FRAME(k) (MT == 1 ? cp_idx[k][0] : (cp_idx[k][0] >> 24))
#define CP_COL(k) (MT == 1 ? cp_idx[k][kIdxWords > 1 ? 1 : 0] : ((cp_idx[k][0] >> 16) & 0xff))
#define CP_SMEM(k) (MT == 1 ? cp_idx[k][kIdxWords > 2 ? 2 : 0] : (cp_idx[k][0] & 0xffff))
  auto issue = [&](long long task, int stage) {
    const int head = static_cast<int>(task % p.heads);
    const long long pix = task / p.heads;
    const long long row0 = (static_cast<long long>(pix / p.D) * p.F) * p.D + pix % p.D;
    bf16* sq = base + stage * stage_elems;
#pragma unroll
    for (int k = 0; k < kMaxCopies; ++k) {
      if (k < n_copies) {
        const long long r = row0 + static_cast<long long>(CP_FRAME(k)) * p.D;
        const int col = head * p.hd + CP_COL(k);
        cp_async16(sq + CP_SMEM(k), p.q + r * p.ldq + col);
        cp_async16(sq + R * pitch + CP_SMEM(k), p.k + r * p.ldk + col);
        cp_async16(sq + 2 * R * pitch + CP_SMEM(k), p.v + r * p.ldv + col);
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
    const bf16* sK = sQ + R * pitch;
    const bf16* sV = sQ + 2 * R * pitch;
    // ---- S = Q K^T (R x R): MT row tiles x 2 MT 8-key column tiles ----
    float sc[MT][2 * MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2 * MT; ++nt) { sc[mt][nt][0] = 0.f; sc[mt][nt][1] = 0.f; sc[mt][nt][2] = 0.f; sc[mt][nt][3] = 0.f; }
    for (int k0 = 0; k0 < hd_pad; k0 += 16) {
      uint32_t bk[2 * MT][2];
#pragma unroll
      for (int nt = 0; nt < 2 * MT; ++nt) {
        bk[nt][0] = *reinterpret_cast<const uint32_t*>(sK + (nt * 8 + g) * pitch + k0 + q4 * 2);
        bk[nt][1] = *reinterpret_cast<const uint32_t*>(sK + (nt * 8 + g) * pitch + k0 + 8 + q4 * 2);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        uint32_t a[4];
        a[0] = *reinterpret_cast<const uint32_t*>(sQ + (mt * 16 + g) * pitch + k0 + q4 * 2);
        a[1] = *reinterpret_cast<const uint32_t*>(sQ + (mt * 16 + g + 8) * pitch + k0 + q4 * 2);
        a[2] = *reinterpret_cast<const uint32_t*>(sQ + (mt * 16 + g) * pitch + k0 + 8 + q4 * 2);
        a[3] = *reinterpret_cast<const uint32_t*>(sQ + (mt * 16 + g + 8) * pitch + k0 + 8 + q4 * 2);
#pragma unroll
        for (int nt = 0; nt < 2 * MT; ++nt) mma_bf16_16816(sc[mt][nt], a, bk[nt][0], bk[nt][1]);
      }
    }
    // ---- softmax over keys (tile mt: row g in sc[mt][.][0..1], row g+8 in sc[mt][.][2..3]; a row is spread over the 4
    //      lanes of a quad) and the A fragments of P ----
    uint32_t pa[MT][MT][4];
    float inv[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 2 * MT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool ok = (nt * 8 + q4 * 2 + e) < p.F;
          sc[mt][nt][e] = ok ? sc[mt][nt][e] * p.scale : -INFINITY;
          sc[mt][nt][2 + e] = ok ? sc[mt][nt][2 + e] * p.scale : -INFINITY;
          mx0 = fmaxf(mx0, sc[mt][nt][e]); mx1 = fmaxf(mx1, sc[mt][nt][2 + e]);
        }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 2 * MT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          sc[mt][nt][e] = __expf(sc[mt][nt][e] - mx0); sum0 += sc[mt][nt][e];
          sc[mt][nt][2 + e] = __expf(sc[mt][nt][2 + e] - mx1); sum1 += sc[mt][nt][2 + e];
        }
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
      inv[mt][0] = 1.0f / sum0; inv[mt][1] = 1.0f / sum1;
#pragma unroll
      for (int ks = 0; ks < MT; ++ks) {              // 16 keys per k-step = column tiles 2 ks, 2 ks + 1
        pa[mt][ks][0] = pack_bf16x2(sc[mt][2 * ks][0], sc[mt][2 * ks][1]);
        pa[mt][ks][1] = pack_bf16x2(sc[mt][2 * ks][2], sc[mt][2 * ks][3]);
        pa[mt][ks][2] = pack_bf16x2(sc[mt][2 * ks + 1][0], sc[mt][2 * ks + 1][1]);
        pa[mt][ks][3] = pack_bf16x2(sc[mt][2 * ks + 1][2], sc[mt][2 * ks + 1][3]);
      }
    }
    // ---- O = P V, 8 head-dim columns per MMA ----
    for (int n0 = 0; n0 < p.hd; n0 += 8) {
      uint32_t bv[MT][2];
#pragma unroll
      for (int ks = 0; ks < MT; ++ks) ldmatrix_x2_trans(bv[ks][0], bv[ks][1], sV + (ks * 16 + (lane & 15)) * pitch + n0);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < MT; ++ks) mma_bf16_16816(o, pa[mt][ks], bv[ks][0], bv[ks][1]);
        *reinterpret_cast<uint32_t*>(sQ + (mt * 16 + g) * pitch + n0 + q4 * 2) = pack_bf16x2(o[0] * inv[mt][0], o[1] * inv[mt][0]);
        *reinterpret_cast<uint32_t*>(sQ + (mt * 16 + g + 8) * pitch + n0 + q4 * 2) = pack_bf16x2(o[2] * inv[mt][1], o[3] * inv[mt][1]);
      }
    }
    __syncwarp();
    {
      const int head = static_cast<int>(task % p.heads);
      const long long pix = task / p.heads;
      const long long row0 = (static_cast<long long>(pix / p.D) * p.F) * p.D + pix % p.D;
#pragma unroll
      for (int k = 0; k < kMaxCopies; ++k) {
        if (k < n_copies)
          *reinterpret_cast<uint4*>(p.o + (row0 + static_cast<long long>(CP_FRAME(k)) * p.D) * p.ldo + head * p.hd + CP_COL(k)) =
              *reinterpret_cast<const uint4*>(sQ + CP_SMEM(k));
      }
    }
    // rows >= F of the staging tile were written with garbage-free zeros only if F == R; restore the zero padding
    if (p.F < R) {
      for (int c = lane; c < (R - p.F) * (pitch / 8); c += 32)
        reinterpret_cast<uint4*>(sQ + p.F * pitch)[c] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
  }
#undef CP_FRAME
#undef CP_COL
#undef CP_SMEM
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
  if (F <= 32) {
    const int mt = F <= 16 ? 1 : 2;                     // 16-row tiles along the frame axis
    const int hd_pad = (head_dim + 15) / 16 * 16;
    const int pitch = hd_pad + 8;                       // +16 bytes per row: conflict-free fragment / ldmatrix reads
    // Warps per SM are what buys bandwidth here (each warp keeps one ~4 KB task in flight while it computes another;
    // measured: 16 -> 20 warps per SM = -7.5 % at head_dim 40, 8 -> 10 = -15 % at head_dim 80), so the launch takes as
    // many warps as ~216 KB of smem and 80 registers per thread allow, in one or two CTAs per SM.  A third stage at the
    // cost of warps (12 per SM) was slower (I360_TA_STAGES3 keeps it for experiments).
    const size_t per_stage = static_cast<size_t>(3) * 16 * mt * pitch * sizeof(bf16);
    const bool three = mt == 1 && getenv("I360_TA_STAGES3") != nullptr && per_stage * 3 * 6 <= 110 * 1024;
    const size_t pw = per_stage * (three ? 3 : 2);
    int total_w = static_cast<int>((216 * 1024) / pw);
    if (total_w > 24) total_w = 24;
    if (mt == 2 && total_w > 12) total_w = 12;          // the two-tile kernel is compiled for <= 384 threads (170 registers)
    if (total_w < 1) return I360_ERR_UNSUPPORTED;
    int ctas = total_w > 12 ? 2 : 1;
    int w2 = total_w / ctas;
    if (const char* e = getenv("I360_TA_WARPS")) { w2 = atoi(e); ctas = (pw * w2 * 2 <= 220 * 1024) ? 2 : 1; }
    if (three) { w2 = 6; ctas = 2; }
    const size_t sm2 = pw * w2;
    if (sm2 > 227 * 1024 || w2 < 1 || w2 > 12) return I360_ERR_UNSUPPORTED;       // compiled for <= 384 threads per CTA
    long long bl = (total + w2 - 1) / w2;
    const long long cap2 = static_cast<long long>(num_sms()) * ctas;
    if (bl > cap2) bl = cap2;
    static bool setm = false;
    if (!setm) {
      cudaFuncSetAttribute(temporal_attn_mma_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(temporal_attn_mma_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(temporal_attn_mma_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      setm = true;
    }
    if (mt == 2)    launch_k(temporal_attn_mma_kernel<2, 2>, dim3(static_cast<unsigned>(bl)), dim3(w2 * 32), sm2, st, p, hd_pad, pitch);
    else if (three) launch_k(temporal_attn_mma_kernel<3, 1>, dim3(static_cast<unsigned>(bl)), dim3(w2 * 32), sm2, st, p, hd_pad, pitch);
    else            launch_k(temporal_attn_mma_kernel<2, 1>, dim3(static_cast<unsigned>(bl)), dim3(w2 * 32), sm2, st, p, hd_pad, pitch);
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
