// Kernels only the conditioning encoders need: SAM ViT's decomposed relative position
// bias (segment_anything ImageEncoderViT, modeling/image_encoder.py `add_decomposed_rel_pos`, release 1.0):
//
//   bias[(item, head), q, k] = q_vec . Rh[qh - kh + S - 1] + q_vec . Rw[qw - kw + S - 1],   q = (qh, qw), k = (kh, kw)
//
// with q_vec the UNSCALED query of that head and Rh / Rw the block's [2S - 1, hd] tables.  One CTA handles QB queries
// of one (item, head): the 2 * S dot products per query go to shared memory in fp32, then the S^2 sums of a row are
// written as bf16 (one rounding), coalesced -- the pass is bound by that write (S^2 * 2 bytes per query and head).  The
// result feeds i360_attention_item_bias_bf16.
#include "common.cuh"
#include "tmap.h"

namespace i360 {

constexpr int kRelQB = 32;        // queries per CTA (the two tables, 32 KB for S = 64, are re-read from L2 by every CTA)
constexpr int kRelThreads = 256;

__global__ void __launch_bounds__(kRelThreads)
relpos_bias_kernel(const bf16* __restrict__ q, long long ldq, int col0, int heads, int hd, int S,
                   const bf16* __restrict__ rel_h, const bf16* __restrict__ rel_w, bf16* __restrict__ bias, int ldb) {
  griddep_wait();        // PDL: see common.cuh
  griddep_launch();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int N = S * S, L = 2 * S - 1, pitch = hd + 2;      // 33-word rows for hd 64: consecutive rows hit distinct banks
  bf16* sRel = reinterpret_cast<bf16*>(smem_raw);           // [2][L][pitch]
  float* sQ = reinterpret_cast<float*>(smem_raw + ((2 * L * pitch * 2 + 15) & ~15));      // [QB][hd], 16-byte aligned (float4 reads of sT)
  float* sT = sQ + kRelQB * hd;                             // [QB][2 S]
  const int q0 = blockIdx.x * kRelQB, head = blockIdx.y, item = blockIdx.z;
  const int nq = min(kRelQB, N - q0);
  {   // tables -> smem, 16-byte global loads, one division per thread (not per element)
    const int cpr = hd >> 3;                                 // 16-byte chunks per table row (hd % 8 == 0)
    const int r0 = threadIdx.x / cpr, c8 = threadIdx.x - r0 * cpr, rstep = kRelThreads / cpr;
    if (r0 < rstep)                                          // threads past the last whole row group idle (hd = 64: none)
      for (int r = r0; r < 2 * L; r += rstep) {
        const bf16* src = (r < L ? rel_h + r * hd : rel_w + (r - L) * hd) + c8 * 8;
        const uint4 v = *reinterpret_cast<const uint4*>(src);
        uint32_t* dst = reinterpret_cast<uint32_t*>(sRel + r * pitch + c8 * 8);      // pitch * 2 bytes is a multiple of 4
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
      }
  }
  for (int i = threadIdx.x; i < nq * hd; i += kRelThreads) {
    const int qi = i / hd, c = i % hd;
    sQ[i] = __bfloat162float(q[(static_cast<long long>(item) * N + q0 + qi) * ldq + col0 + head * hd + c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nq * 2 * S; i += kRelThreads) {
    const int qi = __float2int_rd((static_cast<float>(i) + 0.5f) * (0.5f / static_cast<float>(S))), j = i - qi * 2 * S;
    const int qq = q0 + qi, qh = __float2int_rd((static_cast<float>(qq) + 0.5f) / static_cast<float>(S)), qw = qq - qh * S;
    const int t = j >= S, kk = t ? j - S : j;
    const bf16* r = sRel + (t * L + ((t ? qw : qh) - kk + S - 1)) * pitch;
    const float* qv = sQ + qi * hd;
    float acc = 0.f;
    for (int c = 0; c < hd; c += 2) {
      const float2 rv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(r + c));
      acc = fmaf(qv[c], rv.x, acc);
      acc = fmaf(qv[c + 1], rv.y, acc);
    }
    sT[qi * 2 * S + j] = acc;
  }
  __syncthreads();
  // write-out: the pass is bound by this store (S^2 * 2 bytes per query and head).  No integer divisions in the loop (a
  // runtime k / S, k % S per element made the first version 8x slower than the store bandwidth): k / S comes from one float
  // multiply (exact for k < 2^22), and when S is a multiple of 8 a thread emits 8 keys of one key row with a 16-byte store.
  bf16* out = bias + (static_cast<long long>(item) * heads + head) * N * ldb + static_cast<long long>(q0) * ldb;
  const float invS = 1.0f / static_cast<float>(S);
  if ((S & 7) == 0) {
    const int vecs = ldb >> 3;                               // == N / 8 here (ldb == N)
    for (int qi = 0; qi < nq; ++qi)
    for (int vv = threadIdx.x; vv < vecs; vv += kRelThreads) {
      const int k = vv * 8;
      const float* T = sT + qi * 2 * S;
      const int kh = __float2int_rd((static_cast<float>(k) + 0.5f) * invS), kw = k - kh * S;
      const float th = T[kh];
      const float4 w0 = *reinterpret_cast<const float4*>(T + S + kw), w1 = *reinterpret_cast<const float4*>(T + S + kw + 4);
      uint4 o;
      o.x = pack_bf16x2(th + w0.x, th + w0.y); o.y = pack_bf16x2(th + w0.z, th + w0.w);
      o.z = pack_bf16x2(th + w1.x, th + w1.y); o.w = pack_bf16x2(th + w1.z, th + w1.w);
      *reinterpret_cast<uint4*>(out + static_cast<long long>(qi) * ldb + k) = o;
    }
  } else {
    const int pairs = ldb >> 1;                              // ldb is even (multiple of 8)
    for (int qi = 0; qi < nq; ++qi)
    for (int pp = threadIdx.x; pp < pairs; pp += kRelThreads) {
      const int k = pp * 2;
      const float* T = sT + qi * 2 * S;
      float a = 0.f, b = 0.f;
      if (k < N) { const int kh = __float2int_rd((static_cast<float>(k) + 0.5f) * invS); a = T[kh] + T[S + k - kh * S]; }
      if (k + 1 < N) { const int kh = __float2int_rd((static_cast<float>(k) + 1.5f) * invS); b = T[kh] + T[S + k + 1 - kh * S]; }
      *reinterpret_cast<__nv_bfloat162*>(out + static_cast<long long>(qi) * ldb + k) = __floats2bfloat162_rn(a, b);
    }
  }
}

}  // namespace i360

using namespace i360;

// q: [items * S*S, >= col0 + heads*hd] bf16 rows (row stride ldq elements), the head's query at columns
// col0 + head*hd; rel_h / rel_w: [2S-1, hd] bf16; bias: [items * heads * S*S, ldb] bf16, ldb % 8 == 0, ldb >= S*S
// (columns past S*S are written as 0).
extern "C" int i360_relpos_bias_bf16(const void* q, long long ldq, int col0, int items, int heads, int hd, int S,
                                     const void* rel_h, const void* rel_w, void* bias, int ldb, void* stream) {
  if (!q || !rel_h || !rel_w || !bias || items <= 0 || heads <= 0 || S <= 0) return I360_ERR_ARG;
  if (hd <= 0 || (hd % 8) || hd > 256 || (ldb % 8) || ldb < S * S) return I360_ERR_ARG;   // 16-byte table chunks / aligned sT rows
  const int L = 2 * S - 1, pitch = hd + 2;
  const size_t smem = ((static_cast<size_t>(2 * L * pitch) * 2 + 15) & ~static_cast<size_t>(15)) +
                      static_cast<size_t>(kRelQB) * (hd + 2 * S) * 4;
  if (smem > 200 * 1024) return I360_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) {
    static size_t attr = 0;
    if (smem > attr) {
      if (cudaFuncSetAttribute(relpos_bias_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
        return I360_ERR_CUDA;
      attr = smem;
    }
  }
  const dim3 grid((S * S + kRelQB - 1) / kRelQB, heads, items);
  if (grid.y > 65535u || grid.z > 65535u) return I360_ERR_ARG;
  launch_k(relpos_bias_kernel, dim3(grid), dim3(kRelThreads), smem, static_cast<cudaStream_t>(stream), static_cast<const bf16*>(q), ldq, col0, heads, hd, S, static_cast<const bf16*>(rel_h), static_cast<const bf16*>(rel_w),
      static_cast<bf16*>(bias), ldb);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}
