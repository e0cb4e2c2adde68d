// Fused softmax(Q K^T * scale + bias) V on tcgen05 tensor cores (sm_100a), head_dim 32 or 64.
//
// One CTA = one 128-row query tile of one (batch item, head).  Warp roles (192 threads):
//   warp0  TMA producer : Q once, then K_j / V_j (/ bias_j) tiles through 2-stage mbarrier rings
//   warp1  MMA issuer   : S = Q K_j^T  -> TMEM;  O_j = P_j V_j -> TMEM   (one elected lane)
//   warps 2..9 softmax  : TWO threads per query row (warp w and w+4 share a TMEM lane quarter): each owns 64 of
//                         the 128 key columns of S and half of O's columns; they exchange only the row maximum
//                         through smem.  TMEM S -> online softmax (exp2, fp32) -> P (bf16) into swizzled smem as
//                         the next MMA's A operand; O_j accumulated in registers with the running-max rescale;
//                         final O / l -> bf16 -> global.
// Two CTAs are co-resident per SM (<= 113 KB smem, 256 TMEM columns each) so one CTA's softmax
// overlaps the other's MMAs.
//
// Q, K, V and O are addressed as strided 4-D token views [channels, d1, d2, d3] so that all of the
// reference's attention variants run without any gather/transpose copy:
//   * spatial self-attention (diffusers/models/attention_processor.py:1210-1283): tokens of one frame,
//   * text / image-prompt cross-attention (animatediff/models/attention.py:65-156): K/V rows shared by
//     all frames of a clip (the reference repeats the context per frame, attention.py:257),
//   * WarpAttn perspective<->equirect cross-attention (src/modules/transformer.py:59-74): the "(m h w)"
//     token axis gathers m views that are F frames apart in memory; dense additive bias broadcast over
//     batch and heads (only mask[0] is used, :70).
#include "common.cuh"
#include "tmap.h"
#include <stdlib.h>

namespace i360 {

struct AttnOperand {
  // batch item bi -> fixed coordinates: c2 = (bi % A) / Bdiv, c3base = (bi / A) * mul
  int A, Bdiv, mul;
  int d1, ext3;      // tokens along dim1, dim3 entries per batch item
  int box1, box3;    // TMA box along dim1 / dim3 (box1 * box3 == 128)
  int n1;            // ceil(d1 / box1)
  int col0;          // channel offset of head 0
};

struct AttnParams {
  AttnOperand q, kv;
  int v_col0;                 // channel offset of head 0 inside the V view
  int q_tiles, kv_tiles;
  float scale_log2;           // softmax scale * log2(e)
  // output (same token mapping as q)
  bf16* o; long long os1, os2, os3; int o_col0;
  int accumulate;             // out = bf16(out + bf16(O))   (IP-adapter branch sum, attention.py:148)
  int has_bias; int bias_rows, bias_cols;
  int dbg;                    // I360_ATTN_DBG ablation bits (profiling only): 1 skip pass 1, 2 no MUFU, 4 no P store
};

constexpr int kAttnThreads = 320;   // warp0 TMA, warp1 MMA, warps 2..9 softmax (2 threads per query row)

template <int HD>
struct AttnCfg {
  static constexpr int kRowBytes = HD * 2;
  static constexpr int kQBytes = 128 * kRowBytes;
  static constexpr int kKVBytes = 128 * kRowBytes;
  static constexpr int kPBytes = 128 * 128 * 2;
  static constexpr int kBiasBytes = 128 * 128 * 2;
  static constexpr uint32_t kSwz = (HD == 64) ? SWZ_128B : SWZ_64B;
  static constexpr int kSBO = 8 * kRowBytes;
  static constexpr int kTmemCols = 256;   // S: 128, O: HD
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// One KV tile of the online softmax for one query row (= one thread): two passes over the 128 fp32 logits in
// TMEM.  Pass 1 reduces the row maximum; pass 2 evaluates p = exp2(s*scale - m) with one packed FFMA2 per two
// logits, accumulates the row sum with FADD2 and writes bf16 P into the 128B-swizzled smem tile that is the A
// operand of the P.V MMA.  MASKED (partial tiles) and BIAS (WarpAttn) are compile-time so the common case -- a
// full, unbiased tile of the spatial self-attention -- carries no selects and no per-element scale multiply.
template <bool BIAS, bool MASKED>
__device__ __forceinline__ void softmax_tile(uint32_t tS_row, const uint8_t* sB, uint8_t* sP, int row, int limit,
                                             float scale_log2, float& m_run, float& l_run, float& alpha, int cbeg,
                                             bf16* xmax_mine, const bf16* xmax_other, int dbg = 0) {
  const float LOG2E = 1.4426950408889634f;
  const uint32_t rsw = static_cast<uint32_t>(row & 7);
  float mx = -INFINITY;
  if (dbg & 1) mx = 8.0f;
#pragma unroll 1
  for (int c = cbeg; c < cbeg + ((dbg & 1) ? 0 : 64); c += 32) {
    uint32_t v[32];
    tmem_ld_x32(tS_row + c, v);
    tmem_ld_wait();
    if (!BIAS && !MASKED) {
#pragma unroll
      for (int e = 0; e < 32; e += 2) mx = fmax3(mx, __uint_as_float(v[e]), __uint_as_float(v[e + 1]));
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float bv[8];
        if (BIAS) {
          const int cc = c + g * 8;
          unpack8(*reinterpret_cast<const uint4*>(sB + (cc >> 6) * 16384 + row * 128 + ((((cc & 63) >> 3) ^ rsw) << 4)), bv);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float sv = __uint_as_float(v[g * 8 + e]) * scale_log2;
          if (BIAS) sv = fmaf(bv[e], LOG2E, sv);
          if (MASKED && (c + g * 8 + e >= limit)) sv = -INFINITY;
          mx = fmaxf(mx, sv);
        }
      }
    }
  }
  if (!BIAS && !MASKED) mx *= scale_log2;          // scale > 0: max commutes with the scaling
  // the partner thread (other half of the key columns of this row) contributes its maximum through smem.  Both
  // sides use the bf16-ROUNDED maxima (512 B exchange buffer keeps two CTAs per SM): any common shift is a valid
  // softmax stabiliser, it only has to be identical in both halves and within ~2^0.2 of the true maximum.
  const bf16 mxb = __float2bfloat16(mx);
  *xmax_mine = mxb;
  named_bar_sync(1, 256);
  mx = fmaxf(__bfloat162float(mxb), __bfloat162float(*xmax_other));
  const float m_new = fmaxf(m_run, mx);
  alpha = fast_exp2(m_run - m_new);               // first tile: exp2(-inf) = 0
  const float2 sc2 = make_float2(scale_log2, scale_log2);
  const float2 nm2 = make_float2(-m_new, -m_new);
  const float2 l2e2 = make_float2(LOG2E, LOG2E);
  float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int c = cbeg; c < cbeg + 64; c += 32) {
    uint32_t v[32];
    tmem_ld_x32(tS_row + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int cc = c + g * 8;
      float bv[8];
      if (BIAS) unpack8(*reinterpret_cast<const uint4*>(sB + (cc >> 6) * 16384 + row * 128 + ((((cc & 63) >> 3) ^ rsw) << 4)), bv);
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        float2 off = nm2;
        if (BIAS) off = ffma2(make_float2(bv[e], bv[e + 1]), l2e2, nm2);
        const float2 t = ffma2(make_float2(__uint_as_float(v[g * 8 + e]), __uint_as_float(v[g * 8 + e + 1])), sc2, off);
        float2 pe = (dbg & 2) ? make_float2(t.x * 1e-3f, t.y * 1e-3f) : make_float2(fast_exp2(t.x), fast_exp2(t.y));
        if (MASKED) {
          if (cc + e >= limit) pe.x = 0.f;
          if (cc + e + 1 >= limit) pe.y = 0.f;
        }
        sum2 = fadd2(sum2, pe);
        pk[e >> 1] = pack_bf16x2(pe.x, pe.y);
      }
      if (!(dbg & 4))
        *reinterpret_cast<uint4*>(sP + (cc >> 6) * 16384 + row * 128 + ((((cc & 63) >> 3) ^ rsw) << 4)) =
            make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
  l_run = l_run * alpha + (sum2.x + sum2.y);
  m_run = m_new;
}

__device__ __forceinline__ void tile_coords(const AttnOperand& op, int bi, int tile, int& c1, int& c2, int& c3) {
  c1 = (tile % op.n1) * op.box1;
  c2 = (bi % op.A) / op.Bdiv;
  c3 = (bi / op.A) * op.mul + (tile / op.n1) * op.box3;
}

template <int HD, bool BIAS>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmB,
                 const AttnParams p) {
  using C = AttnCfg<HD>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + C::kQBytes;             // 2 stages
  uint8_t* sV = sK + 2 * C::kKVBytes;        // 2 stages
  uint8_t* sP = sV + 2 * C::kKVBytes;
  uint8_t* sB = sP + C::kPBytes;             // bias tile (BIAS only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (BIAS ? C::kBiasBytes : 0));
  uint64_t* q_full = bars;         // 1
  uint64_t* k_full = bars + 1;     // 2
  uint64_t* k_empty = bars + 3;    // 2
  uint64_t* v_full = bars + 5;     // 2
  uint64_t* v_empty = bars + 7;    // 2
  uint64_t* s_full = bars + 9;     // MMA -> softmax: S_j in TMEM
  uint64_t* p_full = bars + 10;    // softmax -> MMA: P_j in smem, S_j consumed, O_{j-1} consumed
  uint64_t* o_full = bars + 11;    // MMA -> softmax: O_j in TMEM
  uint64_t* b_full = bars + 12;    // TMA -> softmax: bias_j in smem
  uint64_t* b_empty = bars + 13;   // softmax -> TMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  bf16* xch = reinterpret_cast<bf16*>(bars + 16);            // [2][128] row-max exchange between the two column halves

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, bi = blockIdx.z;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    if (BIAS) tma_prefetch_desc(&tmB);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 256); mbar_init(o_full, 1);
    mbar_init(b_full, 1); mbar_init(b_empty, 256);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, C::kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  int qc1, qc2, qc3;
  tile_coords(p.q, bi, qt, qc1, qc2, qc3);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, C::kQBytes);
      tma_load_4d(sQ, &tmQ, q_full, p.q.col0 + head * HD, qc1, qc2, qc3);
      const int q_base = qt * 128;   // the bias arrives tile-padded: [q_tiles*128, kv_tiles*128]
      for (int j = 0; j < p.kv_tiles; ++j) {
        const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
        int c1, c2, c3;
        tile_coords(p.kv, bi, j, c1, c2, c3);
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], C::kKVBytes);
        tma_load_4d(sK + st * C::kKVBytes, &tmK, &k_full[st], p.kv.col0 + head * HD, c1, c2, c3);
        if (BIAS) {
          const int kv_base = j * 128;
          mbar_wait(b_empty, (j & 1) ^ 1);
          mbar_expect_tx(b_full, C::kBiasBytes);
          tma_load_2d(sB, &tmB, b_full, kv_base, q_base);
          tma_load_2d(sB + C::kBiasBytes / 2, &tmB, b_full, kv_base + 64, q_base);
        }
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], C::kKVBytes);
        tma_load_4d(sV + st * C::kKVBytes, &tmV, &v_full[st], p.v_col0 + head * HD, c1, c2, c3);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);   // B (=V) is MN-major
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      mbar_wait(q_full, 0);
      for (int j = 0; j < p.kv_tiles; ++j) {
        const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
        // ---- S_j = Q K_j^T ----
        mbar_wait(&k_full[st], ph);
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + st * C::kKVBytes);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_bf16_ss(tS, make_smem_desc(aQ + ks * 32, C::kSBO, 16, C::kSwz),
                       make_smem_desc(aK + ks * 32, C::kSBO, 16, C::kSwz), idesc_s, ks != 0);
        umma_commit(&k_empty[st]);
        umma_commit(s_full);
        // ---- O_j = P_j V_j ----
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[st], ph);
        tc_fence_after();
        const uint32_t aV = smem_u32(sV + st * C::kKVBytes);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ss(tO, make_smem_desc(aP + (kk >> 2) * 16384 + (kk & 3) * 32, 1024, 16, SWZ_128B),
                       make_smem_desc(aV + kk * 16 * C::kRowBytes, C::kSBO, 16, C::kSwz), idesc_o, kk != 0);
        umma_commit(&v_empty[st]);
        umma_commit(o_full);
      }
    }
  } else {
    // ================================ softmax / epilogue ================================
    const int ew = warp & 3;
    const int half = (warp - 2) >> 2;            // 0: key columns 0..63 / O columns [0, HD/2); 1: the other halves
    const int row = ew * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
    constexpr int HH = HD / 2;
    bf16* xmine = xch + half * 128 + row;
    const bf16* xother = xch + (half ^ 1) * 128 + row;
    // my query token
    const int q_tok = (qt % p.q.n1) * p.q.box1 + row % p.q.box1;
    const int q_view = (qt / p.q.n1) * p.q.box3 + row / p.q.box1;
    const bool q_valid = (q_tok < p.q.d1) && (q_view < p.q.ext3);
    float m_run = -INFINITY, l_run = 0.f;
    float acc[HH];
#pragma unroll
    for (int i = 0; i < HH; ++i) acc[i] = 0.f;

    for (int j = 0; j < p.kv_tiles; ++j) {
      const int kv_i1 = (j % p.kv.n1) * p.kv.box1, kv_i3 = (j / p.kv.n1) * p.kv.box3;
      // columns [0, limit) of this tile hold real keys (tile = 128 tokens of one view, or box3 whole views)
      int limit;
      if (p.kv.box3 == 1) limit = (kv_i3 < p.kv.ext3) ? min(128, p.kv.d1 - kv_i1) : 0;
      else limit = min(128, (p.kv.ext3 - kv_i3) * p.kv.box1);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (BIAS) mbar_wait(b_full, j & 1);
      float alpha;
      if (limit >= 128) softmax_tile<BIAS, false>(tS + lane_sel, sB, sP, row, limit, p.scale_log2, m_run, l_run, alpha, half * 64, xmine, xother, p.dbg);
      else              softmax_tile<BIAS, true>(tS + lane_sel, sB, sP, row, limit, p.scale_log2, m_run, l_run, alpha, half * 64, xmine, xother);
      fence_proxy_async_smem();       // P visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_full);
      if (BIAS) mbar_arrive(b_empty);
      // O_j: my half of the head-dim columns
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      {
        uint32_t v[HH];
        if (HH == 32) tmem_ld_x32(tO + lane_sel + half * HH, v);
        else          tmem_ld_x16(tO + lane_sel + half * HH, v);
        tmem_ld_wait();
        const float2 al2 = make_float2(alpha, alpha);
#pragma unroll
        for (int e = 0; e < HH; e += 2) {
          const float2 r = ffma2(make_float2(acc[e], acc[e + 1]), al2, make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])));
          acc[e] = r.x; acc[e + 1] = r.y;
        }
      }
      tc_fence_before();
    }
    // ---- epilogue: combine the two halves' row sums, write my half of O ----
    float* lsum = reinterpret_cast<float*>(sP);        // the P tile is dead after the last P.V MMA
    lsum[half * 128 + row] = l_run;
    named_bar_sync(1, 256);
    const float l_tot = l_run + lsum[(half ^ 1) * 128 + row];
    if (q_valid) {
      const float inv = 1.0f / l_tot;
      bf16* dst = p.o + p.o_col0 + head * HD + half * HH + static_cast<long long>(q_tok) * p.os1 +
                  static_cast<long long>(qc2) * p.os2 +
                  static_cast<long long>((bi / p.q.A) * p.q.mul + q_view) * p.os3;
#pragma unroll
      for (int c = 0; c < HH; c += 8) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = acc[c + e] * inv;
        if (p.accumulate) {
          const uint4 old = *reinterpret_cast<const uint4*>(dst + c);
          float prev[8];
          unpack8(old, prev);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = prev[e] + __bfloat162float(__float2bfloat16(o[e]));
        }
        *reinterpret_cast<uint4*>(dst + c) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                        pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, C::kTmemCols); }
}

template <int HD, bool BIAS>
static int launch_attn(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const CUtensorMap& b,
                       const AttnParams& p, int heads, int batch, cudaStream_t st) {
  using C = AttnCfg<HD>;
  const int smem = C::kQBytes + 4 * C::kKVBytes + C::kPBytes + (BIAS ? C::kBiasBytes : 0) + 128 + 512;   // barriers + row-max exchange
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(attention_kernel<HD, BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return I360_ERR_CUDA;
    attr_set = true;
  }
  attention_kernel<HD, BIAS><<<dim3(p.q_tiles, heads, batch), kAttnThreads, smem, st>>>(q, k, v, b, p);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}

}  // namespace i360

using namespace i360;

// A strided 4-D token view handed over the C ABI (all strides in ELEMENTS, channel stride == 1).
struct I360TokenView {
  const void* ptr;
  int channels;          // extent of the contiguous channel dim of the underlying rows
  int col0;              // channel offset of head 0
  int d1, d2, d3;        // extents of token dims 1..3
  long long s1, s2, s3;  // strides of token dims 1..3
  int A, Bdiv, mul;      // batch item -> (c2, c3base):  c2 = (bi % A) / Bdiv, c3base = (bi / A) * mul
  int ext3;              // dim3 entries that belong to one batch item (views); 1 for plain sequences
};

static int make_view_map(CUtensorMap* m, const I360TokenView& v, int hd, int box1, int box3) {
  uint64_t d[4] = {(uint64_t)v.channels, (uint64_t)v.d1, (uint64_t)v.d2, (uint64_t)v.d3};
  uint64_t s[3] = {(uint64_t)v.s1 * 2, (uint64_t)v.s2 * 2, (uint64_t)v.s3 * 2};
  uint32_t b[4] = {(uint32_t)hd, (uint32_t)box1, 1u, (uint32_t)box3};
  return get_tmap_bf16(m, v.ptr, 4, d, s, b, hd == 64 ? 3 : 2);
}

static void fill_operand(AttnOperand* o, const I360TokenView& v) {
  o->A = v.A > 0 ? v.A : 0x7fffffff; o->Bdiv = v.Bdiv > 0 ? v.Bdiv : 1; o->mul = v.mul;
  o->d1 = v.d1; o->ext3 = v.ext3 > 0 ? v.ext3 : 1;
  // A 128-row tile is either 128 consecutive tokens of one view (rows past d1 are zero-filled by TMA and
  // masked), or -- for short views that divide 128 -- 128/d1 whole views packed together.
  if (o->ext3 > 1 && v.d1 < 128 && (128 % v.d1) == 0) { o->box1 = v.d1; o->box3 = 128 / v.d1; }
  else { o->box1 = 128; o->box3 = 1; }
  o->n1 = (v.d1 + o->box1 - 1) / o->box1;
  o->col0 = v.col0;
}

// q/k/v/o: token views; heads x head_dim channels starting at col0 of each view; bias: optional bf16
// additive logits bias shared by all batch items and heads, in TILE layout: row qt*128 + r is the query held by
// row r of query tile qt, column j*128 + c the key held by row c of key tile j (see fill_operand: a tile is 128
// consecutive tokens of one view, or 128/d1 whole views when d1 divides 128).  For sequences whose views are
// multiples of 128 tokens (every level of the 512x1024 / 256x512 configurations) this IS the dense [Nq, Nk] matrix.
extern "C" int i360_attention_bf16(const I360TokenView* q, const I360TokenView* k, const I360TokenView* v,
                                   const I360TokenView* o, int heads, int head_dim, int batch, float scale,
                                   const void* bias, int bias_rows, int bias_cols, int accumulate, void* stream) {
  if (!q || !k || !v || !o || !q->ptr || !k->ptr || !v->ptr || !o->ptr) return I360_ERR_ARG;
  if (head_dim != 32 && head_dim != 64) return I360_ERR_UNSUPPORTED;
  if (heads <= 0 || batch <= 0) return I360_ERR_ARG;
  if ((q->s1 % 8) || (k->s1 % 8) || (v->s1 % 8) || (o->s1 % 8) || (q->col0 % 8) || (k->col0 % 8) || (v->col0 % 8) || (o->col0 % 8))
    return I360_ERR_ARG;
  if (k->d1 != v->d1 || k->ext3 != v->ext3 || k->s1 != v->s1 || k->s2 != v->s2 || k->s3 != v->s3) return I360_ERR_ARG;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  fill_operand(&p.q, *q);
  fill_operand(&p.kv, *k);
  p.v_col0 = v->col0;
  p.q_tiles = p.q.n1 * ((p.q.ext3 + p.q.box3 - 1) / p.q.box3);
  p.kv_tiles = p.kv.n1 * ((p.kv.ext3 + p.kv.box3 - 1) / p.kv.box3);
  p.scale_log2 = scale * 1.4426950408889634f;
  p.o = static_cast<bf16*>(const_cast<void*>(o->ptr)); p.os1 = o->s1; p.os2 = o->s2; p.os3 = o->s3; p.o_col0 = o->col0;
  p.accumulate = accumulate;
  p.has_bias = bias != nullptr; p.bias_rows = bias_rows; p.bias_cols = bias_cols;
  { const char* e = getenv("I360_ATTN_DBG"); p.dbg = e ? atoi(e) : 0; }
  CUtensorMap tq, tk, tv, tb;
  int r = make_view_map(&tq, *q, head_dim, p.q.box1, p.q.box3); if (r) return r;
  r = make_view_map(&tk, *k, head_dim, p.kv.box1, p.kv.box3); if (r) return r;
  r = make_view_map(&tv, *v, head_dim, p.kv.box1, p.kv.box3); if (r) return r;
  tb = tq;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bias) {
    if (bias_cols % 8) return I360_ERR_ARG;
    uint64_t d[2] = {(uint64_t)bias_cols, (uint64_t)bias_rows}; uint64_t s[1] = {(uint64_t)bias_cols * 2};
    uint32_t b[2] = {64, 128};
    r = get_tmap_bf16(&tb, bias, 2, d, s, b, 3); if (r) return r;
    if (head_dim == 64) return launch_attn<64, true>(tq, tk, tv, tb, p, heads, batch, st);
    return launch_attn<32, true>(tq, tk, tv, tb, p, heads, batch, st);
  }
  if (head_dim == 64) return launch_attn<64, false>(tq, tk, tv, tb, p, heads, batch, st);
  return launch_attn<32, false>(tq, tk, tv, tb, p, heads, batch, st);
}
