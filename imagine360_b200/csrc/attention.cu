// Fused softmax(Q K^T * scale + bias) V on tcgen05 tensor cores (sm_100a), head_dim 32 or 64.
//
// One CTA = one 128-row query tile of one (batch item, head).  Warp roles (192 threads):
//   warp0  TMA producer : Q once, then K_j / V_j (/ bias_j) tiles through 2-stage mbarrier rings
//   warp1  MMA issuer   : S_j = Q K_j^T -> TMEM;  O_h += P_j[:, h] V_j[h, :] -> TMEM, h = 0, 1   (one elected lane)
//   warps 2..9 softmax  : TWO threads per query row (warp w and w+4 share a TMEM lane quarter): each owns 64 of
//                         the 128 key columns of every tile as an INDEPENDENT online softmax (own max, own sum, own
//                         TMEM accumulator), merged once at the end.  Per tile: S -> registers (one TMEM read, S
//                         released at once so S_{j+1} overlaps), exp2 in fp32, P (bf16) into swizzled smem as the
//                         next MMA's A operand.  The accumulators stay in TMEM (use_acc) and are rescaled lazily.
// Two CTAs are co-resident per SM (<= 113 KB smem, 256 TMEM columns each).
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
  int bias_item_rows;         // 0: one bias for every (batch item, head); else rows of the bias per (batch item, head) block
};

#ifndef I360_POLY_MASK
#define I360_POLY_MASK 0x22   // which of every 8 logit pairs evaluate 2^t on the FMA pipe instead of the MUFU (bit i = pair i)
#endif
#ifndef I360_POLY_MASK2
#define I360_POLY_MASK2 0x22  // same, for attention2_kernel
#endif
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
  static constexpr int kTmemCols = 256;   // S: 128, O_half 0 / 1: HD each
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// One KV tile of the online softmax for one (query row, column half) = one thread.  The thread's 64 fp32 logits are
// read from TMEM ONCE into registers and S is released to the MMA warp immediately (S_{j+1} = Q K_{j+1}^T then runs
// under this tile's exponentials).  The two column halves of a row are independent online softmaxes, each with its
// own running maximum, row sum and TMEM accumulator O_half (+= P_half V_half, accumulated by the tensor core), so
// no per-tile exchange or CTA barrier exists; they are merged once after the last tile.  The accumulator is only
// rescaled when the maximum grew by more than 2^8 ("lazy rescale": any common shift is a valid stabiliser, P <= 256
// is exact enough in bf16 and fp32); that path is rare after the first tiles and warp-uniform (tcgen05.ld/st).
// MASKED (partial tiles) and BIAS (WarpAttn) are compile-time so the common case -- a full, unbiased tile of the
// spatial self-attention -- carries no selects and no per-element scale multiply.
template <int HD, bool BIAS, bool MASKED>
__device__ __forceinline__ void softmax_tile(uint32_t tS_mine, uint32_t tO_mine, const uint8_t* sB, uint8_t* sP,
                                             int row, int limit, float scale_log2, float& m_used, float& l_run,
                                             int cbeg, bool have_acc, bool wait_prev, uint64_t* s_empty,
                                             uint64_t* pv_done, uint32_t pv_parity) {
  const float LOG2E = 1.4426950408889634f;
  const uint32_t rsw = static_cast<uint32_t>(row & 7);
  uint32_t v[64];
  tmem_ld_x32(tS_mine, v);
  tmem_ld_x32(tS_mine + 32, v + 32);
  tmem_ld_wait();
  tc_fence_before();
  mbar_arrive(s_empty);                            // S_j is in registers: the MMA warp may overwrite it
  float mx = -INFINITY;
  if (!BIAS && !MASKED) {
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};      // four independent chains (FMNMX3 latency)
#pragma unroll
    for (int e = 0; e < 64; e += 8) {
#pragma unroll
      for (int c = 0; c < 4; ++c) mx4[c] = fmax3(mx4[c], __uint_as_float(v[e + 2 * c]), __uint_as_float(v[e + 2 * c + 1]));
    }
    mx = fmax3(mx4[0], mx4[1], fmaxf(mx4[2], mx4[3])) * scale_log2;   // scale > 0: max commutes with the scaling
  } else {
    // packed arithmetic (FMUL2 / FFMA2 on logit pairs, FMNMX3 on the pair): WarpAttn's kernel is bound by the issue
    // slots of this loop -- head_dim 32 halves the tensor work per logit -- so every scalar op here costs step time
    const float2 sc2 = make_float2(scale_log2, scale_log2), l2e = make_float2(LOG2E, LOG2E);
    float mx2[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int cc = cbeg + g * 8;
      uint4 bb = make_uint4(0u, 0u, 0u, 0u);
      if (BIAS) bb = *reinterpret_cast<const uint4*>(sB + (cc >> 6) * 16384 + row * 128 + ((((cc & 63) >> 3) ^ rsw) << 4));
      const uint32_t bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 sv = fmul2(make_float2(__uint_as_float(v[g * 8 + 2 * e]), __uint_as_float(v[g * 8 + 2 * e + 1])), sc2);
        if (BIAS) sv = ffma2(unpack_bf16x2(bw[e]), l2e, sv);
        if (MASKED) {
          if (cc + 2 * e >= limit) sv.x = -INFINITY;
          if (cc + 2 * e + 1 >= limit) sv.y = -INFINITY;
        }
        v[g * 8 + 2 * e] = __float_as_uint(sv.x);       // keep the finished logits (log2 units)
        v[g * 8 + 2 * e + 1] = __float_as_uint(sv.y);
        mx2[e & 1] = fmax3(mx2[e & 1], sv.x, sv.y);
      }
    }
    mx = fmaxf(mx2[0], mx2[1]);
  }
  const float m_new = fmaxf(m_used, mx);
  const bool need = m_new > m_used + 8.0f;         // also true for the first finite maximum (m_used = -inf)
  if (__any_sync(0xffffffffu, need)) {
    const float alpha = (m_new == m_used) ? 1.0f : fast_exp2(m_used - m_new);
    if (have_acc) {
      mbar_wait(pv_done, pv_parity);               // every earlier P V (this head's included) has landed in the accumulator
      tc_fence_after();
      const float2 al2 = make_float2(alpha, alpha);
#pragma unroll 1
      for (int c = 0; c < HD; c += 16) {
        uint32_t o[16];
        tmem_ld_x16(tO_mine + c, o);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const float2 r = fmul2(make_float2(__uint_as_float(o[e]), __uint_as_float(o[e + 1])), al2);
          o[e] = __float_as_uint(r.x); o[e + 1] = __float_as_uint(r.y);
        }
        tmem_st_x16(tO_mine + c, o);
      }
      tmem_st_wait();
    }
    l_run *= alpha;
    m_used = m_new;
  }
  // a half row that has not seen a finite logit yet (masked / -inf bias) must not evaluate exp2(-inf + inf)
  const float msub = (BIAS || MASKED) ? ((m_used == -INFINITY) ? 0.f : m_used) : m_used;
  const float2 sc2 = make_float2(scale_log2, scale_log2);
  const float2 nm2 = make_float2(-msub, -msub);
  float2 sum2 = make_float2(0.f, 0.f);
  uint32_t pk[32];
#pragma unroll
  for (int e = 0; e < 64; e += 2) {
    const float2 sv = make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1]));
    const float2 t = (!BIAS && !MASKED) ? ffma2(sv, sc2, nm2) : fadd2(sv, nm2);
    // 2 of every 8 pairs of the plain path evaluate 2^t on the FMA / integer pipes instead of the MUFU.  ncu: the
    // kernel is bound by issue slots (65 % busy, ~511 instructions per thread and tile) and the MUFU pipe (16 ex2 per
    // clock per SM) together; a polynomial pair costs ~10 issue slots against 2 MUFU instructions, and 2/8 balances the
    // two (pano level 0: 3/8 3.61 ms, 2/8 3.51 ms, 1/8 3.57 ms, 0/8 3.73 ms)
    const bool poly = !BIAS && !MASKED && ((I360_POLY_MASK >> ((e >> 1) & 7)) & 1);
    const float2 pe = poly ? exp2_poly2(t) : make_float2(fast_exp2(t.x), fast_exp2(t.y));   // masked: exp2(-inf) = 0
    sum2 = fadd2(sum2, pe);
    pk[e >> 1] = pack_bf16x2(pe.x, pe.y);
  }
  l_run += sum2.x + sum2.y;
  if (wait_prev) mbar_wait(pv_done, pv_parity);    // the tensor core has finished reading the previous P from smem
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const int cc = cbeg + g * 8;
    *reinterpret_cast<uint4*>(sP + (cc >> 6) * 16384 + row * 128 + ((((cc & 63) >> 3) ^ rsw) << 4)) =
        make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
  }
}

__device__ __forceinline__ void tile_coords(const AttnOperand& op, int bi, int tile, int& c1, int& c2, int& c3) {
  c1 = (tile % op.n1) * op.box1;
  c2 = (bi % op.A) / op.Bdiv;
  c3 = (bi / op.A) * op.mul + (tile / op.n1) * op.box3;
}

// G = heads handled by one CTA.  WarpAttn's dense bias tile (32 KB per 128 x 128 logits, the same for every head and batch
// item) is what its kernel waits for: with one head per CTA every (head, batch item) re-streams the whole bias from L2
// (13.4 GB per call at the first level against 0.4 GB of K + V).  With G = 2 the CTA keeps the bias tile in smem for two
// heads -- Q_g, K_g, V_g flow through the same rings as "virtual tiles" i = j * G + g -- which halves that traffic; the
// two extra accumulators fit because head_dim 32 leaves half of the 256 TMEM columns unused (S 128 + 4 x 32).
template <int HD, bool BIAS, int G>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmB,
                 const AttnParams p) {
  using C = AttnCfg<HD>;
  static_assert(G == 1 || G == 2, "heads per CTA");
  static_assert(128 + G * 2 * HD <= C::kTmemCols, "accumulators of all heads must fit next to S");
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) i360_device_fail("dynamic shared memory is not 1024-byte aligned (128B-swizzled TMA / UMMA tiles)");
  uint8_t* sQ = smem;                        // G query tiles (one per head)
  uint8_t* sK = sQ + G * C::kQBytes;         // 2 stages
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
  uint64_t* s_empty = bars + 10;   // softmax -> MMA: S_j copied to registers (256 arrivals)
  uint64_t* p_full = bars + 11;    // 2: softmax half h -> MMA: P_j[:, h*64..] in smem, O_h rescaled if needed (128 arrivals)
  uint64_t* pv_done = bars + 13;   // 2: MMA -> softmax half h: O_h += P_j V_j retired (P smem free, accumulator readable)
  uint64_t* b_full = bars + 15;    // TMA -> softmax: bias_j in smem
  uint64_t* b_empty = bars + 16;   // softmax -> TMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head0 = blockIdx.y * G, bi = blockIdx.z;
  const int n_virtual = p.kv_tiles * G;         // (kv tile, head) pairs, head fastest

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    if (BIAS) tma_prefetch_desc(&tmB);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
      mbar_init(&p_full[s], 128); mbar_init(&pv_done[s], 1);
    }
    mbar_init(s_full, 1); mbar_init(s_empty, 256);
    mbar_init(b_full, 1); mbar_init(b_empty, 256);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, C::kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();        // set-up done under the previous kernel's tail; from here on global memory is touched
  griddep_launch();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;      // O of (head g, column half h) at tO + (g * 2 + h) * HD

  int qc1, qc2, qc3;
  tile_coords(p.q, bi, qt, qc1, qc2, qc3);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, G * C::kQBytes);
#pragma unroll
      for (int g = 0; g < G; ++g)
        tma_load_4d(sQ + g * C::kQBytes, &tmQ, q_full, p.q.col0 + (head0 + g) * HD, qc1, qc2, qc3);
      // the bias arrives tile-padded: [q_tiles*128, kv_tiles*128]; per-item biases (SAM's decomposed relative position
      // term, one block of bias_item_rows rows per (batch item, head)) are stacked along the row axis
      const int q_base = qt * 128 + (bi * static_cast<int>(gridDim.y) * G + head0) * p.bias_item_rows;
      for (int j = 0; j < p.kv_tiles; ++j) {
        int c1, c2, c3;
        tile_coords(p.kv, bi, j, c1, c2, c3);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int i = j * G + g;
          const int st = i & 1; const uint32_t ph = (i >> 1) & 1;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_expect_tx(&k_full[st], C::kKVBytes);
          tma_load_4d(sK + st * C::kKVBytes, &tmK, &k_full[st], p.kv.col0 + (head0 + g) * HD, c1, c2, c3);
          if (BIAS && g == 0) {          // one bias tile serves the G heads of this kv tile
            const int kv_base = j * 128;
            mbar_wait(b_empty, (j & 1) ^ 1);
            mbar_expect_tx(b_full, C::kBiasBytes);
            tma_load_2d(sB, &tmB, b_full, kv_base, q_base);
            tma_load_2d(sB + C::kBiasBytes / 2, &tmB, b_full, kv_base + 64, q_base);
          }
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_expect_tx(&v_full[st], C::kKVBytes);
          tma_load_4d(sV + st * C::kKVBytes, &tmV, &v_full[st], p.v_col0 + (head0 + g) * HD, c1, c2, c3);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);   // B (=V) is MN-major
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      auto issue_s = [&](int i) {                   // S_i = Q_g K_{j,g}^T for the virtual tile i = j * G + g
        const int st = i & 1; const uint32_t ph = (i >> 1) & 1;
        mbar_wait(&k_full[st], ph);
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + st * C::kKVBytes);
        const uint32_t aQg = aQ + (i & (G - 1)) * C::kQBytes;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_bf16_ss(tS, make_smem_desc(aQg + ks * 32, C::kSBO, 16, C::kSwz),
                       make_smem_desc(aK + ks * 32, C::kSBO, 16, C::kSwz), idesc_s, ks != 0);
        umma_commit(&k_empty[st]);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int i = 0; i < n_virtual; ++i) {
        const int st = i & 1; const uint32_t ph = (i >> 1) & 1;
        const int g = i & (G - 1);
        const bool have_acc = i >= G;               // kv tile j > 0: the head's accumulator holds earlier tiles
        if (i + 1 < n_virtual) {                    // runs on the tensor core under the softmax of tile i
          mbar_wait(s_empty, i & 1);
          issue_s(i + 1);
        }
        // ---- O_{g,h} += P_i[:, h*64 .. h*64+64) V_i[h*64 .. h*64+64, :] ----
        mbar_wait(&v_full[st], ph);
        const uint32_t aV = smem_u32(sV + st * C::kKVBytes);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&p_full[h], i & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = h * 4; kk < h * 4 + 4; ++kk)
            umma_bf16_ss(tO + (g * 2 + h) * HD, make_smem_desc(aP + (kk >> 2) * 16384 + (kk & 3) * 32, 1024, 16, SWZ_128B),
                         make_smem_desc(aV + kk * 16 * C::kRowBytes, C::kSBO, 16, C::kSwz), idesc_o,
                         have_acc || (kk != h * 4));
          umma_commit(&pv_done[h]);
        }
        umma_commit(&v_empty[st]);
      }
    }
  } else {
    // ================================ softmax / epilogue ================================
    const int ew = warp & 3;
    const int half = (warp - 2) >> 2;            // 0: key columns 0..63 of every tile; 1: columns 64..127
    const int row = ew * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
    constexpr int HH = HD / 2;
    // my query token
    const int q_tok = (qt % p.q.n1) * p.q.box1 + row % p.q.box1;
    const int q_view = (qt / p.q.n1) * p.q.box3 + row / p.q.box1;
    const bool q_valid = (q_tok < p.q.d1) && (q_view < p.q.ext3);
    float m_used[G], l_run[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { m_used[g] = -INFINITY; l_run[g] = 0.f; }
    const uint32_t tS_mine = tS + lane_sel + half * 64;

    int jn = 0, jq = 0;                          // j % n1, j / n1 kept incrementally (a divide per tile costs ~45 issue slots)
    for (int j = 0; j < p.kv_tiles; ++j) {
      const int kv_i1 = jn * p.kv.box1, kv_i3 = jq * p.kv.box3;
      if (++jn == p.kv.n1) { jn = 0; ++jq; }
      // columns [0, limit) of this tile hold real keys (tile = 128 tokens of one view, or box3 whole views)
      int limit;
      if (p.kv.box3 == 1) limit = (kv_i3 < p.kv.ext3) ? min(128, p.kv.d1 - kv_i1) : 0;
      else limit = min(128, (p.kv.ext3 - kv_i3) * p.kv.box1);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int i = j * G + g;
        const uint32_t tO_mine = tO + lane_sel + (g * 2 + half) * HD;
        mbar_wait(s_full, i & 1);
        tc_fence_after();
        if (BIAS && g == 0) mbar_wait(b_full, j & 1);
        const uint32_t pvp = (i - 1) & 1;
        if (limit >= 128) softmax_tile<HD, BIAS, false>(tS_mine, tO_mine, sB, sP, row, limit, p.scale_log2, m_used[g], l_run[g], half * 64, j > 0, i > 0, s_empty, &pv_done[half], pvp);
        else              softmax_tile<HD, BIAS, true>(tS_mine, tO_mine, sB, sP, row, limit, p.scale_log2, m_used[g], l_run[g], half * 64, j > 0, i > 0, s_empty, &pv_done[half], pvp);
        fence_proxy_async_smem();       // P visible to the tensor core (async proxy)
        tc_fence_before();              // ... and the rescaled accumulator (tcgen05.st) ordered before the arrive
        mbar_arrive(&p_full[half]);
        if (BIAS && g == G - 1) mbar_arrive(b_empty);
      }
    }
    // ---- epilogue: merge the two column halves of my row, write my half of the head-dim columns ----
    const uint32_t lastp = (n_virtual - 1) & 1;
    mbar_wait(&pv_done[0], lastp);
    mbar_wait(&pv_done[1], lastp);
    tc_fence_after();
    float2* xml = reinterpret_cast<float2*>(sP);       // the P tile is dead after the last P.V MMA
#pragma unroll
    for (int g = 0; g < G; ++g) xml[(g * 2 + half) * 128 + row] = make_float2(m_used[g], l_run[g]);
    named_bar_sync(1, 256);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float2 oth = xml[(g * 2 + (half ^ 1)) * 128 + row];
      const float m_all = fmaxf(m_used[g], oth.x);
      const float w_me = (m_used[g] == -INFINITY) ? 0.f : fast_exp2(m_used[g] - m_all);
      const float w_ot = (oth.x == -INFINITY) ? 0.f : fast_exp2(oth.x - m_all);
      const float inv = 1.0f / (l_run[g] * w_me + oth.y * w_ot);
      const float w_lo = (half == 0 ? w_me : w_ot) * inv, w_hi = (half == 0 ? w_ot : w_me) * inv;
      bf16* dst = p.o + p.o_col0 + (head0 + g) * HD + half * HH + static_cast<long long>(q_tok) * p.os1 +
                  static_cast<long long>(qc2) * p.os2 +
                  static_cast<long long>((bi / p.q.A) * p.q.mul + q_view) * p.os3;
#pragma unroll
      for (int c = 0; c < HH; c += 16) {
        uint32_t lo[16], hi[16];
        tmem_ld_x16(tO + lane_sel + (g * 2) * HD + half * HH + c, lo);
        tmem_ld_x16(tO + lane_sel + (g * 2 + 1) * HD + half * HH + c, hi);
        tmem_ld_wait();
        if (q_valid) {
#pragma unroll
          for (int e8 = 0; e8 < 16; e8 += 8) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(lo[e8 + e]) * w_lo + __uint_as_float(hi[e8 + e]) * w_hi;
            if (p.accumulate) {
              const uint4 old = *reinterpret_cast<const uint4*>(dst + c + e8);
              float prev[8];
              unpack8(old, prev);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = prev[e] + __bfloat162float(__float2bfloat16(o[e]));
            }
            *reinterpret_cast<uint4*>(dst + c + e8) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                                 pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, C::kTmemCols); }
}

template <int HD, bool BIAS, int G>
static int launch_attn(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const CUtensorMap& b,
                       const AttnParams& p, int heads, int batch, cudaStream_t st) {
  using C = AttnCfg<HD>;
  const int smem = G * C::kQBytes + 4 * C::kKVBytes + C::kPBytes + (BIAS ? C::kBiasBytes : 0) + 256;   // barriers + TMEM slot
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(attention_kernel<HD, BIAS, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return I360_ERR_CUDA;
    attr_set = true;
  }
  launch_k(attention_kernel<HD, BIAS, G>, dim3(p.q_tiles, heads / G, batch), dim3(kAttnThreads), smem, st, q, k, v, b, p);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}


// ================================================================================================
// attention2_kernel: head_dim 64, no bias, whole tiles (Nq % 256 == 0, Nk % 128 == 0) -- the spatial self-attention of
// the 16x512x1024 step (pano 8192 / 2048 / 512 tokens, views of 1024 / 256).  Differences from attention_kernel:
//   * ONE CTA per SM works on TWO 128-row query tiles against one K/V stream (each K/V tile is loaded once for 256
//     query rows instead of once per 128), 576 threads: warp0 TMA, warp1 MMA, warps 2..17 = 2 tiles x 2 column halves
//     x 4 lane quarters of softmax threads (two threads per query row, as before);
//   * P never touches shared memory: a softmax thread overwrites the first 32 of ITS OWN 64 S columns in TMEM with
//     its 64 bf16 probabilities (tcgen05.st) and P.V is issued with the A operand in TMEM (tcgen05.mma ".ts" form) --
//     no 32 KB P tile, no st.shared + address arithmetic, no fence.proxy.async (a MEMBAR.ALL.CTA) per tile, and the P
//     buffer cannot stall the next tile's softmax;
//   * (tile t, column half h) are FOUR independent pipelines S_x -> softmax -> P_x -> O_x += P_x V that share the tensor
//     pipe round-robin: issue order PV_x(j), S_x(j+1) for x = 0..3; the pipe executes in order, so S_x(j+1) cannot
//     overwrite P_x(j) before P_x(j).V has consumed it and no "S released" barrier is needed; while one pipeline waits
//     for its two MMAs the other three are in their softmax (with only two pipelines -- whole tiles -- the softmax warps
//     waited for S half of the time: profiles/r02_ncu_attention2_pano_l0_first.txt).
// TMEM (all 512 columns): S_t / P_t at t*128, O_{t,h} at 256 + (t*2+h)*64.
#ifndef I360_A2_DELAY_S
#define I360_A2_DELAY_S 1
#endif
constexpr int kAttn2Threads = 576;
constexpr int kAttn2Stages = 4;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
         "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
         "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
         "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

#ifndef I360_A2_SPIN
#define I360_A2_SPIN 0
#endif
// hand-offs on the critical path of attention2_kernel (softmax -> MMA -> softmax): optionally a pure test_wait spin
// instead of the suspending try_wait (experiment knob)
__device__ __forceinline__ void mbar_wait_fast(uint64_t* bar, uint32_t parity) {
#if I360_A2_SPIN
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (++spins > I360_SPIN_LIMIT) i360_device_fail("mbarrier test_wait exceeded the spin limit");
  } while (!ok);
#else
  mbar_wait(bar, parity);
#endif
}

__global__ void __launch_bounds__(kAttn2Threads, 1)
attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int HD = 64;
  using C = AttnCfg<HD>;
  constexpr int NS = kAttn2Stages;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) i360_device_fail("dynamic shared memory is not 1024-byte aligned (128B-swizzled TMA / UMMA tiles)");
  uint8_t* sQ = smem;                              // 2 query tiles
  uint8_t* sK = sQ + 2 * C::kQBytes;               // NS stages
  uint8_t* sV = sK + NS * C::kKVBytes;             // NS stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NS * C::kKVBytes);
  uint64_t* q_full = bars;                         // 1
  uint64_t* k_full = bars + 1;                     // NS
  uint64_t* k_empty = k_full + NS;                 // NS
  uint64_t* v_full = k_empty + NS;                 // NS
  uint64_t* v_empty = v_full + NS;                 // NS
  uint64_t* s_full = v_empty + NS;                 // [2 tiles][2]    MMA -> softmax: S_{t,h}(j) in TMEM
  uint64_t* p_full = s_full + 4;                   // [2 tiles][2]    softmax -> MMA: P_{t,h}(j) in TMEM, O_{t,h} rescaled (128 arrivals)
  uint64_t* pv_done = p_full + 4;                  // [2 tiles][2]    MMA -> softmax: O_{t,h} += P V retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 4);
  float2* xml = reinterpret_cast<float2*>(tmem_slot + 2);      // [2 tiles][2 halves][128 rows] (m, l) exchange, 4 KB

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qp = blockIdx.x, head = blockIdx.y, bi = blockIdx.z;
  const int T = p.kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < NS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int x = 0; x < 4; ++x) { mbar_init(&s_full[x], 1); mbar_init(&p_full[x], 128); mbar_init(&pv_done[x], 1); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();
  griddep_launch();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * C::kQBytes);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        int c1, c2, c3;
        tile_coords(p.q, bi, 2 * qp + t, c1, c2, c3);
        tma_load_4d(sQ + t * C::kQBytes, &tmQ, q_full, p.q.col0 + head * HD, c1, c2, c3);
      }
      int st = 0; uint32_t ph = 0;
      for (int j = 0; j < T; ++j) {
        int c1, c2, c3;
        tile_coords(p.kv, bi, j, c1, c2, c3);
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], C::kKVBytes);
        tma_load_4d(sK + st * C::kKVBytes, &tmK, &k_full[st], p.kv.col0 + head * HD, c1, c2, c3);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], C::kKVBytes);
        tma_load_4d(sV + st * C::kKVBytes, &tmV, &v_full[st], p.v_col0 + head * HD, c1, c2, c3);
        if (++st == NS) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);   // B (= V) is MN-major; A (= P) comes from TMEM
      const uint32_t aQ = smem_u32(sQ);
      // The two column halves of a tile are independent online softmaxes, so (tile t, half h) = x is a pipeline of its
      // own: S_x = Q_t K[h*64 .. h*64+64)^T (N = 64) -> softmax -> P_x -> O_x += P_x V[h*64 .. ).  Four pipelines share
      // the tensor pipe round-robin: while one waits for its two MMAs, the other three are in their softmax.
      auto issue_s = [&](int x, uint32_t aK) {
        const int t = x >> 1, h = x & 1;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_bf16_ss(tmem_base + t * 128 + h * 64, make_smem_desc(aQ + t * C::kQBytes + ks * 32, C::kSBO, 16, C::kSwz),
                       make_smem_desc(aK + h * 64 * C::kRowBytes + ks * 32, C::kSBO, 16, C::kSwz), idesc_s, ks != 0);
        umma_commit(&s_full[x]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
#pragma unroll
      for (int x = 0; x < 4; ++x) issue_s(x, smem_u32(sK));
      umma_commit(&k_empty[0]);
      int st = 0; uint32_t ph = 0;                        // stage / phase of kv tile j
      for (int j = 0; j < T; ++j) {
        int stn = st + 1; uint32_t phn = ph;
        if (stn == NS) { stn = 0; phn ^= 1; }
        const bool more = j + 1 < T;
        mbar_wait(&v_full[st], ph);
        if (more) mbar_wait(&k_full[stn], phn);
        const uint32_t aV = smem_u32(sV + st * C::kKVBytes);
        const uint32_t aKn = smem_u32(sK + stn * C::kKVBytes);
#pragma unroll
        for (int x = 0; x < 4; ++x) {                     // O_x += P_x(j) V_j[h*64 .. h*64+64, :]
          const int t = x >> 1, h = x & 1;
          mbar_wait_fast(&p_full[x], j & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(tmem_base + 256 + x * HD, tmem_base + t * 128 + h * 64 + kk * 8,
                         make_smem_desc(aV + (h * 4 + kk) * 16 * C::kRowBytes, C::kSBO, 16, C::kSwz), idesc_o,
                         (j > 0) || (kk != 0));
          umma_commit(&pv_done[x]);
          // S_x(j+1) overwrites the TMEM columns P_x(j).V reads: issued back to back the pipe drains between the two
          // (measured: ~700 idle clocks per pair).  One independent P.V is slotted in between.
#if I360_A2_DELAY_S
          if (more && x >= 1) issue_s(x - 1, aKn);
#else
          if (more) issue_s(x, aKn);                      // in-order pipe: runs after P_x(j) has been consumed
#endif
        }
#if I360_A2_DELAY_S
        if (more) issue_s(3, aKn);
#endif
        umma_commit(&v_empty[st]);
        if (more) umma_commit(&k_empty[stn]);
        st = stn; ph = phn;
      }
    }
  } else {
    // ================================ softmax / epilogue ================================
    const int idx = warp - 2;
    const int t = idx >> 3;                        // query tile 0 / 1
    const int half = (idx >> 2) & 1;               // key columns half*64 .. half*64+63 of every tile
    const int ew = warp & 3;                       // TMEM lane quarter of this warp
    const int row = ew * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
    constexpr int HH = HD / 2;
    const int qt = 2 * qp + t;
    const int q_tok = (qt % p.q.n1) * p.q.box1 + row;
    int qc1, qc2, qc3;
    tile_coords(p.q, bi, qt, qc1, qc2, qc3);
    const uint32_t tS_mine = tmem_base + lane_sel + t * 128 + half * 64;
    const uint32_t tO_mine = tmem_base + lane_sel + 256 + (t * 2 + half) * HD;
    uint64_t* my_p_full = &p_full[t * 2 + half];
    uint64_t* my_pv_done = &pv_done[t * 2 + half];
    float m_used = -INFINITY, l_run = 0.f;
    const float scale_log2 = p.scale_log2;
    const float2 sc2 = make_float2(scale_log2, scale_log2);

    for (int j = 0; j < T; ++j) {
      mbar_wait_fast(&s_full[t * 2 + half], j & 1);
      tc_fence_after();
      uint32_t v[64];
      tmem_ld_x32(tS_mine, v);
      tmem_ld_x32(tS_mine + 32, v + 32);
      tmem_ld_wait();
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int e = 0; e < 64; e += 8) {
#pragma unroll
        for (int c = 0; c < 4; ++c) mx4[c] = fmax3(mx4[c], __uint_as_float(v[e + 2 * c]), __uint_as_float(v[e + 2 * c + 1]));
      }
      const float mx = fmax3(mx4[0], mx4[1], fmaxf(mx4[2], mx4[3])) * scale_log2;
      const float m_new = fmaxf(m_used, mx);
      const bool need = m_new > m_used + 8.0f;       // lazy rescale (see softmax_tile)
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = (m_new == m_used) ? 1.0f : fast_exp2(m_used - m_new);
        if (j > 0) {
          mbar_wait(my_pv_done, (j - 1) & 1);        // P(j-1) V(j-1) has landed in the accumulator
          tc_fence_after();
          const float2 al2 = make_float2(alpha, alpha);
#pragma unroll 1
          for (int c = 0; c < HD; c += 16) {
            uint32_t o[16];
            tmem_ld_x16(tO_mine + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const float2 r = fmul2(make_float2(__uint_as_float(o[e]), __uint_as_float(o[e + 1])), al2);
              o[e] = __float_as_uint(r.x); o[e + 1] = __float_as_uint(r.y);
            }
            tmem_st_x16(tO_mine + c, o);
          }
        }
        l_run *= alpha;
        m_used = m_new;
      }
      const float2 nm2 = make_float2(-m_used, -m_used);
      float2 sum2 = make_float2(0.f, 0.f);
      uint32_t pk[32];
#pragma unroll
      for (int e = 0; e < 64; e += 2) {
        const float2 tt = ffma2(make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), sc2, nm2);
        const bool poly = (I360_POLY_MASK2 >> ((e >> 1) & 7)) & 1;
        const float2 pe = poly ? exp2_poly2(tt) : make_float2(fast_exp2(tt.x), fast_exp2(tt.y));
        sum2 = fadd2(sum2, pe);
        pk[e >> 1] = pack_bf16x2(pe.x, pe.y);
      }
      l_run += sum2.x + sum2.y;
      tmem_st_x32(tS_mine, pk);                      // P over the first half of my own S columns
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(my_p_full);
    }
    // ---- epilogue: merge the two column halves of my row, write my half of the head-dim columns ----
    const uint32_t lastp = (T - 1) & 1;
    mbar_wait(&pv_done[t * 2], lastp);
    mbar_wait(&pv_done[t * 2 + 1], lastp);
    tc_fence_after();
    xml[(t * 2 + half) * 128 + row] = make_float2(m_used, l_run);
    named_bar_sync(1 + t, 256);
    const float2 oth = xml[(t * 2 + (half ^ 1)) * 128 + row];
    const float m_all = fmaxf(m_used, oth.x);
    const float w_me = fast_exp2(m_used - m_all), w_ot = fast_exp2(oth.x - m_all);
    const float inv = 1.0f / (l_run * w_me + oth.y * w_ot);
    const float w_lo = (half == 0 ? w_me : w_ot) * inv, w_hi = (half == 0 ? w_ot : w_me) * inv;
    bf16* dst = p.o + p.o_col0 + head * HD + half * HH + static_cast<long long>(q_tok) * p.os1 +
                static_cast<long long>(qc2) * p.os2 + static_cast<long long>(qc3) * p.os3;
    const uint32_t tO_t = tmem_base + lane_sel + 256 + (t * 2) * HD;
#pragma unroll
    for (int c = 0; c < HH; c += 16) {
      uint32_t lo[16], hi[16];
      tmem_ld_x16(tO_t + half * HH + c, lo);
      tmem_ld_x16(tO_t + HD + half * HH + c, hi);
      tmem_ld_wait();
#pragma unroll
      for (int e8 = 0; e8 < 16; e8 += 8) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(lo[e8 + e]) * w_lo + __uint_as_float(hi[e8 + e]) * w_hi;
        *reinterpret_cast<uint4*>(dst + c + e8) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                             pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static int launch_attn2(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const AttnParams& p, int heads,
                        int batch, cudaStream_t st) {
  using C = AttnCfg<64>;
  const int smem = 2 * C::kQBytes + 2 * kAttn2Stages * C::kKVBytes + 512 + 4096;   // barriers + TMEM slot + (m, l) exchange
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(attention2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return I360_ERR_CUDA;
    attr_set = true;
  }
  launch_k(attention2_kernel, dim3(p.q_tiles / 2, heads, batch), dim3(kAttn2Threads), smem, st, q, k, v, p);
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
static int attention_impl(const I360TokenView* q, const I360TokenView* k, const I360TokenView* v,
                          const I360TokenView* o, int heads, int head_dim, int batch, float scale,
                          const void* bias, int bias_rows, int bias_cols, int bias_item_rows, int accumulate, void* stream) {
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
  p.bias_item_rows = bias ? bias_item_rows : 0;
  if (p.bias_item_rows && (head_dim != 64 || p.q.box3 != 1 || p.kv.box3 != 1)) return I360_ERR_UNSUPPORTED;
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
    if (head_dim == 64) return launch_attn<64, true, 1>(tq, tk, tv, tb, p, heads, batch, st);
    static const bool pair = getenv("I360_WARP_HEADS_PER_CTA") == nullptr || atoi(getenv("I360_WARP_HEADS_PER_CTA")) != 1;
    if (pair && (heads % 2) == 0) return launch_attn<32, true, 2>(tq, tk, tv, tb, p, heads, batch, st);   // WarpAttn
    return launch_attn<32, true, 1>(tq, tk, tv, tb, p, heads, batch, st);
  }
  if (head_dim == 64) {
    // whole-tile self-attention shapes: two query tiles per CTA, P kept in TMEM (attention2_kernel).  EXPERIMENT, off by
    // default: correct (same test results as attention_kernel) and 28 % fewer instructions per tile, but slower on the
    // B200 -- 3.89 / 4.41 ms (two / four pipelines) against 3.50 ms for the 32 x 8192-token panorama level: its softmax
    // warps wait for S 50-66 % of the time (profiles/r02_ncu_attention2_*.txt).  I360_ATTN_V2=1 selects it.
    static const bool v2 = getenv("I360_ATTN_V2") != nullptr && atoi(getenv("I360_ATTN_V2")) != 0;
    if (v2 && !accumulate && p.q.box3 == 1 && p.kv.box3 == 1 && p.q.ext3 == 1 && p.kv.ext3 == 1 && (p.q.d1 % 256) == 0 &&
        (p.kv.d1 % 128) == 0)
      return launch_attn2(tq, tk, tv, p, heads, batch, st);
    return launch_attn<64, false, 1>(tq, tk, tv, tb, p, heads, batch, st);
  }
  return launch_attn<32, false, 1>(tq, tk, tv, tb, p, heads, batch, st);
}

extern "C" int i360_attention_bf16(const I360TokenView* q, const I360TokenView* k, const I360TokenView* v,
                                   const I360TokenView* o, int heads, int head_dim, int batch, float scale,
                                   const void* bias, int bias_rows, int bias_cols, int accumulate, void* stream) {
  return attention_impl(q, k, v, o, heads, head_dim, batch, scale, bias, bias_rows, bias_cols, 0, accumulate, stream);
}

// Same kernel with one bias block per (batch item, head): bias is [batch * heads * bias_item_rows, bias_cols] bf16, block
// (bi * heads + h) holding the [Nq, Nk] logits bias of that head (rows past Nq of the last query tile run into the next
// block and are never used).  Plain sequences and head_dim 64 only.  Replaces the `attn + rel_h + rel_w` -> softmax ->
// `attn @ v` sequence of segment_anything's ImageEncoderViT Attention.forward (modeling/image_encoder.py, release 1.0).
extern "C" int i360_attention_item_bias_bf16(const I360TokenView* q, const I360TokenView* k, const I360TokenView* v,
                                             const I360TokenView* o, int heads, int head_dim, int batch, float scale,
                                             const void* bias, int bias_item_rows, int bias_cols, void* stream) {
  if (!bias || bias_item_rows <= 0) return I360_ERR_ARG;
  const long long rows = static_cast<long long>(batch) * heads * bias_item_rows;
  if (rows > 0x7fffffffLL) return I360_ERR_ARG;
  return attention_impl(q, k, v, o, heads, head_dim, batch, scale, bias, static_cast<int>(rows), bias_cols, bias_item_rows, 0,
                        stream);
}
