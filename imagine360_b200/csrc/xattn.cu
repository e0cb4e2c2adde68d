// Fused text + image-prompt cross-attention with STATIONARY keys/values (sm_100a, head_dim 64).
//
// Replaces the two attention calls and the bf16 sum of IPCrossAttention (animatediff/models/attention.py:65-156:
// text branch :119-131, image branch :133-146, sum :148) on the xformers path the production run takes.  Both key
// sets are short (77 text tokens, 64 image tokens) and shared by every frame of a clip element, so the generic
// flash kernel (attention.cu) degenerates into one latency chain per (128 queries, head) -- ~6 us per item for
// 0.3 us of HBM traffic.  Here a persistent CTA keeps K = [K_text; K_ip] and V = [V_text; V_ip] of one
// (clip element, head) in shared memory and streams that element's query rows (all frames: they are contiguous
// rows of the [tokens, C] projection output) through a 3-deep software pipeline:
//   warp0  TMA producer : K/V once per (element, head) segment, Q tiles through a 4-stage mbarrier ring
//   warp1  MMA issuer   : S_i = Q_i [K_t; K_ip]^T -> TMEM (double buffered, one N = ntp+nip instruction per k-step);
//                         O_i = P_i [V_t; V_ip] -> TMEM (double buffered)
//   warps 2..9          : TWO threads per query row: one owns the text logits, one the image logits -- the two
//                         softmaxes are independent by construction.  The row is complete in registers (single key
//                         tile), so P is normalised BEFORE it is rounded to bf16 and both branches accumulate into
//                         ONE fp32 accumulator: out = softmax(qK_t^T)V_t + softmax(qK_ip^T)V_ip with a single final
//                         rounding (the reference rounds each branch and the sum to bf16: strictly fewer roundings
//                         here).  The epilogue of tile i-1 (TMEM -> bf16 -> global) runs after the softmax of tile i,
//                         under the tensor core's P_i V.
// Work is the flat list of (element, head, query tile) items cut into equal contiguous ranges, one per SM, so there
// is no wave quantisation; a CTA reloads K/V when its range crosses into the next (element, head).
// Per item the kernel moves 16 KB of Q in and 16 KB of O out: it is HBM bound (algorithmic bytes = 2 * rows * C * 2).
#include "common.cuh"
#include "tmap.h"

namespace i360 {

#ifndef I360_XPOLY_MASK
#define I360_XPOLY_MASK 0x00   // which of every 8 logit pairs take the polynomial exp2 (bit i = pair i).  None: this
                               // kernel is latency / issue bound, not MUFU bound (0x52: 0.298 ms, 0x22: 0.282, 0x00: 0.277)
#endif
constexpr int kXThreads = 320;
constexpr int kXQStages = 4;
constexpr int kXKVBytes = 192 * 128;        // K (and V) rows of both branches, 128 B each, 128B-swizzled
constexpr int kXQBytes = 128 * 128;
constexpr int kXPBlock = 128 * 128;         // one [128 queries x 64 keys] bf16 block, K-major, 128B-swizzled
constexpr int kXPBuf = 3 * kXPBlock;
constexpr int kXSmem = 2 * kXKVBytes + kXQStages * kXQBytes + 2 * kXPBuf + 256;
constexpr int kXTmemCols = 512;             // S0 @0, S1 @192, O0 @384, O1 @448

struct XAttnParams {
  int nt, ni;                  // valid text / image keys
  int ntp, nip;                // padded to multiples of 16 (MMA N / K granularity)
  int nbt;                     // P blocks of the text branch = ceil(ntp / 64)
  int rows_per_ctx;            // query rows of one clip element (frames * tokens)
  int tiles_per_ctx;           // ceil(rows_per_ctx / 128)
  int heads, c;
  long long total_tiles;       // n_ctx * heads * tiles_per_ctx
  float scale_log2;
  bf16* o; long long ldo;
};

struct XSeg { int ctx, head, t0, t1; };

// next contiguous run of query tiles that share one (element, head); f advances through [f, f1)
__device__ __forceinline__ bool x_next_seg(long long& f, long long f1, const XAttnParams& p, XSeg& s) {
  if (f >= f1) return false;
  const long long pair = f / p.tiles_per_ctx;
  const int tb = static_cast<int>(f - pair * p.tiles_per_ctx);
  const long long left = f1 - f;
  const int n = static_cast<int>(left < (p.tiles_per_ctx - tb) ? left : (p.tiles_per_ctx - tb));
  s.ctx = static_cast<int>(pair / p.heads); s.head = static_cast<int>(pair % p.heads);
  s.t0 = tb; s.t1 = tb + n;
  f += n;
  return true;
}

// Softmax of one query row over one branch's NCH*16 (padded) key columns; P / rowsum -> bf16 -> swizzled smem.
template <int NCH>
__device__ __forceinline__ void x_softmax(uint32_t tS_mine, uint8_t* sPb, int row, int nvalid, float scale_log2,
                                          uint64_t* s_empty) {
  constexpr int NC = NCH * 16;
  uint32_t v[NC];
#pragma unroll
  for (int c = 0; c < NCH; ++c) tmem_ld_x16(tS_mine + c * 16, v + c * 16);
  tmem_ld_wait();
  tc_fence_before();
  mbar_arrive(s_empty);                               // S_i is in registers: the buffer may be overwritten
#pragma unroll
  for (int e = NC - 16; e < NC; ++e)
    if (e >= nvalid) v[e] = 0xff800000u;              // padded key columns: -inf
  float mx = -INFINITY;
#pragma unroll
  for (int e = 0; e < NC; e += 2) mx = fmax3(mx, __uint_as_float(v[e]), __uint_as_float(v[e + 1]));
  const float2 sc2 = make_float2(scale_log2, scale_log2);
  const float2 nm2 = make_float2(-mx * scale_log2, -mx * scale_log2);
  float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < NC; e += 2) {
    const float2 t = ffma2(make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), sc2, nm2);
    // optionally some pairs on the FMA pipe (polynomial) instead of the MUFU; never in the chunk that holds -inf
    const int pi = (e >> 1) & 7;
    const bool poly = (e < NC - 16) && ((I360_XPOLY_MASK >> pi) & 1);
    const float2 pe = poly ? exp2_poly2(t) : make_float2(fast_exp2(t.x), fast_exp2(t.y));
    sum2 = fadd2(sum2, pe);
    v[e] = __float_as_uint(pe.x); v[e + 1] = __float_as_uint(pe.y);
  }
  const float inv = 1.0f / (sum2.x + sum2.y);
  const float2 inv2 = make_float2(inv, inv);
  const uint32_t rsw = static_cast<uint32_t>(row & 7);
#pragma unroll
  for (int g = 0; g < NC / 8; ++g) {
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 r = fmul2(make_float2(__uint_as_float(v[g * 8 + 2 * e]), __uint_as_float(v[g * 8 + 2 * e + 1])), inv2);
      pk[e] = pack_bf16x2(r.x, r.y);
    }
    const int cc = g * 8;
    *reinterpret_cast<uint4*>(sPb + (cc >> 6) * kXPBlock + row * 128 + ((((cc & 63) >> 3) ^ rsw) << 4)) =
        make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

__global__ void __launch_bounds__(kXThreads, 1)
xattn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKt,
             const __grid_constant__ CUtensorMap tmKi, const XAttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) i360_device_fail("dynamic shared memory is not 1024-byte aligned (128B-swizzled TMA / UMMA tiles)");
  uint8_t* sK = smem;
  uint8_t* sV = sK + kXKVBytes;
  uint8_t* sQ = sV + kXKVBytes;
  uint8_t* sP = sQ + kXQStages * kXQBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kXPBuf);
  uint64_t* kv_full = bars;            // TMA -> MMA: K/V of the segment landed
  uint64_t* kv_empty = bars + 1;       // MMA -> TMA: every MMA of the segment retired
  uint64_t* q_full = bars + 2;         // [4]
  uint64_t* q_empty = bars + 6;        // [4]
  uint64_t* s_full = bars + 10;        // [2] MMA -> softmax
  uint64_t* s_empty = bars + 12;       // [2] softmax -> MMA (256 arrivals)
  uint64_t* p_full = bars + 14;        // [2] softmax -> MMA (256 arrivals)
  uint64_t* o_full = bars + 16;        // [2] MMA -> epilogue (also: P buffer free)
  uint64_t* o_empty = bars + 18;       // [2] epilogue -> MMA (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long fbeg = p.total_tiles * blockIdx.x / gridDim.x;
  const long long fend = p.total_tiles * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKt); tma_prefetch_desc(&tmKi);
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    for (int s = 0; s < kXQStages; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 256); mbar_init(&p_full[s], 256);
      mbar_init(&o_full[s], 1); mbar_init(&o_empty[s], 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, kXTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();
  griddep_launch();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      long long f = fbeg; XSeg sg; int i = 0, g = 0;
      while (x_next_seg(f, fend, p, sg)) {
        mbar_wait(kv_empty, (g & 1) ^ 1);            // the previous segment's MMAs no longer read K/V
        mbar_expect_tx(kv_full, 2 * (p.ntp + p.nip) * 128);
        tma_load_2d(sK, &tmKt, kv_full, sg.head * 64, sg.ctx * p.nt);
        tma_load_2d(sK + p.ntp * 128, &tmKi, kv_full, sg.head * 64, sg.ctx * p.ni);
        tma_load_2d(sV, &tmKt, kv_full, p.c + sg.head * 64, sg.ctx * p.nt);
        tma_load_2d(sV + p.ntp * 128, &tmKi, kv_full, p.c + sg.head * 64, sg.ctx * p.ni);
        for (int t = sg.t0; t < sg.t1; ++t, ++i) {
          const int st = i % kXQStages; const uint32_t ph = (i / kXQStages) & 1;
          mbar_wait(&q_empty[st], ph ^ 1);
          mbar_expect_tx(&q_full[st], kXQBytes);
          tma_load_2d(sQ + st * kXQBytes, &tmQ, &q_full[st], sg.head * 64, sg.ctx * p.rows_per_ctx + t * 128);
        }
        ++g;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, p.ntp + p.nip, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);    // B (= V) is MN-major
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
      const int kt = p.ntp >> 4, ki = p.nip >> 4;
      auto issue_s = [&](int i) {
        const int st = i % kXQStages; const uint32_t ph = (i / kXQStages) & 1;
        const int b = i & 1; const uint32_t u = (i >> 1) & 1;
        mbar_wait(&s_empty[b], u ^ 1);
        mbar_wait(&q_full[st], ph);
        tc_fence_after();
        const uint32_t aQ = smem_u32(sQ + st * kXQBytes);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16_ss(tmem_base + b * 192, make_smem_desc(aQ + ks * 32, 1024, 16, SWZ_128B),
                       make_smem_desc(aK + ks * 32, 1024, 16, SWZ_128B), idesc_s, ks != 0);
        umma_commit(&q_empty[st]);
        umma_commit(&s_full[b]);
      };
      long long f = fbeg; XSeg sg; int i = 0, g = 0;
      while (x_next_seg(f, fend, p, sg)) {
        mbar_wait(kv_full, g & 1);
        issue_s(i);
        for (int t = sg.t0; t < sg.t1; ++t, ++i) {
          if (t + 1 < sg.t1) issue_s(i + 1);         // runs under the softmax of tile i
          const int b = i & 1; const uint32_t u = (i >> 1) & 1;
          mbar_wait(&o_empty[b], u ^ 1);
          mbar_wait(&p_full[b], u);
          tc_fence_after();
          const uint32_t aP = smem_u32(sP + b * kXPBuf), tO = tmem_base + 384 + b * 64;
          for (int kk = 0; kk < kt; ++kk)
            umma_bf16_ss(tO, make_smem_desc(aP + (kk >> 2) * kXPBlock + (kk & 3) * 32, 1024, 16, SWZ_128B),
                         make_smem_desc(aV + kk * 16 * 128, 1024, 16, SWZ_128B), idesc_o, kk != 0);
          for (int kk = 0; kk < ki; ++kk)
            umma_bf16_ss(tO, make_smem_desc(aP + (p.nbt + (kk >> 2)) * kXPBlock + (kk & 3) * 32, 1024, 16, SWZ_128B),
                         make_smem_desc(aV + (p.ntp + kk * 16) * 128, 1024, 16, SWZ_128B), idesc_o, 1);
          umma_commit(&o_full[b]);
        }
        umma_commit(kv_empty);
        ++g;
      }
    }
  } else {
    const int ew = warp & 3;
    const int branch = (warp - 2) >> 2;              // 0: text keys, 1: image-prompt keys
    const int row = ew * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
    const int nch = (branch ? p.nip : p.ntp) >> 4;
    const int nvalid = branch ? p.ni : p.nt;
    const uint32_t s_col = branch ? p.ntp : 0;
    const int p_blk = branch ? p.nbt : 0;

    auto epilogue = [&](int j, long long grow, int head, bool valid) {
      const int b = j & 1; const uint32_t u = (j >> 1) & 1;
      mbar_wait(&o_full[b], u);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld_x32(tmem_base + 384 + b * 64 + lane_sel + branch * 32, o);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&o_empty[b]);
      if (valid) {
        bf16* dst = p.o + grow * p.ldo + head * 64 + branch * 32;
#pragma unroll
        for (int g = 0; g < 32; g += 8)
          *reinterpret_cast<uint4*>(dst + g) =
              make_uint4(pack_bf16x2(__uint_as_float(o[g]), __uint_as_float(o[g + 1])),
                         pack_bf16x2(__uint_as_float(o[g + 2]), __uint_as_float(o[g + 3])),
                         pack_bf16x2(__uint_as_float(o[g + 4]), __uint_as_float(o[g + 5])),
                         pack_bf16x2(__uint_as_float(o[g + 6]), __uint_as_float(o[g + 7])));
      }
    };

    long long f = fbeg; XSeg sg; int i = 0;
    long long prev_row = 0; int prev_head = 0; bool prev_valid = false;
    while (x_next_seg(f, fend, p, sg)) {
      for (int t = sg.t0; t < sg.t1; ++t, ++i) {
        const int b = i & 1; const uint32_t u = (i >> 1) & 1;
        mbar_wait(&s_full[b], u);
        tc_fence_after();
        const uint32_t tS_mine = tmem_base + b * 192 + lane_sel + s_col;
        uint8_t* sPb = sP + b * kXPBuf + p_blk * kXPBlock;   // free: o_full of tile i-2 was waited on in its epilogue
        switch (nch) {
          case 1: x_softmax<1>(tS_mine, sPb, row, nvalid, p.scale_log2, &s_empty[b]); break;
          case 2: x_softmax<2>(tS_mine, sPb, row, nvalid, p.scale_log2, &s_empty[b]); break;
          case 3: x_softmax<3>(tS_mine, sPb, row, nvalid, p.scale_log2, &s_empty[b]); break;
          case 4: x_softmax<4>(tS_mine, sPb, row, nvalid, p.scale_log2, &s_empty[b]); break;
          case 5: x_softmax<5>(tS_mine, sPb, row, nvalid, p.scale_log2, &s_empty[b]); break;
          default: x_softmax<6>(tS_mine, sPb, row, nvalid, p.scale_log2, &s_empty[b]); break;
        }
        fence_proxy_async_smem();                     // P visible to the tensor core (async proxy)
        mbar_arrive(&p_full[b]);
        if (i > 0) epilogue(i - 1, prev_row, prev_head, prev_valid);
        const int r_in = t * 128 + row;
        prev_valid = r_in < p.rows_per_ctx;
        prev_row = static_cast<long long>(sg.ctx) * p.rows_per_ctx + r_in;
        prev_head = sg.head;
      }
    }
    if (i > 0) epilogue(i - 1, prev_row, prev_head, prev_valid);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kXTmemCols); }
}

}  // namespace i360

using namespace i360;

// q, o: [rows, heads*64] bf16 matrices (row strides ldq / ldo elements); rows = n_ctx * rows_per_ctx, the rows of
// clip element e are [e * rows_per_ctx, (e+1) * rows_per_ctx).  kv_text: [n_ctx * nt, 2*heads*64] = [K | V] projections
// of the text tokens, kv_ip likewise for the image-prompt tokens.
extern "C" int i360_cross_attention_text_ip_bf16(const void* q, long long ldq, void* o, long long ldo, long long rows,
                                                 int n_ctx, const void* kv_text, long long ld_kvt, int nt,
                                                 const void* kv_ip, long long ld_kvi, int ni, int heads, int head_dim,
                                                 float scale, void* stream) {
  if (!q || !o || !kv_text || !kv_ip) return I360_ERR_ARG;
  if (head_dim != 64) return I360_ERR_UNSUPPORTED;
  if (heads <= 0 || n_ctx <= 0 || rows <= 0 || (rows % n_ctx) || nt <= 0 || ni <= 0) return I360_ERR_ARG;
  if ((ldq % 8) || (ldo % 8) || (ld_kvt % 8) || (ld_kvi % 8)) return I360_ERR_ARG;
  XAttnParams p;
  memset(&p, 0, sizeof(p));
  p.nt = nt; p.ni = ni;
  p.ntp = (nt + 15) & ~15; p.nip = (ni + 15) & ~15;
  p.nbt = (p.ntp + 63) / 64;
  const int nbi = (p.nip + 63) / 64;
  // each branch's logits row lives in one thread's registers: 96 columns keep the kernel free of local-memory spills
  if (p.ntp > 96 || p.nip > 96 || p.nbt + nbi > 3) return I360_ERR_UNSUPPORTED;
  if (rows / n_ctx > 0x7fffffff / 2) return I360_ERR_ARG;
  p.rows_per_ctx = static_cast<int>(rows / n_ctx);
  p.tiles_per_ctx = (p.rows_per_ctx + 127) / 128;
  p.heads = heads; p.c = heads * 64;
  p.total_tiles = static_cast<long long>(n_ctx) * heads * p.tiles_per_ctx;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.o = static_cast<bf16*>(o); p.ldo = ldo;
  CUtensorMap tq, tkt, tki;
  {
    uint64_t d[2] = {(uint64_t)p.c, (uint64_t)rows}; uint64_t s[1] = {(uint64_t)ldq * 2}; uint32_t b[2] = {64, 128};
    int r = get_tmap_bf16(&tq, q, 2, d, s, b, 3); if (r) return r;
  }
  {
    uint64_t d[2] = {(uint64_t)2 * p.c, (uint64_t)n_ctx * nt}; uint64_t s[1] = {(uint64_t)ld_kvt * 2};
    uint32_t b[2] = {64, (uint32_t)p.ntp};
    int r = get_tmap_bf16(&tkt, kv_text, 2, d, s, b, 3); if (r) return r;
  }
  {
    uint64_t d[2] = {(uint64_t)2 * p.c, (uint64_t)n_ctx * ni}; uint64_t s[1] = {(uint64_t)ld_kvi * 2};
    uint32_t b[2] = {64, (uint32_t)p.nip};
    int r = get_tmap_bf16(&tki, kv_ip, 2, d, s, b, 3); if (r) return r;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(xattn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kXSmem) != cudaSuccess)
      return I360_ERR_CUDA;
    attr_set = true;
  }
  const long long sms = num_sms();
  const int grid = static_cast<int>(p.total_tiles < sms ? p.total_tiles : sms);
  launch_k(xattn_kernel, dim3(grid), dim3(kXThreads), kXSmem, static_cast<cudaStream_t>(stream), tq, tkt, tki, p);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}
