// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] * W[N,K]^T )           bf16 in, fp32 accumulate in TMEM, bf16 out
//
// One kernel, two A-operand producers:
//   GEMM mode : A is a row-major [M,K] matrix, tiles fetched with 2-D TMA boxes {64, 128}.
//   CONV mode : A is an NHWC activation [B,H,W,C]; the 128 rows of an M tile are a pixel box
//               {TB,TH,TW}; for each of the 9 taps the producer issues a 4-D TMA box shifted by
//               (kh-1, kw-1). Out-of-bounds rows/columns are zero-filled by the TMA unit, which is
//               exactly conv2d's zero padding -> no im2col buffer is ever materialised.
//               Up to two extra 1x1 sources are accumulated into the same tile (ResnetBlock's
//               conv_shortcut over the un-normalised input / the skip-concat halves).
//   HALO mode : CONV with a 16 x 8 pixel tile whose (16+2) x (8+2) halo is loaded ONCE per 64-channel block (one 4-D
//               TMA box, 180 rows of 128 B); the nine taps are nine tcgen05.mma groups whose A descriptor starts
//               (kh * 10 + kw) rows into that tile with SBO = 10 rows (1280 B): each 8-row core group of the operand
//               is one 8-pixel image row of the tile, and the tensor core applies the 128-byte swizzle to the absolute
//               shared-memory address (probed: tools/probe_halo_desc.cu), so shifted starts read exactly what TMA
//               wrote.  Activation traffic from L2 drops from 9 x 16 KB to 22.5 KB per channel block; the weight
//               tiles (one per tap and channel block) flow through their own ring.
//
// Warp roles (320 threads): warp0 = TMA producer, warp1 = TMEM alloc + MMA issuer (one elected
// lane), warps 2..9 = two epilogue groups (TMEM -> registers -> bias/temb/residual/GEGLU -> bf16 ->
// swizzled smem -> TMA tensor store), each group taking every other 64-column chunk of the tile.
// Two TMEM accumulator stages let the epilogue of tile i overlap the mainloop of tile i+1.
//
// Replaces the reference's library calls: nn.Conv2d inside InflatedConv3d
// (animatediff/models/resnet.py:19-27), nn.Linear / LoRACompatibleLinear everywhere
// (diffusers/models/lora.py:368), GEGLU (diffusers/models/activations.py:93-122).
#include "common.cuh"
#include "tmap.h"
#include <stdio.h>
#include <stdlib.h>

namespace i360 {

struct GemmConvParams {
  // problem
  int M, N, K;        // GEMM mode (conv: M = padded pixel count, K unused)
  int conv;           // 0 = GEMM, 1 = CONV3x3
  // conv geometry (input dims include any materialised halo)
  int B, H, W;
  int TW, TH, TB;     // pixel box, TW*TH*TB == 128
  int n_wt, n_ht, n_bt;
  int Cin, C2, C3;    // channels of the 3x3 source and of up to two extra 1x1 sources
  int halo;           // CONV through the halo-tile kernels (TW = 8, TH = 16, TB = 1)
  // CONV tap list (tap-by-tap mode): n_taps shifted boxes (dh, dw) per channel block, weights tap-major.  3x3: the nine
  // (kh-1, kw-1); the sub-pixel form of nearest-upsample + 3x3 has four taps per output parity (i360_conv_upsample2x_bf16)
  int n_taps; signed char tap_dh[9]; signed char tap_dw[9];
  // direct (cropped) stores into a strided output lattice: element offsets of one step in w / h / image (0 = dense)
  long long os_w, os_h, os_b;
  // GroupNorm statistics of the OUTPUT, accumulated by the epilogue (CONV, one image per tile, group size 4 / 8 / 16 so that
  // groups never straddle a 32-column chunk): gn_stats[(image * 32 + group) * 2 + {0, 1}] += (sum, sum of squares), fp64
  double* gn_stats; int gn_gs;
  // ... or, for ANY group size (the UNet's 10 / 20 / 40 channels per group straddle the 32-column chunks): per-CHANNEL statistics
  // gn_chan[(image * N + channel) * 2 + {0, 1}] += (sum, sum of squares) of the stored bf16 values, fp64 atomics of fp32 warp sums (one coalesced
  // 32-lane request per warp and chunk after a recursive-halving reduction over the warp's 32 pixel rows); the consuming
  // GroupNorm folds channels into groups (i360_groupnorm_apply_chanstats)
  double* gn_chan;
  int in_stride;      // CONV: input pixel = in_stride * output pixel + tap offset (2: the stride-2 downsample convs, whose
                      // activation boxes are TMA boxes with traversal stride 2 -- no im2col buffer)
  int crop;           // output columns cropped on each side (pano halo)
  int xoff;           // column offset applied to the extra 1x1 sources (= -crop: they are stored un-padded)
  int Hout, Wout;
  // tiling
  int m_tiles, n_tiles;
  int k_iters;        // 64-wide K blocks per tile
  // epilogue
  bf16* D; long long ldd;
  const bf16* bias;                 // [N] (GEGLU: [2*N_out], values then gates like the weight)
  const bf16* resid; long long ldr; // same row mapping as D
  const float* rowvec; int rowvec_div; int rowvec_ld;  // += rowvec[(img_or_row / div) * ld + col]
  int act;            // 0 none, 1 GEGLU (tile = [values | gates]), 2 GELU(erf), 3 SiLU
  float out_scale;
  int n_out;          // logical output columns (GEGLU: N/2)
  int rowvec_mod;     // > 0: rowvec row = (row / rowvec_div) % rowvec_mod  (temporal PE: frame of a (b, f, d) row)
  int rowvec_in_table; // LNF: every row of a tile maps to the same rowvec row (rowvec_div % 128 == 0): the vector is folded
                      // into the tile's smem bias table instead of being re-read by every thread for every chunk
  // LayerNorm folded into the GEMM (LNF kernels): D = rstd_r * (A W'^T - mean_r * u) + c, W' = W * gamma (per column
  // of K), u[n] = sum_k W'[n,k], c[n] = sum_k beta[k] W[n,k] + bias[n]; the row statistics come from the GEMM that
  // PRODUCED A (its epilogue wrote partial sums per row, see rowstats)
  const float* ln_u; const float* ln_c; float ln_eps;
  const float2* ln_stats; int ln_slots;   // [ln_slots][M] partial (sum, sum of squares) of every A row, written by the producer
  // producer side of the fold: this GEMM also writes, per output row, the (sum, sum of squares) of the values it stores,
  // one partial per (N tile, epilogue group) slot: rowstats[slot * M + row]
  float2* rowstats;
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 320;   // warp0 TMA, warp1 MMA, warps 2..9 epilogue (two groups of 4)
constexpr int kMaxStatSlots = 10; // row-statistics slots a LayerNorm-folding GEMM can consume (5 N tiles x 2 epilogue groups)

// MT: M sub-tiles per CTA tile (1 or 2).  MT = 2 (GEMM mode, BN = 128 only) computes a 256 x 128 tile as two 128-row
// accumulators that share the weight tile in smem: 683 instead of 569 FLOP per byte loaded from L2 for the N = 640 /
// 1920 projections, which are bound by L2 -> SM traffic at the tensor-core pace; epilogue group g owns accumulator g.
constexpr int kHaloTW = 8, kHaloTH = 16;                       // pixel tile of HALO mode (8-pixel rows = UMMA core groups)
constexpr int kHaloW = kHaloTW + 2, kHaloH = kHaloTH + 2;      // loaded box
constexpr int kHaloRows = kHaloW * kHaloH;                     // 180 rows of 128 bytes
constexpr int kHaloTxBytes = kHaloRows * BK * 2;               // 23 040
constexpr int kHaloBytes = 23 * 1024;                          // ... rounded up to the 1024-byte swizzle atom
constexpr int kHaloStages = 2;                                 // a halo tile lives for 9 taps: double buffering suffices

template <int BN, bool RING = false, int MT = 1, bool LNF = false, bool HALO = false, int NG = 2> struct Cfg {
  static_assert(!HALO || (MT == 1 && !LNF), "HALO is a CONV mode");
  static_assert(NG == 2 || (NG == 4 && MT == 1 && !RING && !HALO && BN >= 160), "four epilogue groups: GEMM mode, table epilogues");
  static constexpr int kThreadsTotal = 64 + NG * 128;      // warp0 TMA, warp1 MMA, then NG epilogue groups of four warps
  static constexpr int kABytes = MT * BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = HALO ? kBBytes : kABytes + kBBytes;   // HALO: the ring holds weight tiles only
  static constexpr int kHaloTotal = HALO ? kHaloStages * kHaloBytes : 0;
  // epilogue staging: output columns leave in 32-column chunks (64-byte rows, 64B swizzle) through 2 smem
  // buffers per epilogue group
  static constexpr int CH = 32;
  static constexpr int kStgBytes = 128 * CH * 2;
  // RING: the residual tile is TMA-prefetched INTO the staging buffers (added in place, then stored), which needs
  // 4 buffers per group so that a residual chunk can be requested ~3 chunks before it is consumed
  // (NG = 4: ONE buffer per group -- a group's chunks are four chunks apart, so waiting for its previous store to have read
  // the buffer costs nothing, and the staging area stays at 32 KB)
  static constexpr int kBufPerGrp = RING ? 4 : (NG == 4 ? 1 : 2);
  static constexpr int kNumStg = NG * kBufPerGrp;
  // The epilogue-bound tile shapes (160-wide K=320/640 projections, 256-wide GEGLU) keep the tile's bias as an fp32
  // table in smem -- one copy per epilogue group, filled one tile ahead through registers -- instead of every
  // thread re-loading and unpacking the same bf16 values for every chunk.
  static constexpr bool kBiasTable = BN >= 160;
  static_assert(!LNF || kBiasTable, "the LayerNorm-folding epilogue reads c and u from the smem tables");
  static constexpr int kTableBytes = kBiasTable ? (LNF ? 2 : 1) * NG * BN * 4 : 0;   // [group][BN] bias (LNF: c), then [group][BN] u
  static constexpr int kStatBytes = 512;      // [2 epilogue groups][32 GroupNorm groups][sum, sumsq] fp32 partials (gn_stats)
  // dynamic smem is declared __align__(1024) (checked at run time), so no alignment slack is reserved
  static constexpr int kBudget = 227 * 1024 - kNumStg * kStgBytes - kTableBytes - kStatBytes - 256 - kHaloTotal;
  static constexpr int kMaxStages = HALO ? 6 : 8;
  static constexpr int kStages = kBudget / kStageBytes > kMaxStages ? kMaxStages : kBudget / kStageBytes;
  static constexpr int kSmemBytes = kHaloTotal + kStages * kStageBytes + kNumStg * kStgBytes + kTableBytes + kStatBytes + 256 /*barriers*/;
  static constexpr int kAccCols = MT * BN;        // TMEM columns of one accumulator stage
  static constexpr int kTmemCols = (2 * kAccCols <= 32) ? 32 : (2 * kAccCols <= 64) ? 64 : (2 * kAccCols <= 128) ? 128
                                   : (2 * kAccCols <= 256) ? 256 : 512;
  static_assert(2 * kAccCols <= 512, "two accumulator stages must fit TMEM");
};

// Epilogue flavours (compile-time: the epilogue is the critical path of the small-K layers, and one generic,
// runtime-branchy version thrashes the instruction cache)
enum : int { EPI_PLAIN = 0, EPI_GEGLU = 1, EPI_RESID = 2, EPI_ROWVEC = 3, EPI_ACT = 4, EPI_GENERAL = 5, EPI_RESID_DIRECT = 6 };
// residual tiles go through the TMA ring for the narrow-N tiles (the K=320/640 projections, where the per-thread
// strided residual reads were the bottleneck); BN=256 keeps 4 pipeline stages and reads the residual directly, and
// so do the cropped pano tiles (EPI_RESID_DIRECT, chosen by the host), whose rows are not a TMA box of the output.
// Ring or not is a compile-time property, so the ring kernels carry no (predicated-off) direct-read instructions.
template <int BN, int EPI> struct UseRing { static constexpr bool value = (EPI == EPI_RESID) && (BN <= 192); };

// LNF: LayerNorm folded into the GEMM.  The GEMM that produced the token matrix A wrote, from its epilogue, the partial
// (sum, sum of squares) of every row it stored (GemmConvParams::rowstats: one slot per N tile and epilogue group, no
// atomics, so the result is deterministic); the epilogue here adds the slots of its row, forms mean / rstd and applies
//   rstd * acc - mean * rstd * u[n] + c[n].
// The normalised activations are never written to or read from HBM: the LayerNorm pass (one read + one write of the
// token matrix per norm, 192 launches and ~13 ms of the 16x512x1024 step) disappears.  (A first version took the
// statistics from the A tiles in shared memory with four extra warps; they competed with the epilogue warps for issue
// slots and the step got 9 ms slower -- see git history / profiles/r02_bench_c3_ln_fold_on.json.)
// GroupNorm statistics of one 32-column chunk of my row: per group of GS channels (sum, sum of squares), reduced over the
// 32 rows of the warp by shuffles and added to the CTA's shared-memory partials by lane 0 (see GemmConvParams::gn_stats).
template <int GS>
__device__ __forceinline__ void gn_chunk_stats(const float2* f2, bool valid, int first_group, float* acc, int lane) {
  constexpr int G = 32 / GS, PP = GS / 2;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float2 a = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < PP; ++j) { a = fadd2(a, f2[g * PP + j]); q = ffma2(f2[g * PP + j], f2[g * PP + j], q); }
    float s = valid ? a.x + a.y : 0.f, qq = valid ? q.x + q.y : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); qq += __shfl_xor_sync(0xffffffffu, qq, o); }
    if (lane == 0) { atomicAdd(acc + (first_group + g) * 2, s); atomicAdd(acc + (first_group + g) * 2 + 1, qq); }
  }
}

// Per-channel statistics of one 32-column chunk: cs / cq hold my row's 32 values and their squares (zeros for rows outside the
// image).  Recursive halving over the warp: at offset o every lane keeps the half of its columns selected by its lane bit and
// receives the partner's partial sums of that half, so after five steps lane l holds column l summed over the warp's 32 rows
// (31 shuffles per array instead of 32 x 5).
__device__ __forceinline__ void warp_column_sums(float* cs, float* cq, int lane) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send_s = up ? cs[i] : cs[i + o], keep_s = up ? cs[i + o] : cs[i];
      const float send_q = up ? cq[i] : cq[i + o], keep_q = up ? cq[i + o] : cq[i];
      cs[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, o);
      cq[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, o);
    }
  }
}

template <int BN, int EPI, int MT = 1, bool LNF = false, bool HALO = false, int NG = 2>
__global__ void __launch_bounds__(64 + NG * 128, 1)
gemm_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR,
                 const GemmConvParams p) {
  constexpr bool kRing = UseRing<BN, EPI>::value;
  using C = Cfg<BN, kRing, MT, LNF, HALO, NG>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0) i360_device_fail("dynamic shared memory is not 1024-byte aligned (128B-swizzled TMA / UMMA tiles)");
  uint8_t* shalo = smem_raw;                          // HALO: kHaloStages activation tiles in front of the weight ring
  uint8_t* smem = smem_raw + C::kHaloTotal;
  const uint32_t base = smem_u32(smem);
  uint8_t* stg = smem + C::kStages * C::kStageBytes;            // 2 staging buffers (1024-aligned)
  float* sbias = reinterpret_cast<float*>(stg + C::kNumStg * C::kStgBytes);   // [2 groups][BN] (kBiasTable only)
  float* su = sbias + NG * BN;                                                 // [groups][BN] (LNF only)
  float* sgn = reinterpret_cast<float*>(stg + C::kNumStg * C::kStgBytes + C::kTableBytes);     // [2][32][2] (gn_stats)
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + C::kNumStg * C::kStgBytes + C::kTableBytes + C::kStatBytes);
  uint64_t* full = bars;                       // [kStages]
  uint64_t* empty = bars + C::kStages;         // [kStages]
  uint64_t* tfull = bars + 2 * C::kStages;     // [2]
  uint64_t* tempty = bars + 2 * C::kStages + 2;// [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);
  uint64_t* rbar = bars + 2 * C::kStages + 5;   // [2 groups][4]: residual chunk landed in staging buffer
  uint64_t* afull = rbar + 8;                   // [2] (HALO, never together with LNF): halo tile landed
  uint64_t* aempty = rbar + 10;                 // [2] (HALO): the nine taps that read the halo tile have retired

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmD);
    if (p.C2 > 0) tma_prefetch_desc(&tmA2);
    if (p.C3 > 0) tma_prefetch_desc(&tmA3);
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4 * NG); }
    if (kRing) { for (int s = 0; s < 8; ++s) mbar_init(&rbar[s], 1); tma_prefetch_desc(&tmR); }
    if (HALO) { for (int s = 0; s < kHaloStages; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); } }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, C::kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();        // barriers, TMEM and descriptors were set up under the previous kernel's tail (see common.cuh)
  griddep_launch();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int hstage = 0; uint32_t hphase = 0;      // HALO: activation ring
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int n_blk = t % p.n_tiles, m_blk = t / p.n_tiles;
        const int n0 = n_blk * BN;
        int w0 = 0, h0 = 0, b0 = 0;
        if (p.conv) {
          w0 = (m_blk % p.n_wt) * p.TW;
          h0 = ((m_blk / p.n_wt) % p.n_ht) * p.TH;
          b0 = (m_blk / (p.n_wt * p.n_ht)) * p.TB;
        }
        auto issue = [&](const CUtensorMap* ma, int c0, int dw, int dh, int kcoord) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          mbar_expect_tx(&full[stage], C::kStageBytes);
          if (p.conv) tma_load_4d(sa, ma, &full[stage], c0, p.in_stride * w0 + dw, p.in_stride * h0 + dh, b0);
          else        tma_load_2d(sa, ma, &full[stage], c0, m_blk * (BM * MT));
          tma_load_2d(sb, &tmW, &full[stage], kcoord, n0);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        };
        if (HALO) {
          // activation tile (halo box, or a plain 128-pixel box for the fused 1x1 sources) into the A ring, then the
          // weight tiles that multiply it into the B ring
          auto issue_a = [&](const CUtensorMap* ma, int c0, int cw, int ch, int bytes) {
            mbar_wait(&aempty[hstage], hphase ^ 1);
            mbar_expect_tx(&afull[hstage], bytes);
            tma_load_4d(shalo + hstage * kHaloBytes, ma, &afull[hstage], c0, cw, ch, b0);
            if (++hstage == kHaloStages) { hstage = 0; hphase ^= 1; }
          };
          auto issue_b = [&](int kcoord) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], C::kBBytes);
            tma_load_2d(smem + stage * C::kStageBytes, &tmW, &full[stage], kcoord, n0);
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          };
          for (int c0 = 0; c0 < p.Cin; c0 += BK) {
            issue_a(&tmA, c0, w0 - 1, h0 - 1, kHaloTxBytes);
            for (int tap = 0; tap < 9; ++tap) issue_b(tap * p.Cin + c0);
          }
          for (int c0 = 0; c0 < p.C2; c0 += BK) { issue_a(&tmA2, c0, w0 + p.xoff, h0, BM * BK * 2); issue_b(9 * p.Cin + c0); }
          for (int c0 = 0; c0 < p.C3; c0 += BK) { issue_a(&tmA3, c0, w0 + p.xoff, h0, BM * BK * 2); issue_b(9 * p.Cin + p.C2 + c0); }
        } else if (!p.conv) {
          for (int kb = 0; kb < p.k_iters; ++kb) issue(&tmA, kb * BK, 0, 0, kb * BK);
        } else {
          for (int tap = 0; tap < p.n_taps; ++tap) {
            const int dh = p.tap_dh[tap], dw = p.tap_dw[tap];
            for (int c0 = 0; c0 < p.Cin; c0 += BK) issue(&tmA, c0, dw, dh, tap * p.Cin + c0);
          }
          for (int c0 = 0; c0 < p.C2; c0 += BK) issue(&tmA2, c0, p.xoff, 0, 9 * p.Cin + c0);
          for (int c0 = 0; c0 < p.C3; c0 += BK) issue(&tmA3, c0, p.xoff, 0, 9 * p.Cin + p.C2 + c0);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int hstage = 0; uint32_t hphase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * C::kAccCols;
        if (HALO) {
          const int n_cin = (p.Cin + BK - 1) / BK, n_extra = (p.C2 + BK - 1) / BK + (p.C3 + BK - 1) / BK;
          uint32_t acc = 0;
          for (int cb = 0; cb < n_cin + n_extra; ++cb) {
            mbar_wait(&afull[hstage], hphase);
            const uint32_t sa = smem_u32(shalo) + hstage * kHaloBytes;
            const bool halo = cb < n_cin;          // else: a plain 128-pixel tile of a fused 1x1 source
            const int taps = halo ? 9 : 1;
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(&full[stage], phase);
              tc_fence_after();
              const uint32_t sb = base + stage * C::kStageBytes;
              const uint32_t sa_tap = halo ? sa + ((tap / 3) * kHaloW + tap % 3) * (BK * 2) : sa;
              const uint32_t sbo = halo ? kHaloW * BK * 2 : 1024;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                umma_bf16_ss(d_tmem, make_smem_desc(sa_tap + k * 32, sbo, 16, SWZ_128B),
                             make_smem_desc(sb + k * 32, 1024, 16, SWZ_128B), idesc, acc);
                acc = 1;
              }
              umma_commit(&empty[stage]);
              if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(&aempty[hstage]);          // every MMA that reads this activation tile has been issued before
            if (++hstage == kHaloStages) { hstage = 0; hphase ^= 1; }
          }
        } else
        for (int kb = 0; kb < p.k_iters; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = base + stage * C::kStageBytes;
          const uint32_t sb = sa + C::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc(sa + k * 32, 1024, 16, SWZ_128B);
            const uint64_t db = make_smem_desc(sb + k * 32, 1024, 16, SWZ_128B);
            umma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0);
            if (MT == 2)      // rows 128..255 of the tile: second accumulator, same weight tile
              umma_bf16_ss(d_tmem + BN, make_smem_desc(sa + BM * BK * 2 + k * 32, 1024, 16, SWZ_128B), db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);   // frees the smem slot once these MMAs retire
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[as]);        // accumulator complete -> epilogue
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (NG == 4) {
    // =============================== epilogue, four groups (warps 2..17) ===============================
    // The K = 320 / 640 GEGLU and QKV projections are bound by their epilogue: with two groups only two warps per SM
    // sub-partition work on ~35 instructions per output pair (erf-GELU, the LayerNorm fold) and their dependency
    // latency is exposed -- 6 200 clocks per 128 x 256 GEGLU tile against 2 560 of tensor-core time.  Four groups put four
    // epilogue warps on every sub-partition; to fit 576 threads (112 registers) a 32-column chunk is computed as two
    // 16-column halves.  Group g takes chunks g, g + 4, ... and owns ONE staging buffer (see Cfg).  Plain and GEGLU
    // epilogues with the smem bias table only (bias / LayerNorm-fold constants); nothing per-row but the LN statistics.
    static_assert(NG != 4 || EPI == EPI_PLAIN || EPI == EPI_GEGLU, "four-group epilogue: plain / GEGLU");
    constexpr int CH = C::CH;
    constexpr bool kGeglu = (EPI == EPI_GEGLU);
    const int ew = warp & 3;                      // TMEM lane quarter of this warp
    const int grp = (warp - 2) >> 2;
    const int row = ew * 32 + lane;
    const int gtid = ((warp - 2) & 3) * 32 + lane;
    const bool store_thread = (lane == 0) && (((warp - 2) & 3) == 0);
    const uint32_t swz = (row >> 1) & 3;
    uint8_t* buf = stg + grp * C::kStgBytes;
    uint8_t* my = buf + row * (CH * 2);
    const bool has_bias = p.bias != nullptr;
    float* tbias_w = sbias + grp * BN;
    float* tu_w = su + grp * BN;
    float nb0 = 0.f, nb1 = 0.f, nu0 = 0.f, nu1 = 0.f;
    auto load_tables = [&](int tt) {
      if (tt >= total_tiles) return;
      const int c0 = (tt % p.n_tiles) * BN + gtid;
      const bool in0 = c0 < p.N, in1 = (gtid + 128 < BN) && (c0 + 128 < p.N);
      if (LNF) {
        nb0 = in0 ? p.ln_c[c0] : 0.f; nu0 = in0 ? p.ln_u[c0] : 0.f;
        nb1 = in1 ? p.ln_c[c0 + 128] : 0.f; nu1 = in1 ? p.ln_u[c0 + 128] : 0.f;
        if (p.rowvec_in_table) {
          const float* rvp = p.rowvec + static_cast<long long>((((tt / p.n_tiles) * BM) / p.rowvec_div) % p.rowvec_mod) * p.rowvec_ld;
          if (in0) nb0 += rvp[c0];
          if (in1) nb1 += rvp[c0 + 128];
        }
      } else {
        nb0 = (has_bias && in0) ? __bfloat162float(p.bias[c0]) : 0.f;
        nb1 = (has_bias && in1) ? __bfloat162float(p.bias[c0 + 128]) : 0.f;
      }
    };
    load_tables(blockIdx.x);
    int as = 0; uint32_t aphase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int n_blk = t % p.n_tiles, m_blk = t / p.n_tiles;
      const long long r = static_cast<long long>(m_blk) * BM + row;
      // this tile's tables (fetched into registers one tile ago).  Everybody of the group is past the previous tile's
      // last chunk barrier, i.e. done reading the old table.
      tbias_w[gtid] = nb0;
      if (gtid + 128 < BN) tbias_w[gtid + 128] = nb1;
      if (LNF) { tu_w[gtid] = nu0; if (gtid + 128 < BN) tu_w[gtid + 128] = nu1; }
      named_bar_sync(1 + grp, 128);
      load_tables(t + gridDim.x);
      float2 rstd2 = make_float2(1.f, 1.f), nmr2 = make_float2(0.f, 0.f);
      if (LNF) {      // row statistics: the load latency hides under the wait for the accumulator (the kernel is MMA bound now)
        float s1 = 0.f, s2 = 0.f;
        if (r < p.M) {
          for (int sl = 0; sl < p.ln_slots; ++sl) {
            const float2 st = __ldg(p.ln_stats + static_cast<long long>(sl) * p.M + r);
            s1 += st.x; s2 += st.y;
          }
        }
        const float inv_k = 1.0f / static_cast<float>(p.K);
        const float mean = s1 * inv_k;
        const float rstd = rsqrtf(fmaxf(s2 * inv_k - mean * mean, 0.f) + p.ln_eps);
        rstd2 = make_float2(rstd, rstd); nmr2 = make_float2(-mean * rstd, -mean * rstd);
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * C::kAccCols;
      constexpr int n_chunks = kGeglu ? (BN / 2) / CH : BN / CH;
      const int oc0 = kGeglu ? n_blk * (BN / 2) : n_blk * BN;
      const float* tb = sbias + grp * BN;
      const float* tuu = su + grp * BN;
#pragma unroll 1
      for (int ci = grp; ci < n_chunks; ci += 4) {
        const int ocol = oc0 + ci * CH;
        const bool live = ocol < p.n_out;
        uint32_t pk[16];                                   // the chunk's 32 bf16 outputs of my row
        if (live) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int cc = ci * CH + h * 16;               // tile column of this half
            uint32_t v[16];
            tmem_ld_x16(t_acc + cc, v);
            if (kGeglu) {
              uint32_t vg[16];
              tmem_ld_x16(t_acc + BN / 2 + cc, vg);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 4; ++q) {                // 4 columns per step
                const float4 c4 = *reinterpret_cast<const float4*>(tb + cc + q * 4);
                const float4 g4 = *reinterpret_cast<const float4*>(tb + BN / 2 + cc + q * 4);
                float2 ba0 = make_float2(c4.x, c4.y), ba1 = make_float2(c4.z, c4.w);
                float2 bg0 = make_float2(g4.x, g4.y), bg1 = make_float2(g4.z, g4.w);
                if (LNF) {
                  const float4 u4 = *reinterpret_cast<const float4*>(tuu + cc + q * 4);
                  const float4 w4 = *reinterpret_cast<const float4*>(tuu + BN / 2 + cc + q * 4);
                  ba0 = ffma2(make_float2(u4.x, u4.y), nmr2, ba0); ba1 = ffma2(make_float2(u4.z, u4.w), nmr2, ba1);
                  bg0 = ffma2(make_float2(w4.x, w4.y), nmr2, bg0); bg1 = ffma2(make_float2(w4.z, w4.w), nmr2, bg1);
                }
                const float2 a0 = make_float2(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]));
                const float2 a1 = make_float2(__uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
                const float2 g0 = make_float2(__uint_as_float(vg[q * 4]), __uint_as_float(vg[q * 4 + 1]));
                const float2 g1 = make_float2(__uint_as_float(vg[q * 4 + 2]), __uint_as_float(vg[q * 4 + 3]));
                const float2 val0 = LNF ? ffma2(a0, rstd2, ba0) : fadd2(a0, ba0), val1 = LNF ? ffma2(a1, rstd2, ba1) : fadd2(a1, ba1);
                const float2 gat0 = LNF ? ffma2(g0, rstd2, bg0) : fadd2(g0, bg0), gat1 = LNF ? ffma2(g1, rstd2, bg1) : fadd2(g1, bg1);
                const float2 o0 = fmul2(val0, gelu_gate2(gat0)), o1 = fmul2(val1, gelu_gate2(gat1));
                pk[h * 8 + q * 2] = pack_bf16x2(o0.x, o0.y);
                pk[h * 8 + q * 2 + 1] = pack_bf16x2(o1.x, o1.y);
              }
            } else {
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float2 a0 = make_float2(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]));
                float2 a1 = make_float2(__uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
                if (LNF || has_bias) {
                  const float4 c4 = *reinterpret_cast<const float4*>(tb + cc + q * 4);
                  float2 b0 = make_float2(c4.x, c4.y), b1 = make_float2(c4.z, c4.w);
                  if (LNF) {
                    const float4 u4 = *reinterpret_cast<const float4*>(tuu + cc + q * 4);
                    b0 = ffma2(make_float2(u4.x, u4.y), nmr2, b0); b1 = ffma2(make_float2(u4.z, u4.w), nmr2, b1);
                    a0 = ffma2(a0, rstd2, b0); a1 = ffma2(a1, rstd2, b1);
                  } else {
                    a0 = fadd2(a0, b0); a1 = fadd2(a1, b1);
                  }
                }
                pk[h * 8 + q * 2] = pack_bf16x2(a0.x, a0.y);
                pk[h * 8 + q * 2 + 1] = pack_bf16x2(a1.x, a1.y);
              }
            }
          }
        }
        // the group's previous store must have finished READING the buffer before it is overwritten
        if (store_thread) bulk_wait_read<0>();
        named_bar_sync(1 + grp, 128);
        if (live) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(my + ((static_cast<uint32_t>(g) ^ swz) << 4)) = make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + grp, 128);
        if (store_thread && live) {
          tma_store_2d(&tmD, buf, ocol, m_blk * BM);
          bulk_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (store_thread) bulk_wait<0>();
  } else {
    // =============================== epilogue (warps 2..9) ===============================
    // TMEM -> registers -> (bias / temb / activation / residual) -> bf16 -> swizzled smem -> TMA tensor store.
    // Two groups of 4 warps take alternate 32-column chunks; each group double-buffers its staging tile and
    // needs ONE named barrier per chunk: the group's store thread waits (before that barrier) until the store
    // issued one chunk earlier has finished reading, so after the barrier everybody may overwrite that buffer.
    // The TMA store clips rows/columns outside the output tensor (M/N tails).  Cropped pano tiles (crop > 0)
    // would need a negative start coordinate, which TMA stores reject: those use predicated global stores.
    constexpr int CH = C::CH;
    constexpr bool kGeglu = (EPI == EPI_GEGLU);
    constexpr bool kResid = (EPI == EPI_RESID) || (EPI == EPI_RESID_DIRECT) || (EPI == EPI_GENERAL);
    constexpr bool kRowvec = (EPI == EPI_ROWVEC) || (EPI == EPI_GENERAL);
    constexpr bool kAct = (EPI == EPI_ACT) || (EPI == EPI_GENERAL);
    const int ew = warp & 3;                 // TMEM lane quarter this warp may read
    const int grp = (warp - 2) >> 2;         // epilogue group 0/1: chunks grp, grp+2, ...
    const int row = ew * 32 + lane;          // row inside the 128-row tile
    const bool store_thread = (lane == 0) && (warp == 2 || warp == 6);
    const bool direct = !kRing && p.conv && p.crop > 0;
    const uint32_t swz = (row >> 1) & 3;     // 64B swizzle: 16-byte chunk index ^= address bits [7,9)
    constexpr int NB = C::kBufPerGrp;
    uint8_t* gbuf = stg + grp * NB * C::kStgBytes;
    uint32_t kchunk = 0;                     // chunks consumed by this group so far (buffer = kchunk % NB)
    // residual prefetch iterator of the group's store thread: runs NB-1 chunks ahead of consumption
    constexpr int n_chunks_c = (EPI == EPI_GEGLU) ? (BN / 2) / C::CH : BN / C::CH;
    constexpr int kCi0 = 0, kCiStep = (MT == 2) ? 1 : 2;   // MT = 2: the group walks every chunk of ITS accumulator
    const int ci_first = (MT == 2) ? kCi0 : grp;
    const int row_off = (MT == 2) ? grp * BM : 0;          // tile row of this group's accumulator row 0
    int pf_t = blockIdx.x, pf_ci = ci_first; uint32_t pf_k = 0;
    auto prefetch_resid = [&]() {            // request the residual of the next not-yet-requested chunk
      while (pf_t < total_tiles && pf_ci >= n_chunks_c) { pf_t += gridDim.x; pf_ci = ci_first; }
      if (pf_t >= total_tiles) return;
      const int pn = pf_t % p.n_tiles, pm = pf_t / p.n_tiles;
      const int col = pn * BN + pf_ci * C::CH;
      uint8_t* dstb = gbuf + (pf_k % NB) * C::kStgBytes;
      uint64_t* bar = &rbar[grp * 4 + (pf_k % NB)];
      if (col < p.n_out) {
        mbar_expect_tx(bar, C::kStgBytes);
        if (p.conv) {
          const int pw0 = (pm % p.n_wt) * p.TW, ph0 = ((pm / p.n_wt) % p.n_ht) * p.TH, pb0 = (pm / (p.n_wt * p.n_ht)) * p.TB;
          tma_load_4d(dstb, &tmR, bar, col, pw0, ph0, pb0);
        } else {
          tma_load_2d(dstb, &tmR, bar, col, pm * (BM * MT) + row_off);
        }
      } else {
        mbar_arrive(bar);                    // dead chunk: keep the phase sequence in step
      }
      pf_ci += kCiStep; ++pf_k;
    };
    constexpr bool ring_on = kRing;          // host dispatch guarantees: residual present, tile rows are a TMA box
    if (ring_on && store_thread) { for (int i = 0; i < NB - 1; ++i) prefetch_resid(); }
    const bool has_bias = p.bias != nullptr;
    const int gtid = (warp & 3) * 32 + lane;             // thread index inside the epilogue group
    float nb0 = 0.f, nb1 = 0.f;                          // bias of columns gtid, gtid + 128 of the NEXT tile
    float nu0 = 0.f, nu1 = 0.f;                          // LNF: u of the same columns
    auto load_bias = [&](int tt) {
      if (tt < total_tiles) {
        const int c0 = (tt % p.n_tiles) * BN + gtid;
        if (LNF) {
          nb0 = (c0 < p.N) ? p.ln_c[c0] : 0.f;
          nu0 = (c0 < p.N) ? p.ln_u[c0] : 0.f;
          nb1 = (gtid + 128 < BN && c0 + 128 < p.N) ? p.ln_c[c0 + 128] : 0.f;
          nu1 = (gtid + 128 < BN && c0 + 128 < p.N) ? p.ln_u[c0 + 128] : 0.f;
          if (p.rowvec_in_table) {
            const float* rvp = p.rowvec + static_cast<long long>((((tt / p.n_tiles) * (BM * MT)) / p.rowvec_div) % p.rowvec_mod) * p.rowvec_ld;
            if (c0 < p.N) nb0 += rvp[c0];
            if (gtid + 128 < BN && c0 + 128 < p.N) nb1 += rvp[c0 + 128];
          }
        } else {
          nb0 = (has_bias && c0 < p.N) ? __bfloat162float(p.bias[c0]) : 0.f;
          nb1 = (has_bias && gtid + 128 < BN && c0 + 128 < p.N) ? __bfloat162float(p.bias[c0 + 128]) : 0.f;
        }
      }
    };
    if (C::kBiasTable && (has_bias || kGeglu || LNF)) load_bias(blockIdx.x);
    float2 pst[kMaxStatSlots];                           // LNF: row-statistics slots of my row of the NEXT tile
    auto load_stats = [&](int tt) {
      if (!LNF) return;
      const long long r = (tt < total_tiles) ? static_cast<long long>(tt / p.n_tiles) * (BM * MT) + row_off + row : p.M;
#pragma unroll
      for (int sl = 0; sl < kMaxStatSlots; ++sl)
        pst[sl] = (r < p.M && sl < p.ln_slots) ? __ldg(p.ln_stats + static_cast<long long>(sl) * p.M + r) : make_float2(0.f, 0.f);
    };
    load_stats(blockIdx.x);
    // GroupNorm statistics of the output (gn_stats): this epilogue group's fp32 partials of the image it is working on live in
    // smem and are flushed (fp64 atomics, 64 per group) when the CTA's tile sequence moves on to the next image and at the end
    const bool gn_on = !kGeglu && !LNF && p.conv && p.gn_stats != nullptr;
    float* gacc = sgn + grp * 64;
    int gn_img = -1;
    auto gn_flush = [&](int img) {
      named_bar_sync(1 + grp, 128);                      // every warp's smem atomics of the image have landed
      if (gtid < 64) {
        const float v = gacc[gtid];
        if (v != 0.f) atomicAdd(p.gn_stats + static_cast<long long>(img) * 64 + gtid, static_cast<double>(v));
        gacc[gtid] = 0.f;
      }
      named_bar_sync(1 + grp, 128);
    };
    if (gn_on) { if (gtid < 64) gacc[gtid] = 0.f; named_bar_sync(1 + grp, 128); }
    int as = 0; uint32_t aphase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int n_blk = t % p.n_tiles, m_blk = t / p.n_tiles;
      // ---- row mapping (residual / per-image vector reads, direct stores) ----
      bool valid = true; long long orow = 0; int vec_idx = 0;
      long long dstr = -1;                     // strided output lattice: element offset of my row (direct stores only)
      int w0 = 0, h0 = 0, b0 = 0;
      if (p.conv) {
        w0 = (m_blk % p.n_wt) * p.TW; h0 = ((m_blk / p.n_wt) % p.n_ht) * p.TH; b0 = (m_blk / (p.n_wt * p.n_ht)) * p.TB;
        if (gn_on && b0 != gn_img) { if (gn_img >= 0) gn_flush(gn_img); gn_img = b0; }
        if (kResid || kRowvec || direct || gn_on || p.gn_chan != nullptr) {
          const int tw = row % p.TW, th = (row / p.TW) % p.TH, tb = row / (p.TW * p.TH);
          const int w = w0 + tw, h = h0 + th, b = b0 + tb;
          valid = (b < p.B) && (h < p.H) && (w >= p.crop) && (w < p.W - p.crop);
          orow = (static_cast<long long>(b) * p.Hout + h) * p.Wout + (w - p.crop);
          if (p.os_w != 0) dstr = static_cast<long long>(b) * p.os_b + static_cast<long long>(h) * p.os_h + static_cast<long long>(w - p.crop) * p.os_w;
          vec_idx = b;
        }
      } else {
        const int r = m_blk * (BM * MT) + row_off + row;
        valid = r < p.M; orow = r; vec_idx = r;
      }
      if (!valid) { orow = 0; vec_idx = 0; }
      const int rv_row = p.rowvec_mod > 0 ? (vec_idx / p.rowvec_div) % p.rowvec_mod : vec_idx / p.rowvec_div;
      const float* rv = (kRowvec && p.rowvec != nullptr) ? p.rowvec + static_cast<long long>(rv_row) * p.rowvec_ld : nullptr;
      const bf16* rrow = (kResid && !ring_on && p.resid != nullptr && valid) ? p.resid + orow * p.ldr : nullptr;
      bf16* drow = (dstr >= 0 && valid) ? p.D + dstr : p.D + orow * p.ldd;

      const float* tbias = sbias + grp * BN;
      const float* tu = su + grp * BN;
      if (C::kBiasTable && (has_bias || kGeglu || LNF)) {
        // this tile's values were fetched into registers during the previous tile.  Every thread of the group has
        // finished reading the previous tile's table once it passed that tile's last chunk barrier; the cropped
        // direct-store mode has no chunk barriers, hence the extra one.
        if (direct) named_bar_sync(1 + grp, 128);
        float* tw = sbias + grp * BN;
        tw[gtid] = nb0;
        if (gtid + 128 < BN) tw[gtid + 128] = nb1;
        if (LNF) {
          float* uw = su + grp * BN;
          uw[gtid] = nu0;
          if (gtid + 128 < BN) uw[gtid + 128] = nu1;
        }
        named_bar_sync(1 + grp, 128);
        load_bias(t + gridDim.x);                      // lands under this tile's chunks
      }
      float2 rstd2 = make_float2(1.f, 1.f), nmr2 = make_float2(0.f, 0.f);     // LNF: rstd and -mean * rstd of my row
      if (LNF) {
        // the partial sums of this row were requested one tile ago (a global-load latency per tile on the epilogue's
        // critical path cost more than the LayerNorm pass saved); the next tile's are requested right away
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int sl = 0; sl < kMaxStatSlots; ++sl) { s1 += pst[sl].x; s2 += pst[sl].y; }
        load_stats(t + gridDim.x);
        const float inv_k = 1.0f / static_cast<float>(p.K);
        const float mean = s1 * inv_k;
        const float rstd = rsqrtf(fmaxf(s2 * inv_k - mean * mean, 0.f) + p.ln_eps);
        rstd2 = make_float2(rstd, rstd); nmr2 = make_float2(-mean * rstd, -mean * rstd);
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * C::kAccCols + (MT == 2 ? grp * BN : 0);
      constexpr int n_chunks = kGeglu ? (BN / 2) / CH : BN / CH;
      const int oc0 = kGeglu ? n_blk * (BN / 2) : n_blk * BN;     // first OUTPUT column of this tile
      float2 rs1 = make_float2(0.f, 0.f), rs2 = make_float2(0.f, 0.f);   // row statistics of the values this thread stores

#pragma unroll 1
      for (int ci = ci_first; ci < n_chunks; ci += kCiStep) {
        const int ocol = oc0 + ci * CH;                 // output column of staged column 0
        const bool live = ocol < p.n_out;
        uint8_t* buf = gbuf + (kchunk % NB) * C::kStgBytes;
        uint8_t* my = buf + row * (CH * 2);
        if (live) {
          float2 f2[16];                     // 32 consecutive output columns as 16 packed pairs
          uint32_t v[32];
          tmem_ld_x16(t_acc + ci * CH, v);
          tmem_ld_x16(t_acc + ci * CH + 16, v + 16);
          if (kGeglu) {
            uint32_t vg[32];
            tmem_ld_x16(t_acc + BN / 2 + ci * CH, vg);
            tmem_ld_x16(t_acc + BN / 2 + ci * CH + 16, vg + 16);
            tmem_ld_wait();
            const int wc = n_blk * BN + ci * CH;          // packed weight/bias row of value column 0
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float2 ba[4], bg[4];
              if (C::kBiasTable) {
                // fp32 table (zeros when the layer has no bias): broadcast LDS.128, no unpacking, no branches
                const float4 x0 = *reinterpret_cast<const float4*>(tbias + ci * CH + g * 8);
                const float4 x1 = *reinterpret_cast<const float4*>(tbias + ci * CH + g * 8 + 4);
                const float4 y0 = *reinterpret_cast<const float4*>(tbias + BN / 2 + ci * CH + g * 8);
                const float4 y1 = *reinterpret_cast<const float4*>(tbias + BN / 2 + ci * CH + g * 8 + 4);
                ba[0] = make_float2(x0.x, x0.y); ba[1] = make_float2(x0.z, x0.w); ba[2] = make_float2(x1.x, x1.y); ba[3] = make_float2(x1.z, x1.w);
                bg[0] = make_float2(y0.x, y0.y); bg[1] = make_float2(y0.z, y0.w); bg[2] = make_float2(y1.x, y1.y); bg[3] = make_float2(y1.z, y1.w);
              } else if (has_bias) {
                const uint4 x = *reinterpret_cast<const uint4*>(p.bias + wc + g * 8);
                const uint4 y = *reinterpret_cast<const uint4*>(p.bias + wc + BN / 2 + g * 8);
                ba[0] = unpack_bf16x2(x.x); ba[1] = unpack_bf16x2(x.y); ba[2] = unpack_bf16x2(x.z); ba[3] = unpack_bf16x2(x.w);
                bg[0] = unpack_bf16x2(y.x); bg[1] = unpack_bf16x2(y.y); bg[2] = unpack_bf16x2(y.z); bg[3] = unpack_bf16x2(y.w);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { ba[j] = make_float2(0.f, 0.f); bg[j] = make_float2(0.f, 0.f); }
              }
              if (LNF) {       // rstd * acc - mean * rstd * u + c for the value and the gate column
                const float4 p0 = *reinterpret_cast<const float4*>(tu + ci * CH + g * 8);
                const float4 p1 = *reinterpret_cast<const float4*>(tu + ci * CH + g * 8 + 4);
                const float4 q0 = *reinterpret_cast<const float4*>(tu + BN / 2 + ci * CH + g * 8);
                const float4 q1 = *reinterpret_cast<const float4*>(tu + BN / 2 + ci * CH + g * 8 + 4);
                ba[0] = ffma2(make_float2(p0.x, p0.y), nmr2, ba[0]); ba[1] = ffma2(make_float2(p0.z, p0.w), nmr2, ba[1]);
                ba[2] = ffma2(make_float2(p1.x, p1.y), nmr2, ba[2]); ba[3] = ffma2(make_float2(p1.z, p1.w), nmr2, ba[3]);
                bg[0] = ffma2(make_float2(q0.x, q0.y), nmr2, bg[0]); bg[1] = ffma2(make_float2(q0.z, q0.w), nmr2, bg[1]);
                bg[2] = ffma2(make_float2(q1.x, q1.y), nmr2, bg[2]); bg[3] = ffma2(make_float2(q1.z, q1.w), nmr2, bg[3]);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 av = make_float2(__uint_as_float(v[g * 8 + 2 * j]), __uint_as_float(v[g * 8 + 2 * j + 1]));
                const float2 ag = make_float2(__uint_as_float(vg[g * 8 + 2 * j]), __uint_as_float(vg[g * 8 + 2 * j + 1]));
                const float2 val = LNF ? ffma2(av, rstd2, ba[j]) : fadd2(av, ba[j]);
                const float2 gate = LNF ? ffma2(ag, rstd2, bg[j]) : fadd2(ag, bg[j]);
                f2[g * 4 + j] = fmul2(val, gelu_gate2(gate));
              }
            }
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) f2[j] = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            if (LNF) {                       // rstd * acc + (-mean * rstd) * u + c
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 b4 = *reinterpret_cast<const float4*>(tbias + ci * CH + q * 4);
                const float4 u4 = *reinterpret_cast<const float4*>(tu + ci * CH + q * 4);
                f2[2 * q] = ffma2(f2[2 * q], rstd2, ffma2(make_float2(u4.x, u4.y), nmr2, make_float2(b4.x, b4.y)));
                f2[2 * q + 1] = ffma2(f2[2 * q + 1], rstd2, ffma2(make_float2(u4.z, u4.w), nmr2, make_float2(b4.z, b4.w)));
              }
            } else if (has_bias) {
              if (C::kBiasTable) {           // fp32 table in smem (zero past N): broadcast LDS.128, packed adds
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 b4 = *reinterpret_cast<const float4*>(tbias + ci * CH + q * 4);
                  f2[2 * q] = fadd2(f2[2 * q], make_float2(b4.x, b4.y));
                  f2[2 * q + 1] = fadd2(f2[2 * q + 1], make_float2(b4.z, b4.w));
                }
              } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  if (ocol + g * 8 < p.N) {
                    const uint4 bb = *reinterpret_cast<const uint4*>(p.bias + ocol + g * 8);
                    f2[g * 4 + 0] = fadd2(f2[g * 4 + 0], unpack_bf16x2(bb.x)); f2[g * 4 + 1] = fadd2(f2[g * 4 + 1], unpack_bf16x2(bb.y));
                    f2[g * 4 + 2] = fadd2(f2[g * 4 + 2], unpack_bf16x2(bb.z)); f2[g * 4 + 3] = fadd2(f2[g * 4 + 3], unpack_bf16x2(bb.w));
                  }
                }
              }
            }
            if (kRowvec && rv) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int col = ocol + g * 8;
                if (col < p.N) {
                  const float4 r0 = *reinterpret_cast<const float4*>(rv + col);
                  const float4 r1 = *reinterpret_cast<const float4*>(rv + col + 4);
                  f2[g * 4 + 0] = fadd2(f2[g * 4 + 0], make_float2(r0.x, r0.y)); f2[g * 4 + 1] = fadd2(f2[g * 4 + 1], make_float2(r0.z, r0.w));
                  f2[g * 4 + 2] = fadd2(f2[g * 4 + 2], make_float2(r1.x, r1.y)); f2[g * 4 + 3] = fadd2(f2[g * 4 + 3], make_float2(r1.z, r1.w));
                }
              }
            }
            if (kAct) {
              if (p.act == 2) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f2[j] = gelu_erf2(f2[j]);
              } else if (p.act == 3) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f2[j] = make_float2(silu(f2[j].x), silu(f2[j].y));
              }
            }
            if (kResid && !kRing && rrow) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int col = ocol + g * 8;
                if (col < p.N) {
                  const uint4 rr = *reinterpret_cast<const uint4*>(rrow + col);
                  f2[g * 4 + 0] = fadd2(f2[g * 4 + 0], unpack_bf16x2(rr.x)); f2[g * 4 + 1] = fadd2(f2[g * 4 + 1], unpack_bf16x2(rr.y));
                  f2[g * 4 + 2] = fadd2(f2[g * 4 + 2], unpack_bf16x2(rr.z)); f2[g * 4 + 3] = fadd2(f2[g * 4 + 3], unpack_bf16x2(rr.w));
                }
              }
            }
          }
          if (ring_on) {                     // residual chunk was TMA-prefetched into my staging row
            mbar_wait(&rbar[grp * 4 + (kchunk % NB)], (kchunk / NB) & 1);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 rr = *reinterpret_cast<const uint4*>(my + ((static_cast<uint32_t>(g) ^ swz) << 4));
              f2[g * 4 + 0] = fadd2(f2[g * 4 + 0], unpack_bf16x2(rr.x)); f2[g * 4 + 1] = fadd2(f2[g * 4 + 1], unpack_bf16x2(rr.y));
              f2[g * 4 + 2] = fadd2(f2[g * 4 + 2], unpack_bf16x2(rr.z)); f2[g * 4 + 3] = fadd2(f2[g * 4 + 3], unpack_bf16x2(rr.w));
            }
          }
          if (p.out_scale != 1.0f) {
            const float2 sc = make_float2(p.out_scale, p.out_scale);
#pragma unroll
            for (int j = 0; j < 16; ++j) f2[j] = fmul2(f2[j], sc);
          }
          if (!kGeglu && !LNF && p.conv && p.gn_chan != nullptr) {     // per-channel GroupNorm statistics of the STORED values
            float cs[32], cq[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 r = unpack_bf16x2(pack_bf16x2(f2[j].x, f2[j].y));
              cs[2 * j] = valid ? r.x : 0.f; cs[2 * j + 1] = valid ? r.y : 0.f;
              cq[2 * j] = cs[2 * j] * cs[2 * j]; cq[2 * j + 1] = cs[2 * j + 1] * cs[2 * j + 1];
            }
            warp_column_sums(cs, cq, lane);
            const int col = ocol + lane;
            if (col < p.n_out) {
              // the warp's 32 rows belong to ONE image (host guarantees TW * TH >= 32 pixels of an image per tile)
              const int img = b0 + (ew * 32) / (p.TW * p.TH);
              if (img < p.B) {
                double* dst = p.gn_chan + (static_cast<long long>(img) * p.n_out + col) * 2;
                atomicAdd(dst, static_cast<double>(cs[0]));       // fp64: the order of the adds does not reach the bf16 result
                atomicAdd(dst + 1, static_cast<double>(cq[0]));
              }
            }
          }
          if (gn_on) {                                          // GroupNorm statistics of what this chunk stores
            const int fg = ocol / p.gn_gs;
            if (p.gn_gs == 4) gn_chunk_stats<4>(f2, valid, fg, gacc, lane);
            else if (p.gn_gs == 8) gn_chunk_stats<8>(f2, valid, fg, gacc, lane);
            else gn_chunk_stats<16>(f2, valid, fg, gacc, lane);
          }
          if (!kGeglu && !LNF && p.rowstats != nullptr) {     // producer side of the LayerNorm fold (columns past N are zeros)
#pragma unroll
            for (int j = 0; j < 16; ++j) { rs1 = fadd2(rs1, f2[j]); rs2 = ffma2(f2[j], f2[j], rs2); }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 packed = make_uint4(pack_bf16x2(f2[g * 4 + 0].x, f2[g * 4 + 0].y), pack_bf16x2(f2[g * 4 + 1].x, f2[g * 4 + 1].y),
                                            pack_bf16x2(f2[g * 4 + 2].x, f2[g * 4 + 2].y), pack_bf16x2(f2[g * 4 + 3].x, f2[g * 4 + 3].y));
            if (direct) {
              const int col = ocol + g * 8;
              if (valid && col < p.n_out) *reinterpret_cast<uint4*>(drow + col) = packed;
            } else {
              *reinterpret_cast<uint4*>(my + ((static_cast<uint32_t>(g) ^ swz) << 4)) = packed;
            }
          }
        }
        if (!direct) {
          if (ring_on && !live) mbar_wait(&rbar[grp * 4 + (kchunk % NB)], (kchunk / NB) & 1);   // consume the dead chunk's phase
          fence_proxy_async_smem();
          if (store_thread) bulk_wait_read<0>();     // every earlier store of this group is done reading its buffer
          named_bar_sync(1 + grp, 128);
          if (store_thread) {
            if (live) {
              if (p.conv) tma_store_4d(&tmD, buf, ocol, w0, h0, b0);
              else        tma_store_2d(&tmD, buf, ocol, m_blk * (BM * MT) + row_off);
              bulk_commit();
            }
            // buffer (kchunk+NB-1) % NB == (kchunk-1) % NB was read by the store of chunk kchunk-1, which the
            // wait above retired -> it can receive the residual of chunk kchunk+NB-1
            if (ring_on) prefetch_resid();
          }
          ++kchunk;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
      if (!kGeglu && !LNF && p.rowstats != nullptr && !p.conv && valid) {
        const int slot = (MT == 2) ? n_blk : n_blk * 2 + grp;
        p.rowstats[static_cast<long long>(slot) * p.M + orow] = make_float2(rs1.x + rs1.y, rs2.x + rs2.y);
      }
    }
    if (gn_on && gn_img >= 0) gn_flush(gn_img);
    if (store_thread) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, C::kTmemCols); }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <int BN, int EPI, int MT = 1, bool LNF = false, bool HALO = false, int NG = 2>
static int launch(const CUtensorMap& a, const CUtensorMap& a2, const CUtensorMap& a3,
                  const CUtensorMap& w, const CUtensorMap& d, const CUtensorMap& rmap, const GemmConvParams& p,
                  cudaStream_t st) {
  using C = Cfg<BN, UseRing<BN, EPI>::value, MT, LNF, HALO, NG>;
  static_assert(C::kStages >= 3, "weight / operand ring too shallow");
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_conv_kernel<BN, EPI, MT, LNF, HALO, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             C::kSmemBytes) != cudaSuccess)
      return I360_ERR_CUDA;
    attr_set = true;
  }
  int grid = p.m_tiles * p.n_tiles;
  if (grid > num_sms()) grid = num_sms();
  if (grid <= 0) return I360_OK;
  launch_k(gemm_conv_kernel<BN, EPI, MT, LNF, HALO, NG>, dim3(grid), dim3(C::kThreadsTotal), C::kSmemBytes, st, a, a2, a3, w, d, rmap, p);
  I360_CUDA_CHECK_LAUNCH();
  return I360_OK;
}

// experiment hook: I360_BN_MAP="960:192,640:128" overrides the tile width chosen for a given N (non-GEGLU)
static int bn_override(int N) {
  static int table[16][2]; static int n_entries = -1;
  if (n_entries < 0) {
    n_entries = 0;
    const char* e = getenv("I360_BN_MAP");
    while (e && *e && n_entries < 16) {
      int n = 0, b = 0;
      if (sscanf(e, "%d:%d", &n, &b) == 2) { table[n_entries][0] = n; table[n_entries][1] = b; ++n_entries; }
      e = strchr(e, ',');
      if (e) ++e;
    }
  }
  for (int i = 0; i < n_entries; ++i) if (table[i][0] == N) return table[i][1];
  return 0;
}

static int pick_bn(int N, int act) {
  // GEGLU tiles hold [values | gates] halves that must be whole 64-column store chunks
  if (act == 1) return (N % 256 == 0) ? 256 : 128;
  if (const int o = bn_override(N)) return o;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  // minimise padded columns; prefer the wider tile on ties (better smem traffic per FLOP)
  int best = 256; long best_waste = ((N + 255) / 256) * 256L - N;
  // 192 before 160: 6 whole 32-column epilogue chunks split 3:3 over the two epilogue groups (160 gives 3:2), and
  // N = 960 / 1920 (the fused QKV projections) need fewer tiles: measured -10 % / -4..8 % on those GEMMs
  const int cands[4] = {192, 160, 128, 64};
  for (int i = 0; i < 4; ++i) {
    long w = ((N + cands[i] - 1) / cands[i]) * (long)cands[i] - N;
    if (w < best_waste) { best_waste = w; best = cands[i]; }
  }
  return best;
}

static int pick_epi(const GemmConvParams& p) {
  if (p.act == 1) return EPI_GEGLU;
  const bool r = p.resid != nullptr, v = p.rowvec != nullptr, a = p.act != 0;
  if (!r && !v && !a) return EPI_PLAIN;
  if (r && !v && !a) return (p.conv && p.crop > 0) ? EPI_RESID_DIRECT : EPI_RESID;
  if (!r && v && !a) return EPI_ROWVEC;
  if (!r && !v && a) return EPI_ACT;
  return EPI_GENERAL;
}

// Epilogue groups of the LayerNorm-folded / GEGLU projections.  Measured on one B200 (tools/ln_fold_probe.py): with four
// groups (576 threads) the folded QKV projection 655 360 x 960 x 320 goes 0.503 -> 0.451 ms and the plain GEGLU x 2560 x 320
// 1.10 -> 0.99 ms, but the folded GEGLU 1.22 -> 1.43 ms (its per-tile preamble -- tables, row statistics, three barriers -- is
// paid by every group and is as long as the one chunk a group then computes) and N = 320 / 1920 are unchanged; with four
// groups everywhere the step was 312.6 / 313.8 ms against 312.7 / 312.8 ms with two, with four groups for the folded PLAIN
// projections only 305.6 ms against 304.7 / 305.8 ms: no gain inside the step -> the two-group kernels (which prefetch the row
// statistics one tile ahead) stay the default.  I360_EPI_GROUPS=4 selects four groups everywhere, =3 for the plain ones only.
static int epi_groups_env() {
  static const int v = getenv("I360_EPI_GROUPS") ? atoi(getenv("I360_EPI_GROUPS")) : 0;
  return v;
}
static bool four_groups() { return epi_groups_env() == 4; }                 // GEGLU kernels
static bool four_groups_plain() { return epi_groups_env() >= 3; }           // LayerNorm-folded plain projections

template <int BN>
static int dispatch_epi(const CUtensorMap& a, const CUtensorMap& a2, const CUtensorMap& a3, const CUtensorMap& w,
                        const CUtensorMap& d, const CUtensorMap& r, const GemmConvParams& p, cudaStream_t st) {
  if (p.halo) {
    switch (pick_epi(p)) {
      case EPI_PLAIN: return launch<BN, EPI_PLAIN, 1, false, true>(a, a2, a3, w, d, r, p, st);
      case EPI_RESID: return launch<BN, EPI_RESID, 1, false, true>(a, a2, a3, w, d, r, p, st);
      case EPI_ROWVEC: return launch<BN, EPI_ROWVEC, 1, false, true>(a, a2, a3, w, d, r, p, st);
      case EPI_RESID_DIRECT: return launch<BN, EPI_RESID_DIRECT, 1, false, true>(a, a2, a3, w, d, r, p, st);
      default: return I360_ERR_UNSUPPORTED;      // the host only selects HALO for these four epilogues
    }
  }
  switch (pick_epi(p)) {
    case EPI_PLAIN: return launch<BN, EPI_PLAIN>(a, a2, a3, w, d, r, p, st);
    case EPI_GEGLU:
      if constexpr (BN == 256) { if (four_groups()) return launch<BN, EPI_GEGLU, 1, false, false, 4>(a, a2, a3, w, d, r, p, st); }
      return launch<BN, EPI_GEGLU>(a, a2, a3, w, d, r, p, st);
    case EPI_RESID: return launch<BN, EPI_RESID>(a, a2, a3, w, d, r, p, st);
    case EPI_ROWVEC: return launch<BN, EPI_ROWVEC>(a, a2, a3, w, d, r, p, st);
    case EPI_ACT: return launch<BN, EPI_ACT>(a, a2, a3, w, d, r, p, st);
    case EPI_RESID_DIRECT: return launch<BN, EPI_RESID_DIRECT>(a, a2, a3, w, d, r, p, st);
    default: return launch<BN, EPI_GENERAL>(a, a2, a3, w, d, r, p, st);
  }
}

static int dispatch(int bn, const CUtensorMap& a, const CUtensorMap& a2, const CUtensorMap& a3,
                    const CUtensorMap& w, const CUtensorMap& d, const CUtensorMap& r, const GemmConvParams& p,
                    cudaStream_t st) {
  switch (bn) {
    case 64: return dispatch_epi<64>(a, a2, a3, w, d, r, p, st);
    case 128: return dispatch_epi<128>(a, a2, a3, w, d, r, p, st);
    case 160: return dispatch_epi<160>(a, a2, a3, w, d, r, p, st);
    case 192: return dispatch_epi<192>(a, a2, a3, w, d, r, p, st);
    case 256: return dispatch_epi<256>(a, a2, a3, w, d, r, p, st);
  }
  return I360_ERR_ARG;
}

}  // namespace i360

using namespace i360;

// GEGLU packing contract: the caller packs W (and bias) so that each BN-row block holds BN/2 value
// rows followed by the BN/2 matching gate rows; i360_gemm_geglu_block(N) returns that BN.
extern "C" int i360_gemm_geglu_block(int n_total) { return pick_bn(n_total, 1); }

// tile plan of a GEMM call: N tile width and whether the 256 x 128 (two M sub-tiles) variant runs
static void plan_gemm(int M, int N, int K, int act, bool has_rowvec, bool has_resid, float out_scale, int* bn_out, bool* mt2_out) {
  int bn = pick_bn(N, act);
  // 256 x 128 tiles (two M sub-tiles sharing the weight tile) for the L2-bound mid-size projections: N a multiple of
  // 128 but not of 256 (640, 1920), K >= 512, plain / bias / bias + residual epilogues, enough rows to fill the GPU
  // Measured against the 160 / 192-wide tiles on one box: N = 640: -2..-5 % (K = 640 and 2560, with and without the
  // residual ring); N = 1920: -3 % without, +5 % with a residual (3-stage ring) -> not used there.  I360_GEMM_MT2=0 disables.
  static const bool mt2_on = getenv("I360_GEMM_MT2") == nullptr || atoi(getenv("I360_GEMM_MT2")) != 0;
  const bool mt2 = mt2_on && act == 0 && !has_rowvec && out_scale == 1.0f && (N % 128 == 0) && (N % 256 != 0) && N >= 640 && K >= 512 &&
                   M >= 256 * 148 && !(has_resid && N > 640) && !bn_override(N);
  if (mt2) bn = 128;
  *bn_out = bn; *mt2_out = mt2;
}

// number of row-statistics slots i360_gemm_rowstats_bf16 writes for this problem (the caller allocates slots * M float2)
extern "C" int i360_gemm_rowstats_slots(int M, int N, int K, int has_resid) {
  int bn; bool mt2;
  plan_gemm(M, N, K, 0, false, has_resid != 0, 1.0f, &bn, &mt2);
  const int n_tiles = (N + bn - 1) / bn;
  return mt2 ? n_tiles : 2 * n_tiles;
}

static int gemm_impl(const void* A, long long lda, const void* W, long long ldw, void* D,
                     long long ldd, int M, int N, int K, const void* bias,
                     const void* resid, long long ldr, const float* rowvec, int rowvec_div,
                     int rowvec_ld, int act, float out_scale, float* rowstats, void* stream) {
  if (!A || !W || !D || M <= 0 || N <= 0 || K <= 0) return I360_ERR_ARG;
  if ((K % 8) || (lda % 8) || (ldw % 8) || (ldd % 8) || (N % 8)) return I360_ERR_ARG;
  if (resid && (ldr % 8)) return I360_ERR_ARG;
  if (act == 1 && (N % 128)) return I360_ERR_ARG;
  if (act == 1 && (resid || rowvec)) return I360_ERR_UNSUPPORTED;
  if (rowstats && (act != 0 || rowvec || out_scale != 1.0f)) return I360_ERR_UNSUPPORTED;
  int bn; bool mt2;
  plan_gemm(M, N, K, act, rowvec != nullptr, resid != nullptr, out_scale, &bn, &mt2);
  const int bm = mt2 ? 2 * BM : BM;
  GemmConvParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.conv = 0;
  p.m_tiles = (M + bm - 1) / bm; p.n_tiles = (N + bn - 1) / bn; p.k_iters = (K + BK - 1) / BK;
  p.D = static_cast<bf16*>(D); p.ldd = ldd;
  p.bias = static_cast<const bf16*>(bias);
  p.resid = static_cast<const bf16*>(resid); p.ldr = ldr;
  p.rowvec = rowvec; p.rowvec_div = rowvec_div > 0 ? rowvec_div : 1; p.rowvec_ld = rowvec_ld;
  p.act = act; p.out_scale = out_scale; p.n_out = (act == 1) ? N / 2 : N;
  p.rowstats = reinterpret_cast<float2*>(rowstats);
  CUtensorMap ta, tw;
  uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}; uint64_t sA[1] = {(uint64_t)lda * 2};
  uint32_t bA[2] = {BK, (uint32_t)bm};
  uint64_t dW[2] = {(uint64_t)K, (uint64_t)N}; uint64_t sW[1] = {(uint64_t)ldw * 2};
  uint32_t bW[2] = {BK, (uint32_t)bn};
  int r = get_tmap_bf16(&ta, A, 2, dA, sA, bA, 3); if (r) return r;
  r = get_tmap_bf16(&tw, W, 2, dW, sW, bW, 3); if (r) return r;
  CUtensorMap td;
  uint64_t dD[2] = {(uint64_t)p.n_out, (uint64_t)M}; uint64_t sD[1] = {(uint64_t)ldd * 2};
  uint32_t bD[2] = {32, BM};
  r = get_tmap_bf16(&td, D, 2, dD, sD, bD, 2); if (r) return r;
  CUtensorMap tr = td;
  if (resid) {
    uint64_t sR[1] = {(uint64_t)ldr * 2};
    r = get_tmap_bf16(&tr, resid, 2, dD, sR, bD, 2); if (r) return r;
  }
  if (mt2) {
    if (resid) return launch<128, EPI_RESID, 2>(ta, ta, ta, tw, td, tr, p, static_cast<cudaStream_t>(stream));
    return launch<128, EPI_PLAIN, 2>(ta, ta, ta, tw, td, tr, p, static_cast<cudaStream_t>(stream));
  }
  return dispatch(bn, ta, ta, ta, tw, td, tr, p, static_cast<cudaStream_t>(stream));
}

extern "C" int i360_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* D,
                              long long ldd, int M, int N, int K, const void* bias,
                              const void* resid, long long ldr, const float* rowvec, int rowvec_div,
                              int rowvec_ld, int act, float out_scale, void* stream) {
  return gemm_impl(A, lda, W, ldw, D, ldd, M, N, K, bias, resid, ldr, rowvec, rowvec_div, rowvec_ld, act, out_scale, nullptr, stream);
}

// D = A W^T + bias (+ resid), and additionally rowstats[slot * M + r] = (sum, sum of squares) over the columns of row r
// this (N tile, epilogue group) slot stored, for slot < i360_gemm_rowstats_slots(M, N, K, resid != NULL): the producer
// half of the LayerNorm fold (i360_gemm_ln_bf16 consumes it).
extern "C" int i360_gemm_rowstats_bf16(const void* A, long long lda, const void* W, long long ldw, void* D, long long ldd,
                                       int M, int N, int K, const void* bias, const void* resid, long long ldr,
                                       float* rowstats, void* stream) {
  if (!rowstats) return I360_ERR_ARG;
  return gemm_impl(A, lda, W, ldw, D, ldd, M, N, K, bias, resid, ldr, nullptr, 1, 0, 0, 1.0f, rowstats, stream);
}

// HALO mode (fixed 16 x 8 pixel tiles) is used when its padded work is within 4 % of the best free-form pixel box, the
// input has at least one full 64-channel block, the epilogue is one the halo kernels are instantiated for and there is no
// fused 1x1 source: those k-steps need a fresh activation tile each, and the two-deep activation ring then exposes the TMA
// latency (measured: conv2 + shortcut of the 16x512x1024 step 6 % slower).  Measured on one B200 (tools/microbench.py
// conv): 960->320 at 640x32x32 +7 %, pano 320->320 +6 %, VAE 256->256 +12 %, 512->512 +11 %, 320->320 at 32x32 +-0 %,
// 640->640 at 32x68 (6 % padded columns) -9 % -> excluded by the 4 % rule.  Inside the 16x512x1024 step (inputs warm in L2
// after the GroupNorm pass) the 32x32 / 16x16 perspective convs gained nothing (75.9 ms of convs with, 75.1 ms without), the
// VAE decode of 4 frames went 24.8 -> 23.4 ms: the rule therefore also asks for images of at least 64 x 64 pixels (the
// VAE, panorama level 0).  I360_CONV_HALO=0 disables it (A/B runs).
struct HaloPolicy { int on; double tol; int allow_extra; int min_hw; };
static HaloPolicy& halo_policy() {
  static HaloPolicy p = {getenv("I360_CONV_HALO") == nullptr || atoi(getenv("I360_CONV_HALO")) != 0,
                         getenv("I360_CONV_HALO_TOL") ? atof(getenv("I360_CONV_HALO_TOL")) : 1.04,
                         getenv("I360_CONV_HALO_EXTRA") != nullptr && atoi(getenv("I360_CONV_HALO_EXTRA")) != 0,
                         getenv("I360_CONV_HALO_MIN_HW") ? atoi(getenv("I360_CONV_HALO_MIN_HW")) : 64};
  return p;
}
static bool use_halo(int H, int W, int Cin, bool resid, bool rowvec, bool extra_sources, double waste_best) {
  const HaloPolicy& hp = halo_policy();
  const double waste_h = (double)((W + kHaloTW - 1) / kHaloTW * kHaloTW) * ((H + kHaloTH - 1) / kHaloTH * kHaloTH) / ((double)W * H);
  return hp.on && !(resid && rowvec) && (hp.allow_extra || !extra_sources) && Cin >= 64 && H >= hp.min_hw && W >= hp.min_hw &&
         waste_h <= hp.tol * waste_best;
}

// Selection rule of the halo kernels (see use_halo): on / padded-work tolerance / allow fused 1x1 sources / smallest image
// side; a negative
// argument keeps the current value.  For tests (which widen the rule to cover every halo kernel path) and A/B runs.
extern "C" void i360_conv3x3_halo_policy(int on, double tol, int allow_extra, int min_hw) {
  HaloPolicy& hp = halo_policy();
  if (min_hw >= 0) hp.min_hw = min_hw;
  if (on >= 0) hp.on = on;
  if (tol >= 0) hp.tol = tol;
  if (allow_extra >= 0) hp.allow_extra = allow_extra;
}

static void best_box(int B, int H, int W, int* TW, int* TH, int* TB, double* waste, int max_tb = 128) {
  int bestTW = 16, bestTH = 8, bestTB = 1; double bestw = 1e30;
  for (int tw = 1; tw <= 128; tw *= 2)
    for (int th = 1; tw * th <= 128; th *= 2) {
      const int tb = 128 / (tw * th);
      if (tb > max_tb) continue;
      if (tw > 1 && tw / 2 >= W) continue;
      if (th > 1 && th / 2 >= H) continue;
      const double w = (double)((W + tw - 1) / tw * tw) * ((H + th - 1) / th * th) *
                       ((B + tb - 1) / tb * tb) / ((double)W * H * B);
      const double score = w - 1e-6 * tw;  // ties -> wider rows
      if (score < bestw) { bestw = score; bestTW = tw; bestTH = th; bestTB = tb; }
    }
  *TW = bestTW; *TH = bestTH; *TB = bestTB; *waste = bestw;
}

// 1 when i360_conv3x3_bf16 would run this problem through the halo-tile kernels (tests / benchmarks ask)
extern "C" int i360_conv3x3_uses_halo(int B, int H, int W, int Cin, int has_resid, int has_rowvec, int has_extra_sources) {
  int tw, th, tb; double w;
  best_box(B, H, W, &tw, &th, &tb, &w);
  return use_halo(H, W, Cin, has_resid != 0, has_rowvec != 0, has_extra_sources != 0, w + 1e-6 * tw) ? 1 : 0;
}

// x: NHWC [B,H,W,Cin] (H,W include any materialised halo); Wt: [Cout, 9*Cin + C2 + C3] with the
// 3x3 part ordered (kh, kw, cin). x2/x3: optional NHWC [B,H,W-2*crop,C2|C3] sources for a fused 1x1
// (stored WITHOUT the halo, i.e. aligned with the output). Output NHWC [B, H, W-2*crop, Cout].
// Output placement of a convolution: dense NHWC [B, H, Wo, Cout] (lattice == nullptr), or a strided lattice inside a larger
// NHWC tensor (the parity sub-lattices of the 2x-upsampled output): element strides of one step in w / h / image.
struct OutLattice { long long s_w, s_h, s_b; };

static int conv_impl(const void* x, int B, int H, int W, int Cin, const void* x2, int C2, const void* x3, int C3, const void* Wt,
                     int Cout, void* D, const OutLattice* lattice, int crop, const void* bias, const void* resid, const float* rowvec,
                     int rowvec_div, int rowvec_ld, float out_scale, int n_taps, const signed char* tap_dh, const signed char* tap_dw,
                     void* stream, int in_stride = 1, int Hin = 0, int Win = 0, double* gn_stats = nullptr, int gn_groups = 0,
                     double* gn_chan = nullptr) {
  // H, W: the OUTPUT domain that is tiled (before the crop); Hin, Win: the input tensor when in_stride != 1
  if (in_stride == 1) { Hin = H; Win = W; }
  if (!x || !Wt || !D || B <= 0 || H <= 0 || W <= 0) return I360_ERR_ARG;
  if ((Cin % 8) || (Cout % 8) || (C2 % 8) || (C3 % 8) || crop < 0 || 2 * crop >= W)
    return I360_ERR_ARG;
  if ((C2 > 0 && !x2) || (C3 > 0 && !x3)) return I360_ERR_ARG;
  const bool is3x3 = (n_taps == 9);          // the halo kernels hard-wire the nine (kh-1, kw-1) taps
  // K segments need no 64-alignment: channels past a source's extent are zero-filled by TMA, so whatever
  // weight columns a partial K block overlaps are multiplied by zeros.
  // choose the pixel box minimising padded work
  int bestTW, bestTH, bestTB; double bestw;
  if (gn_stats) {       // one image per tile; groups of 4 / 8 / 16 channels never straddle a 32-column chunk
    if (gn_groups != 32 || (Cout % 32) || (Cout / 32 != 4 && Cout / 32 != 8 && Cout / 32 != 16) || rowvec) return I360_ERR_UNSUPPORTED;
    if (static_cast<long long>(H) * W < 128) return I360_ERR_UNSUPPORTED;
  }
  // per-channel statistics reduce over a warp's 32 tile rows: those must lie in one image (TW * TH >= 32)
  if (gn_chan && static_cast<long long>(H) * W < 32) return I360_ERR_UNSUPPORTED;
  best_box(B, H, W, &bestTW, &bestTH, &bestTB, &bestw, gn_stats ? 1 : (gn_chan ? 4 : 128));
  GemmConvParams p;
  memset(&p, 0, sizeof(p));
  p.gn_stats = gn_stats; p.gn_gs = gn_stats ? Cout / 32 : 0;
  p.gn_chan = gn_chan;
  p.halo = is3x3 && !lattice && in_stride == 1 && use_halo(H, W, Cin, resid != nullptr, rowvec != nullptr, C2 + C3 > 0, bestw + 1e-6 * bestTW);
  p.in_stride = in_stride;
  if (in_stride != 1 && (C2 > 0 || C3 > 0)) return I360_ERR_UNSUPPORTED;
  if (p.halo) { bestTW = kHaloTW; bestTH = kHaloTH; bestTB = 1; }
  p.conv = 1; p.B = B; p.H = H; p.W = W; p.TW = bestTW; p.TH = bestTH; p.TB = bestTB;
  p.n_wt = (W + p.TW - 1) / p.TW; p.n_ht = (H + p.TH - 1) / p.TH; p.n_bt = (B + p.TB - 1) / p.TB;
  p.Cin = Cin; p.C2 = C2; p.C3 = C3; p.crop = crop; p.xoff = -crop; p.Hout = H; p.Wout = W - 2 * crop;
  p.N = Cout; p.M = p.n_wt * p.n_ht * p.n_bt * BM;
  p.n_taps = n_taps;
  for (int t = 0; t < n_taps; ++t) { p.tap_dh[t] = tap_dh[t]; p.tap_dw[t] = tap_dw[t]; }
  const int bn = pick_bn(Cout, 0);
  p.m_tiles = p.n_wt * p.n_ht * p.n_bt; p.n_tiles = (Cout + bn - 1) / bn;
  p.k_iters = n_taps * ((Cin + BK - 1) / BK) + (C2 + BK - 1) / BK + (C3 + BK - 1) / BK;
  p.D = static_cast<bf16*>(D); p.ldd = Cout;
  p.bias = static_cast<const bf16*>(bias);
  p.resid = static_cast<const bf16*>(resid); p.ldr = Cout;
  p.rowvec = rowvec; p.rowvec_div = rowvec_div > 0 ? rowvec_div : 1; p.rowvec_ld = rowvec_ld;
  p.act = 0; p.out_scale = out_scale; p.n_out = Cout;
  if (lattice) { p.os_w = lattice->s_w; p.os_h = lattice->s_h; p.os_b = lattice->s_b; }
  const long long Ktot = static_cast<long long>(n_taps) * Cin + C2 + C3;
  CUtensorMap ta, ta2, ta3, tw;
  auto act_map = [&](CUtensorMap* m, const void* ptr, int C, int Wt_) {
    uint64_t d[4] = {(uint64_t)C, (uint64_t)Wt_, (uint64_t)H, (uint64_t)B};
    uint64_t s[3] = {(uint64_t)C * 2, (uint64_t)Wt_ * C * 2, (uint64_t)H * Wt_ * C * 2};
    uint32_t b[4] = {BK, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TB};
    return get_tmap_bf16(m, ptr, 4, d, s, b, 3);
  };
  int r;
  if (p.halo) {
    uint64_t d[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t s[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t b[4] = {BK, (uint32_t)kHaloW, (uint32_t)kHaloH, 1u};
    r = get_tmap_bf16(&ta, x, 4, d, s, b, 3);
  } else if (in_stride != 1) {
    // traversal stride along w and h: the box spans in_stride * T input pixels and loads every in_stride-th one
    uint64_t d[4] = {(uint64_t)Cin, (uint64_t)Win, (uint64_t)Hin, (uint64_t)B};
    uint64_t s[3] = {(uint64_t)Cin * 2, (uint64_t)Win * Cin * 2, (uint64_t)Hin * Win * Cin * 2};
    uint32_t b[4] = {BK, (uint32_t)(p.TW * in_stride), (uint32_t)(p.TH * in_stride), (uint32_t)p.TB};
    uint32_t es[4] = {1u, (uint32_t)in_stride, (uint32_t)in_stride, 1u};
    if (b[1] > 256 || b[2] > 256) return I360_ERR_UNSUPPORTED;
    r = get_tmap_bf16(&ta, x, 4, d, s, b, 3, es);
  } else {
    r = act_map(&ta, x, Cin, W);
  }
  if (r) return r;
  ta2 = ta; ta3 = ta;
  if (C2 > 0) { r = act_map(&ta2, x2, C2, W - 2 * crop); if (r) return r; }
  if (C3 > 0) { r = act_map(&ta3, x3, C3, W - 2 * crop); if (r) return r; }
  uint64_t dW[2] = {(uint64_t)Ktot, (uint64_t)Cout}; uint64_t sW[1] = {(uint64_t)Ktot * 2};
  uint32_t bW[2] = {BK, (uint32_t)bn};
  r = get_tmap_bf16(&tw, Wt, 2, dW, sW, bW, 3); if (r) return r;
  CUtensorMap td, tr;
  {
    const int Wo = W - 2 * crop;
    uint64_t d[4] = {(uint64_t)Cout, (uint64_t)Wo, (uint64_t)H, (uint64_t)B};
    uint64_t s[3] = {(uint64_t)Cout * 2, (uint64_t)Wo * Cout * 2, (uint64_t)H * Wo * Cout * 2};
    if (lattice) { s[0] = (uint64_t)lattice->s_w * 2; s[1] = (uint64_t)lattice->s_h * 2; s[2] = (uint64_t)lattice->s_b * 2; }
    uint32_t b[4] = {32, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TB};
    r = get_tmap_bf16(&td, D, 4, d, s, b, 2); if (r) return r;
    tr = td;
    if (resid) {
      if (lattice) return I360_ERR_UNSUPPORTED;
      r = get_tmap_bf16(&tr, resid, 4, d, s, b, 2); if (r) return r;
    }
  }
  return dispatch(bn, ta, ta2, ta3, tw, td, tr, p, static_cast<cudaStream_t>(stream));
}

extern "C" int i360_conv3x3_bf16(const void* x, int B, int H, int W, int Cin, const void* x2, int C2,
                                 const void* x3, int C3, const void* Wt, int Cout, void* D,
                                 int crop, const void* bias, const void* resid,
                                 const float* rowvec, int rowvec_div, int rowvec_ld,
                                 float out_scale, void* stream) {
  static const signed char dh[9] = {-1, -1, -1, 0, 0, 0, 1, 1, 1}, dw[9] = {-1, 0, 1, -1, 0, 1, -1, 0, 1};
  return conv_impl(x, B, H, W, Cin, x2, C2, x3, C3, Wt, Cout, D, nullptr, crop, bias, resid, rowvec, rowvec_div, rowvec_ld,
                   out_scale, 9, dh, dw, stream);
}

// i360_conv3x3_bf16 that ALSO accumulates the GroupNorm statistics of its output (32 groups of 4 / 8 / 16 channels: the VAE's
// widths) from the epilogue: stats[(image * 32 + group) * 2 + {0, 1}] = (sum, sum of squares) over the image, fp64, zeroed
// here.  The GroupNorm that reads the output then needs no statistics pass (i360_groupnorm_apply takes the same buffer).
// Replaces conv -> GroupNorm's first read (diffusers/models/resnet.py:454-496 ResnetBlock2D: norm2 after conv1, the next
// block's norm1 after conv2 / the up-sampler).
extern "C" int i360_conv3x3_gnstats_bf16(const void* x, int B, int H, int W, int Cin, const void* x2, int C2, const void* Wt,
                                         int Cout, void* D, const void* bias, const void* resid, int groups, double* stats,
                                         void* stream) {
  if (!stats) return I360_ERR_ARG;
  if (cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * 32, static_cast<cudaStream_t>(stream)) != cudaSuccess) return I360_ERR_CUDA;
  static const signed char dh[9] = {-1, -1, -1, 0, 0, 0, 1, 1, 1}, dw[9] = {-1, 0, 1, -1, 0, 1, -1, 0, 1};
  return conv_impl(x, B, H, W, Cin, x2, C2, nullptr, 0, Wt, Cout, D, nullptr, 0, bias, resid, nullptr, 1, 0, 1.0f, 9, dh, dw, stream,
                   1, 0, 0, stats, groups);
}

// i360_conv3x3_bf16 (same arguments) that also accumulates PER-CHANNEL statistics of its output from the epilogue:
// chan_stats[(image * Cout + channel) * 2 + {0, 1}] = (sum, sum of squares) of the stored bf16 values over the image, fp64,
// zeroed here.  Any channel count / group size: the GroupNorm reading the output folds channels into groups
// (i360_groupnorm_apply_chanstats) and needs no statistics pass.  Replaces the first read of InflatedGroupNorm after a conv in
// ResnetBlock3D (norm2 after conv1, animatediff/models/resnet.py:243) and of Transformer3DModel.norm after conv2
// (animatediff/models/attention.py:262).
extern "C" int i360_conv3x3_chanstats_bf16(const void* x, int B, int H, int W, int Cin, const void* x2, int C2, const void* x3,
                                           int C3, const void* Wt, int Cout, void* D, int crop, const void* bias,
                                           const void* resid, const float* rowvec, int rowvec_div, int rowvec_ld, float out_scale,
                                           double* chan_stats, void* stream) {
  if (!chan_stats) return I360_ERR_ARG;
  if (cudaMemsetAsync(chan_stats, 0, sizeof(double) * 2 * B * Cout, static_cast<cudaStream_t>(stream)) != cudaSuccess) return I360_ERR_CUDA;
  static const signed char dh[9] = {-1, -1, -1, 0, 0, 0, 1, 1, 1}, dw[9] = {-1, 0, 1, -1, 0, 1, -1, 0, 1};
  return conv_impl(x, B, H, W, Cin, x2, C2, x3, C3, Wt, Cout, D, nullptr, crop, bias, resid, rowvec, rowvec_div, rowvec_ld, out_scale,
                   9, dh, dw, stream, 1, 0, 0, nullptr, 0, chan_stats);
}

// 3x3 / stride 2 convolution as an implicit GEMM: the activation box of tap (kh, kw) is a TMA box with traversal stride 2
// starting at (2 w0 + kw - pad_lo, 2 h0 + kh - pad_lo); rows / columns outside the image are zero-filled = the conv's padding
// (pad_lo = 1: symmetric pad 1 of Downsample3D, animatediff/models/resnet.py:117-140; pad_lo = 0: the VAE's asymmetric
// F.pad(0, 1, 0, 1), diffusers/models/resnet.py:184).  x: NHWC [B, Hin, Win, Cin] (Win includes 2 * crop circular halo
// columns per side), D: NHWC [B, Hin / 2, Win / 2 - 2 crop, Cout].  Replaces the im2col buffer + GEMM of round 1.
extern "C" int i360_conv3x3_s2_bf16(const void* x, int B, int Hin, int Win, int Cin, const void* Wt, int Cout, void* D,
                                    int pad_lo, int crop, const void* bias, void* stream) {
  if (pad_lo < 0 || pad_lo > 1 || (Hin % 2) || (Win % 2)) return I360_ERR_ARG;
  signed char dh[9], dw[9];
  for (int t = 0; t < 9; ++t) { dh[t] = static_cast<signed char>(t / 3 - pad_lo); dw[t] = static_cast<signed char>(t % 3 - pad_lo); }
  return conv_impl(x, B, Hin / 2, Win / 2, Cin, nullptr, 0, nullptr, 0, Wt, Cout, D, nullptr, crop, bias, nullptr, nullptr, 1, 0, 1.0f,
                   9, dh, dw, stream, 2, Hin, Win);
}

// Nearest 2x upsample followed by a 3x3 / pad 1 convolution, WITHOUT the upsampled tensor and with 4/9 of the multiply-adds:
// output pixel (2i + a, 2j + b) reads input rows {i-1, i} (a = 0) or {i, i+1} (a = 1) and likewise for columns, so each of the
// four output parities is a 2x2-tap convolution of the LOW-resolution input whose weights are sums of the 3x3 taps that fall
// on the same input pixel (a = 0: {w[0], w[1] + w[2]}, a = 1: {w[0] + w[1], w[2]}; the caller pre-sums in fp32 and rounds once:
// ops.pack_upsample_conv).  x: NHWC [B, H, W, Cin] low resolution (W includes `crop` circular halo columns per side);
// Weff: [4 parities (a*2+b)][Cout][4 taps (dr*2+dc) * Cin]; D: NHWC [B, 2H, 2(W - 2 crop), Cout].  Zero padding of the
// upsampled image coincides with TMA's zero fill of the low-resolution box.  Replaces Upsample3D.forward = F.interpolate(
// scale_factor=2, mode="nearest") + InflatedConv3d (animatediff/models/resnet.py:86-114; with pad_pano(1) / unpad_pano(2) of
// MVGenModel.py:449-456 as crop = 1) and the VAE's Upsample2D (diffusers/models/resnet.py:108-143).
static int upsample2x_impl(const void* x, int B, int H, int W, int Cin, const void* Weff, int Cout, void* D, int crop,
                           const void* bias, double* stats, int groups, void* stream) {
  if (!x || !Weff || !D) return I360_ERR_ARG;
  const int Wo = W - 2 * crop;
  const OutLattice lat = {2LL * Cout, 2LL * (2LL * Wo) * Cout, (2LL * H) * (2LL * Wo) * Cout};
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      const signed char dh[4] = {static_cast<signed char>(a - 1), static_cast<signed char>(a - 1), static_cast<signed char>(a), static_cast<signed char>(a)};
      const signed char dw[4] = {static_cast<signed char>(b - 1), static_cast<signed char>(b), static_cast<signed char>(b - 1), static_cast<signed char>(b)};
      const bf16* wp = static_cast<const bf16*>(Weff) + static_cast<long long>(a * 2 + b) * Cout * (4LL * Cin);
      bf16* dp = static_cast<bf16*>(D) + (static_cast<long long>(a) * (2LL * Wo) + b) * Cout;
      const int r = conv_impl(x, B, H, W, Cin, nullptr, 0, nullptr, 0, wp, Cout, dp, &lat, crop, bias, nullptr, nullptr, 1, 0, 1.0f,
                              4, dh, dw, stream, 1, 0, 0, stats, groups);
      if (r) return r;
    }
  return I360_OK;
}

extern "C" int i360_conv_upsample2x_bf16(const void* x, int B, int H, int W, int Cin, const void* Weff, int Cout, void* D,
                                         int crop, const void* bias, void* stream) {
  return upsample2x_impl(x, B, H, W, Cin, Weff, Cout, D, crop, bias, nullptr, 0, stream);
}

// ... with the GroupNorm statistics of the (2H x 2W) output accumulated by the four parity launches (see
// i360_conv3x3_gnstats_bf16; the image index of a tile is the same in all four).
extern "C" int i360_conv_upsample2x_gnstats_bf16(const void* x, int B, int H, int W, int Cin, const void* Weff, int Cout, void* D,
                                                 const void* bias, int groups, double* stats, void* stream) {
  if (!stats) return I360_ERR_ARG;
  if (cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * 32, static_cast<cudaStream_t>(stream)) != cudaSuccess) return I360_ERR_CUDA;
  return upsample2x_impl(x, B, H, W, Cin, Weff, Cout, D, 0, bias, stats, groups, stream);
}

// LayerNorm folded into the projection that consumes it:  D = act( LN(A; gamma, beta, eps) W^T + bias (+ rowvec) ).
// Wf: [N, K] = W * gamma (bf16);  u[n] = sum_k Wf[n,k] (fp32, from the bf16-rounded Wf);  c[n] = sum_k beta[k] W[n,k] + bias[n]
// (fp32).  K must be the LayerNorm width (the whole row), a multiple of 64.  act: 0 none, 1 GEGLU (Wf / u / c packed with
// i360_gemm_geglu_block like the un-folded weight).  rowvec (fp32 [rowvec_mod or M/div, N]) is added per row at
// (row / rowvec_div) % rowvec_mod -- the temporal module's PE term (LN(x) + pe_f) W^T = LN(x) W^T + (pe W^T)[f].
// Replaces nn.LayerNorm followed by nn.Linear (animatediff/models/attention.py:463-508, motion_module.py:247-259,:350).
extern "C" int i360_gemm_ln_bf16(const void* A, long long lda, const void* Wf, long long ldw, void* D, long long ldd, int M,
                                 int N, int K, const float* u, const float* c, float eps, const float* rowstats, int slots,
                                 const float* rowvec, int rowvec_div, int rowvec_mod, int rowvec_ld, int act, void* stream) {
  if (!A || !Wf || !D || !u || !c || !rowstats || slots <= 0 || slots > kMaxStatSlots || M <= 0 || N <= 0 || K <= 0) return I360_ERR_ARG;
  if ((K % 64) || (lda % 8) || (ldw % 8) || (ldd % 8) || (N % 8)) return I360_ERR_ARG;
  if (act != 0 && act != 1) return I360_ERR_UNSUPPORTED;
  if (act == 1 && ((N % 256) || rowvec)) return I360_ERR_UNSUPPORTED;
  const int bn = pick_bn(N, act);
  if (bn < 160) return I360_ERR_UNSUPPORTED;        // narrow outputs keep the separate LayerNorm pass (host decides)
  GemmConvParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.conv = 0;
  p.m_tiles = (M + BM - 1) / BM; p.n_tiles = (N + bn - 1) / bn; p.k_iters = K / BK;
  p.D = static_cast<bf16*>(D); p.ldd = ldd;
  p.rowvec = rowvec; p.rowvec_div = rowvec_div > 0 ? rowvec_div : 1; p.rowvec_ld = rowvec_ld; p.rowvec_mod = rowvec_mod;
  p.act = act; p.out_scale = 1.0f; p.n_out = (act == 1) ? N / 2 : N;
  p.ln_u = u; p.ln_c = c; p.ln_eps = eps; p.ln_stats = reinterpret_cast<const float2*>(rowstats); p.ln_slots = slots;
  CUtensorMap ta, tw, td;
  uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}; uint64_t sA[1] = {(uint64_t)lda * 2};
  uint32_t bA[2] = {BK, (uint32_t)BM};
  uint64_t dW[2] = {(uint64_t)K, (uint64_t)N}; uint64_t sW[1] = {(uint64_t)ldw * 2};
  uint32_t bW[2] = {BK, (uint32_t)bn};
  int r = get_tmap_bf16(&ta, A, 2, dA, sA, bA, 3); if (r) return r;
  r = get_tmap_bf16(&tw, Wf, 2, dW, sW, bW, 3); if (r) return r;
  uint64_t dD[2] = {(uint64_t)p.n_out, (uint64_t)M}; uint64_t sD[1] = {(uint64_t)ldd * 2};
  uint32_t bD[2] = {32, BM};
  r = get_tmap_bf16(&td, D, 2, dD, sD, bD, 2); if (r) return r;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (act == 1) {
    if (four_groups()) return launch<256, EPI_GEGLU, 1, true, false, 4>(ta, ta, ta, tw, td, td, p, st);
    return launch<256, EPI_GEGLU, 1, true>(ta, ta, ta, tw, td, td, p, st);
  }
  // temporal PE term: when the 128 rows of a tile share one frame the vector rides in the tile's bias table (plain epilogue)
  p.rowvec_in_table = (rowvec && rowvec_mod > 0 && (p.rowvec_div % BM) == 0) ? 1 : 0;
  if (rowvec && !p.rowvec_in_table) {
    switch (bn) {
      case 160: return launch<160, EPI_ROWVEC, 1, true>(ta, ta, ta, tw, td, td, p, st);
      case 192: return launch<192, EPI_ROWVEC, 1, true>(ta, ta, ta, tw, td, td, p, st);
      case 256: return launch<256, EPI_ROWVEC, 1, true>(ta, ta, ta, tw, td, td, p, st);
    }
  } else if (four_groups_plain()) {
    switch (bn) {
      case 160: return launch<160, EPI_PLAIN, 1, true, false, 4>(ta, ta, ta, tw, td, td, p, st);
      case 192: return launch<192, EPI_PLAIN, 1, true, false, 4>(ta, ta, ta, tw, td, td, p, st);
      case 256: return launch<256, EPI_PLAIN, 1, true, false, 4>(ta, ta, ta, tw, td, td, p, st);
    }
  } else {
    switch (bn) {
      case 160: return launch<160, EPI_PLAIN, 1, true>(ta, ta, ta, tw, td, td, p, st);
      case 192: return launch<192, EPI_PLAIN, 1, true>(ta, ta, ta, tw, td, td, p, st);
      case 256: return launch<256, EPI_PLAIN, 1, true>(ta, ta, ta, tw, td, td, p, st);
    }
  }
  return I360_ERR_UNSUPPORTED;
}

// tile width i360_gemm_ln_bf16 would use for N output columns (0: not supported, keep LayerNorm + GEMM)
extern "C" int i360_gemm_ln_supported(int N, int K, int act) {
  if ((K % 64) || (N % 8)) return 0;
  if (act == 1) return (N % 256 == 0) ? 256 : 0;
  const int bn = pick_bn(N, 0);
  return bn >= 160 ? bn : 0;
}
