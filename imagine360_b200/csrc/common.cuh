// Shared device-side helpers for the sm_100a kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (MMA / TMEM alloc / TMEM load / commit / fences) as thin inline-PTX wrappers, plus the
// shared-memory matrix descriptor and instruction descriptor encoders.
//
// Everything here is written for sm_100a only (B200). No fallbacks.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace i360 {

typedef __nv_bfloat16 bf16;

#ifndef I360_SPIN_LIMIT
#define I360_SPIN_LIMIT (1u << 27)   // ~seconds; a stuck pipeline traps instead of hanging the GPU
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05 reading smem)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// A failed invariant inside a kernel traps (a device printf in these register-tight kernels costs stack frames and spills:
// measured with ptxas -v).  The two invariants are: dynamic shared memory 1024-byte aligned (128B-swizzled TMA / UMMA tiles),
// and no mbarrier wait longer than I360_SPIN_LIMIT polls (pipeline deadlock guard).  The host names the failing entry point
// (tmap.h::I360_CUDA_CHECK_LAUNCH) when the sticky error surfaces.
#define i360_device_fail(what) __trap()
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"   // %3: suspend-time hint (ns): the
      "selp.u32 %0, 1, 0, p;\n\t}\n"                                     // warp sleeps in hardware instead of
      : "=r"(ok)                                                            // burning issue slots in a spin loop
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > I360_SPIN_LIMIT) i360_device_fail("mbarrier wait exceeded the spin limit (pipeline deadlock guard)");
  }
}

// ----------------------------------------------------------------------------------------------
// TMA tiled loads (global -> shared), completion on an mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA tiled stores (shared -> global), bulk async-group completion; out-of-bounds elements are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL) -- an EXPERIMENT that is off by default (tmap.h::pdl_enabled holds the numbers: it made
// the step 0.4-2 % slower).  With I360_PDL=1 every kernel of the library is launched with the programmatic-stream-serialization
// attribute (tmap.h::launch_k), executes griddep_wait() BEFORE its first access to global memory that an earlier kernel
// may have written (or may still be reading), and griddep_launch() right after its own set-up: the next kernel's CTAs are
// then scheduled onto SMs as this grid's CTAs retire and run THEIR set-up (barrier init, TMEM allocation, descriptor
// prefetch, ~2-4 us with the launch latency) under this grid's tail instead of after it.  griddep_wait() returns only when
// the prerequisite grid has completed and its memory operations are visible, so data dependencies are exactly those of
// plain stream order; without the launch attribute both instructions are no-ops.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef I360_PDL_EARLY_TRIGGER
#define I360_PDL_EARLY_TRIGGER 1      // 0: no explicit trigger (dependents are released when the CTAs exit)
#endif
__device__ __forceinline__ void griddep_launch() {
#if I360_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, TMEM loads, fences
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread retired
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base_lane+i)
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// registers -> TMEM, same lane / column mapping as tmem_ld_x16
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 "
      "[%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
         "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// Descriptors (bit layouts per cute/arch/mma_sm100_desc.hpp: SmemDescriptor / InstrDescriptor)
// ----------------------------------------------------------------------------------------------
enum : uint32_t { SWZ_NONE = 0, SWZ_128B = 2, SWZ_64B = 4, SWZ_32B = 6 };

// K-major operand tile whose rows are exactly one swizzle span wide (128B/64B/32B):
//   8-row core groups are `sbo_bytes` apart (8 * row bytes for a dense tile).
// MN-major operand tile ([k rows][mn span]): same encoding, sbo = stride between 8-k-row groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t sbo_bytes,
                                                   uint32_t lbo_bytes, uint32_t swizzle) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(swizzle & 7) << 61;
  return d;
}
// bf16 x bf16 -> fp32, M rows x N cols per instruction (K = 16)
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                       uint32_t b_mn_major) {
  return (1u << 4)            // D format: F32
         | (1u << 7)          // A format: BF16
         | (1u << 10)         // B format: BF16
         | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------------
// small math / packing helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
// gelu(x) = 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, i.e. far
// below the bf16 rounding of the result): 2 MUFU + ~12 FMA instead of erff's ~30-instruction expansion.  The
// epilogue of the GEGLU GEMMs evaluates this once per output element, so its cost bounds the small-K layers.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = fast_rcp(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = 1.0f - poly * t * __expf(-z * z);       // erf(|x| / sqrt 2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}
// Blackwell packed fp32 pairs (FFMA2 / FADD2) and the 3-input max (FMNMX3): halve the issue slots of the
// softmax inner loops, which are instruction-bound next to the tensor pipe.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^t for a pair without the MUFU: round-to-nearest split t = n + f (magic-number add), degree-3 minimax
// polynomial of 2^f on [-0.5, 0.5] (max rel. error 7.5e-5, well under the bf16 rounding of P) on packed FFMA2, and
// n added straight into the exponent field.  t is clamped at -126 (result ~1e-38, i.e. 0 for the row sum).
__device__ __forceinline__ float2 exp2_poly2(float2 t) {
  const float kMagic = 12582912.0f;               // 1.5 * 2^23: the low mantissa bits of (t + kMagic) hold round(t)
  const float2 tc = make_float2(fmaxf(t.x, -126.0f), fmaxf(t.y, -126.0f));
  const float2 r = fadd2(tc, make_float2(kMagic, kMagic));
  const float2 n = fadd2(r, make_float2(-kMagic, -kMagic));
  const float2 f = ffma2(n, make_float2(-1.f, -1.f), tc);
  float2 q = ffma2(f, make_float2(0.0551716685f, 0.0551716685f), make_float2(0.2426111251f, 0.2426111251f));
  q = ffma2(q, f, make_float2(0.6932609677f, 0.6932609677f));
  q = ffma2(q, f, make_float2(0.9999280572f, 0.9999280572f));
  return make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(r.x) << 23)),
                     __int_as_float(__float_as_int(q.y) + (__float_as_int(r.y) << 23)));
}

// gelu(x) = 0.5 x (1 + erf(x / sqrt 2)) on two values at once, for the GEGLU epilogue, which is MUFU / issue bound.
// erf from Abramowitz-Stegun 7.1.28:  erf(z) = 1 - 1 / (1 + a1 z + ... + a6 z^6)^16,  z >= 0,  |error| <= 3e-7
// (far below the bf16 rounding of the result): ONE MUFU op (rcp) per element instead of rcp + ex2, the rest is packed
// FFMA2 / FMUL2 (6-term Horner with 1/sqrt2 folded into the coefficients, four squarings).  With r = 1 - erf(|x|/sqrt2):
// gelu(x) = max(x, 0) - 0.5 |x| r  for either sign of x.  Large |x|: the power overflows to +inf, r = 0, exact limit.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  float2 p = ffma2(ax, make_float2(5.38297500e-6f, 5.38297500e-6f), make_float2(4.88906061e-5f, 4.88906061e-5f));   // a6/8, a5/(4 sqrt2)
  p = ffma2(p, ax, make_float2(3.80035750e-5f, 3.80035750e-5f));        // a4 / 4
  p = ffma2(p, ax, make_float2(3.27762634e-3f, 3.27762634e-3f));        // a3 / (2 sqrt2)
  p = ffma2(p, ax, make_float2(2.11410062e-2f, 2.11410062e-2f));        // a2 / 2
  p = ffma2(p, ax, make_float2(4.98673469e-2f, 4.98673469e-2f));        // a1 / sqrt2
  p = ffma2(p, ax, make_float2(1.f, 1.f));
  p = fmul2(p, p); p = fmul2(p, p); p = fmul2(p, p); p = fmul2(p, p);   // ^16
  const float2 r = make_float2(fast_rcp(p.x), fast_rcp(p.y));
  const float2 m = fmul2(ax, make_float2(0.5f, 0.5f));
  const float2 h = ffma2(x, make_float2(0.5f, 0.5f), m);                 // max(x, 0) = 0.5 x + 0.5 |x|, exact
  return ffma2(make_float2(-m.x, -m.y), r, h);
}
// EXPERIMENT (not the default, see I360_GEGLU_ERF): a cheaper gelu(x) = x * Phi(x) for the GEGLU epilogue (SASS: 304 packed
// FMA-pipe instructions per 32-column chunk with the A-S 7.1.28 erf above, 19 per output pair):
// Phi(x) ~= sigmoid(x (c1 + c3 x^2 + c5 x^4)), odd quintic
// fitted by minimax on the GELU error (tools/fit_gelu_sigmoid.py): max |error| 2.6e-5 on the whole real line in fp32 --
// two orders of magnitude below the bf16 rounding of the result -- for 6 FMA-pipe instructions and 4 MUFU ops (ex2, rcp)
// per pair instead of 14 and 2.  The quintic turns around past |x| ~ 11, so its argument is clamped to +-10 (ALU pipe);
// the result still multiplies the unclamped x.  Constants carry the -log2(e) of exp(-p) = 2^(-p log2 e).
// Measured: 655 360 x 2560 x 320 LayerNorm-folded GEGLU 1.205 ms with it, 1.215 ms with the erf -- the epilogue is bound by
// dependency latency (ncu: stall "wait" 27 %, FMA pipe 33 %, tensor 46 %; profiles/r02b_ncu_geglu_ln_k320.txt), not by the
// FMA pipe, so the 40 % fewer FMA instructions buy nothing and the 100 x more accurate erf stays the default.
__device__ __forceinline__ float2 gelu_sig2(float2 x) {
  const float2 xc = make_float2(fminf(fmaxf(x.x, -10.f), 10.f), fminf(fmaxf(x.y, -10.f), 10.f));
  const float2 x2 = fmul2(xc, xc);
  float2 p = ffma2(x2, make_float2(1.0142630198970437e-3f, 1.0142630198970437e-3f), make_float2(-0.10677571594715118f, -0.10677571594715118f));
  p = ffma2(p, x2, make_float2(-2.301121234893799f, -2.301121234893799f));
  const float2 a = fmul2(p, xc);
  const float2 d = fadd2(make_float2(fast_exp2(a.x), fast_exp2(a.y)), make_float2(1.f, 1.f));
  return fmul2(x, make_float2(fast_rcp(d.x), fast_rcp(d.y)));
}
#ifndef I360_GEGLU_ERF
#define I360_GEGLU_ERF 1      // 0: the GEGLU epilogue uses gelu_sig2 (A/B builds)
#endif
__device__ __forceinline__ float2 gelu_gate2(float2 x) { return I360_GEGLU_ERF ? gelu_erf2(x) : gelu_sig2(x); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }
// Packed SiLU for the HBM-bound GroupNorm+SiLU pass: ONE MUFU op per element (ex2 of -|y|) and the reciprocal of
// 1 + e in [1, 2] by a linear seed + two Newton steps on the packed FMA pipe (rel. error ~1e-5, far below the bf16
// rounding of the result), instead of ex2 + an IEEE divide.  sigmoid(y) = y >= 0 ? 1/(1+e) : e/(1+e), e = exp(-|y|).
__device__ __forceinline__ float2 silu2(float2 y) {
  const float ex = fast_exp2(-1.4426950408889634f * fabsf(y.x));
  const float ey = fast_exp2(-1.4426950408889634f * fabsf(y.y));
  const float2 nd = ffma2(make_float2(ex, ey), make_float2(-1.f, -1.f), make_float2(-1.f, -1.f));   // -(1 + e)
  float2 r = ffma2(nd, make_float2(8.f / 17.f, 8.f / 17.f), make_float2(24.f / 17.f, 24.f / 17.f));
  r = fmul2(r, ffma2(nd, r, make_float2(2.f, 2.f)));
  r = fmul2(r, ffma2(nd, r, make_float2(2.f, 2.f)));
  const float2 num = make_float2(y.x >= 0.f ? 1.f : ex, y.y >= 0.f ? 1.f : ey);
  return fmul2(y, fmul2(num, r));
}

}  // namespace i360
