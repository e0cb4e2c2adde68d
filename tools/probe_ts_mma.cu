// Probe: tcgen05.mma with the A operand in TMEM (".ts" form).  Establishes the TMEM layout of a bf16 A tile:
// hypothesis H1 -- lane = row, 32-bit column c holds A[row][2c] (low half) and A[row][2c+1] (high half), a K=16
// instruction step advances the A address by 8 columns.  B = identity (K-major, 128B swizzle) so D == A as the MMA reads it.
#include "../imagine360_b200/csrc/common.cuh"
#include <stdio.h>
#include <vector>
using namespace i360;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
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

// A [128][64] bf16 (global), B = identity [64][64]; out D [128][64] fp32.  a_col0: TMEM column where A starts.
__global__ void __launch_bounds__(128, 1) probe(const bf16* A, float* D, int a_col0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* sB = smem;                      // [64 rows n][64 k] bf16, 128-byte rows, 128B swizzle
  const int tid = threadIdx.x, warp = tid >> 5;
  // identity in swizzled layout: element (n, k) at n*128 + ((k/8 ^ (n&7))*16) + (k%8)*2
  for (int i = tid; i < 64 * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    *reinterpret_cast<bf16*>(sB + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
  // my row of A -> TMEM columns a_col0 .. a_col0+31 (H1 packing)
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) {
    const uint32_t lo = *reinterpret_cast<const unsigned short*>(A + tid * 64 + 2 * c);
    const uint32_t hi = *reinterpret_cast<const unsigned short*>(A + tid * 64 + 2 * c + 1);
    v[c] = lo | (hi << 16);
  }
  tmem_st_x32(tb + lane_sel + a_col0, v);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    for (int ks = 0; ks < 4; ++ks)
      umma_bf16_ts(tb + 128, tb + a_col0 + ks * 8, make_smem_desc(smem_u32(sB) + ks * 32, 1024, 16, SWZ_128B), idesc, ks != 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t o[32];
  for (int h = 0; h < 2; ++h) {
    tmem_ld_x32(tb + lane_sel + 128 + h * 32, o);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) D[tid * 64 + h * 32 + c] = __uint_as_float(o[c]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 256); }
}

int main() {
  std::vector<bf16> hA(128 * 64);
  for (int r = 0; r < 128; ++r) for (int k = 0; k < 64; ++k) hA[r * 64 + k] = __float2bfloat16((float)((r % 4) * 64 + k));   // exact in bf16
  bf16* dA; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  for (int a_col0 : {0, 64}) {
    cudaMemset(dD, 0, 128 * 64 * 4);
    probe<<<1, 128, 16384>>>(dA, dD, a_col0);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> hD(128 * 64);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 128 * 64; ++i) if (hD[i] != __bfloat162float(hA[i])) ++bad;
    printf("a_col0=%d  cuda=%s  mismatches=%d / 8192\n", a_col0, cudaGetErrorString(e), bad);
    if (bad) for (int r : {0, 1, 33, 127}) { printf("row %d:", r); for (int k = 0; k < 64; ++k) printf(" %g", hD[r * 64 + k]); printf("\n"); }
  }
  return 0;
}
