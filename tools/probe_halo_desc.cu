// Probe: can one tcgen05.mma read its K-major, 128B-swizzled A operand from a HALO tile -- 8-pixel image rows that are
// (TW + 2) = 10 smem rows apart (SBO = 1280 B, not a multiple of the 1024 B swizzle atom) and whose start is shifted by
// (kh * 10 + kw) rows for tap (kh, kw)?  Then the nine taps of a 3x3 convolution can share ONE loaded tile per channel
// block instead of nine.  Variants: descriptor "base offset" field (bits 49-51) = 0, or (start_address >> 7) & 7.
// B = identity (N = 64), so D[m][k] is the A element the tensor core actually read for tile row m.
#include "../imagine360_b200/csrc/common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <vector>
using namespace i360;

constexpr int TW = 8, TH = 16, HW = TW + 2, HH = TH + 2, HR = HW * HH;   // 180 halo rows of 128 bytes

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t sbo, int base_off_mode) {
  uint64_t d = make_smem_desc(addr, sbo, 16, SWZ_128B);
  if (base_off_mode == 1) d |= static_cast<uint64_t>((addr >> 7) & 7) << 49;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe(const bf16* X, float* D, int kh, int kw, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* sA = smem;                      // halo tile: HR rows x 64 ch, swizzled by ABSOLUTE row index (what TMA writes)
  uint8_t* sB = smem + 24 * 1024;          // identity [64 n][64 k]
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < HR * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<bf16*>(sA + r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2) = X[i];
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    *reinterpret_cast<bf16*>(sB + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 64); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (tid == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    const uint32_t a0 = smem_u32(sA) + (kh * HW + kw) * 128;
    for (int ks = 0; ks < 4; ++ks)
      umma_bf16_ss(tb, desc(a0 + ks * 32, HW * 128, mode), make_smem_desc(smem_u32(sB) + ks * 32, 1024, 16, SWZ_128B), idesc, ks != 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
  uint32_t o[32];
  for (int h = 0; h < 2; ++h) {
    tmem_ld_x32(tb + lane_sel + h * 32, o);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) D[tid * 64 + h * 32 + c] = __uint_as_float(o[c]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 64); }
}

int main() {
  std::vector<bf16> hX(HR * 64);
  for (int r = 0; r < HR; ++r) for (int k = 0; k < 64; ++k) hX[r * 64 + k] = __float2bfloat16((float)(r + (k % 2 ? 0.5f : 0.f) + 256 * (k / 8 % 2)));
  // pass 0: value = halo row r (+256 for odd 16-byte chunks); pass 1: value = (r % 16) * 8 + chunk index (all exact in
  // bf16): together they pin the row AND all three swizzle bits of what the tensor core read
  const int pass = getenv("PROBE_PASS") ? atoi(getenv("PROBE_PASS")) : 0;
  for (int r = 0; r < HR; ++r) for (int k = 0; k < 64; ++k)
    hX[r * 64 + k] = __float2bfloat16(pass == 0 ? (float)(r + ((k >> 3) & 1) * 256) : (float)((r % 16) * 8 + (k >> 3)));
  bf16* dX; float* dD;
  cudaMalloc(&dX, hX.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 24 * 1024 + 8 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 2; ++mode)
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        cudaMemset(dD, 0, 128 * 64 * 4);
        probe<<<1, 128, smem>>>(dX, dD, kh, kw, mode);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> hD(128 * 64);
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int m = 0; m < 128; ++m)
          for (int k = 0; k < 64; ++k) {
            const int r = (m / TW + kh) * HW + (m % TW) + kw;
            if (hD[m * 64 + k] != __bfloat162float(hX[r * 64 + k])) { if (first < 0) first = m * 64 + k; ++bad; }
          }
        printf("mode=%d tap=(%d,%d) cuda=%s mismatches=%d/8192", mode, kh, kw, cudaGetErrorString(e), bad);
        if (bad) { const int m = first / 64; printf("  first at m=%d k=%d got %g want %g | row m: ", m, first % 64, hD[first], __bfloat162float(hX[((m / TW + kh) * HW + (m % TW) + kw) * 64 + first % 64]));
                   for (int k = 0; k < 64; k += 8) printf("%g ", hD[m * 64 + k]); }
        printf("\n");
        if (e != cudaSuccess) return 1;
      }
  return 0;
}
