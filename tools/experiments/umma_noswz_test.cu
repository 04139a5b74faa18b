// Micro-test: tcgen05.mma with NO-SWIZZLE K-major operands (core matrices 8 rows x 16 B).
// Layout under test: K core-columns are separate "slots" (slot j holds k = 8j..8j+7 of every row, rows 16 B apart),
// so a K = 16 step spans two slots -- the layout a per-tap 8-channel stem would use (slot = filter tap).
// Descriptor: start address, LBO = byte distance between core matrices adjacent in K, SBO = between 8-row groups.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I scrfd_arcface_facerecognition_b200/csrc -o gpurun_out/umma_noswz tools/experiments/umma_noswz_test.cu
#include "b2f_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_fp16.h>
using namespace b2f;

constexpr int M = 128, N = 64, K = 32;
constexpr int A_SLOT = M * 16, B_SLOT = N * 16;   // bytes per K core-column slot

__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;   // layout type 0 = no swizzle
}

__global__ void __launch_bounds__(128) test_kernel(const __half* A, const __half* B, float* D, int swap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                       // K/8 slots of M rows x 16 B
  uint8_t* sB = smem + (K / 8) * A_SLOT;    // K/8 slots of N rows x 16 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < M * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__half*>(sA + (k / 8) * A_SLOT + r * 16 + (k % 8) * 2) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__half*>(sB + (k / 8) * B_SLOT + r * 16 + (k % 8) * 2) = B[i];
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 64); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = umma_idesc(M, N, 0);
    for (int s = 0; s < K / 16; ++s) {
      const uint32_t a = smem_u32(sA) + s * 2 * A_SLOT, b = smem_u32(sB) + s * 2 * B_SLOT;
      const uint64_t da = swap ? desc_noswz(a, 128, A_SLOT) : desc_noswz(a, A_SLOT, 128);
      const uint64_t db = swap ? desc_noswz(b, 128, B_SLOT) : desc_noswz(b, B_SLOT, 128);
      umma_f16(tmem, da, db, idesc, s ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<__half> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
  srand(1);
  for (int i = 0; i < M * K; ++i) { fA[i] = (rand() % 17 - 8) / 8.f; hA[i] = __float2half(fA[i]); }
  for (int i = 0; i < N * K; ++i) { fB[i] = (rand() % 13 - 6) / 4.f; hB[i] = __float2half(fB[i]); }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += fA[m * K + k] * fB[n * K + k];
      ref[m * N + n] = s;
    }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  const int smem = (K / 8) * (A_SLOT + B_SLOT) + 1024;
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int swap = 0; swap < 2; ++swap) {
    cudaMemset(dD, 0, M * N * 4);
    test_kernel<<<1, 128, smem>>>(dA, dB, dD, swap);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("swap %d: %s\n", swap, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int i = 0; i < M * N; ++i) err = fmax(err, fabs(out[i] - ref[i]));
    printf("LBO/SBO %s: max abs err %.4g  (D[0][0] %.3f ref %.3f, D[9][3] %.3f ref %.3f)\n",
           swap ? "swapped (LBO = 8-row group, SBO = K slot)" : "as documented (LBO = K slot, SBO = 8-row group)", err, out[0],
           ref[0], out[9 * N + 3], ref[9 * N + 3]);
  }
  return 0;
}
