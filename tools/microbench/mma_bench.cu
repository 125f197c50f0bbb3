// Microbenchmark: tcgen05.mma issue cost and execution rate for the small shapes of the FNN decoder.
// One CTA, one issuing thread; reports clocks from first issue to (a) last issue returned, (b) commit observed.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "ptx.cuh"
using namespace nvd;

__global__ void __launch_bounds__(128, 1) mma_issue(int N, int count, int groups, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bars[4];
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  const uint32_t sbase = ptx::smem_u32(smem);
  if (tid == 0) { for (int i = 0; i < 4; ++i) ptx::mbar_init(ptx::smem_u32(&bars[i]), 1); ptx::mbar_fence_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem(); ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = __shfl_sync(0xFFFFFFFFu, slot, 0);
  const uint32_t idesc = ptx::make_idesc_bf16_f32(128, N);
  // each of `groups` warps issues `count` MMAs into its own 128-column accumulator (like the decoder's groups)
  const uint64_t a_desc0 = ptx::make_smem_desc(sbase + warp * 16384, 2048, 128);
  const uint64_t b_desc0 = ptx::make_smem_desc(sbase + 65536, N * 16, 128);
  if (warp < groups) {
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < 4; ++rep) {
      __syncwarp();
      t0 = clock64();
      if (ptx::elect_one()) {
        uint64_t ad = a_desc0, bd = b_desc0;
        for (int j = 0; j < count; ++j) {
          ptx::umma_f16_ss(tm + warp * 128, ad, bd, idesc, j > 0);
          ad += 4096 >> 4; bd += 4096 >> 4;
          if ((j & 3) == 3) { ad = a_desc0; bd = b_desc0; }
        }
        ptx::umma_commit(ptx::smem_u32(&bars[warp]));
      }
      __syncwarp();
      t1 = clock64();
      while (!ptx::mbar_try_wait(ptx::smem_u32(&bars[warp]), rep & 1)) {}
      t2 = clock64();
    }
    if ((tid & 31) == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0; }
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64); long long h[8];
  cudaFuncSetAttribute(mma_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
  for (int groups : {1, 4})
    for (int N : {32, 64, 128})
      for (int count : {1, 2, 5, 9, 18}) {
        cudaMemset(d, 0, 64);
        mma_issue<<<1, 128, 98304>>>(N, count, groups, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("groups=%d N=%3d count=%2d : issue %5lld clk  done %5lld clk  (%.1f clk/MMA, ideal %d)  [g3: %lld/%lld] %s\n", groups, N, count,
               h[0], h[1], double(h[1]) / count, 128 * N * 16 / 4096, h[6], h[7], e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
