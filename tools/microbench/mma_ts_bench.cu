// Microbenchmark: tcgen05.mma execution rate, SS mode (A from shared memory) vs TS mode (A from tensor memory),
// long chains from one issuing thread, N = 64 / 128 / 256.  clk per MMA vs the ideal 128 * N * 16 / 4096.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../neural-network-based-viterbi-decoding_b200/csrc -o mma_ts_bench mma_ts_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "ptx.cuh"
using namespace nvd;

__global__ void __launch_bounds__(128, 1) chain(int N, int count, int ts, int same_b, long long* out, const uint8_t* gsrc, int tma_kb_per_mma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t tbar[4];
  __shared__ volatile int stop;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  const uint32_t sbase = ptx::smem_u32(smem);
  if (tid == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); for (int i = 0; i < 4; ++i) ptx::mbar_init(ptx::smem_u32(&tbar[i]), 1); stop = 0; ptx::mbar_fence_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish(); }
  for (int i = tid; i < 196608 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  ptx::fence_proxy_async_smem(); ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = __shfl_sync(0xFFFFFFFFu, slot, 0);
  const uint32_t idesc = ptx::make_idesc_bf16_f32(128, N);
  if (warp == 0) {
    long long t0 = 0, t2 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      __syncwarp();
      t0 = clock64();
      if (ptx::elect_one()) {
        for (int j = 0; j < count; ++j) {
          const int kk = j & 15;                                           // 16 K-steps of A (256 TMEM columns / 64 KB)
          const uint64_t bd = ptx::make_smem_desc(sbase + 65536 + (same_b ? 0 : (j & 7) * N * 32), N * 16, 128);
          if (ts) ptx::umma_f16_ts(tm, tm + 256 + 8 * kk, bd, idesc, j > 0);
          else ptx::umma_f16_ss(tm, ptx::make_smem_desc(sbase + kk * 4096, 2048, 128), bd, idesc, j > 0);
        }
        ptx::umma_commit(ptx::smem_u32(&bar));
      }
      __syncwarp();
      while (!ptx::mbar_try_wait(ptx::smem_u32(&bar), rep & 1)) {}
      t2 = clock64();
    }
    if (tid == 0) out[0] = t2 - t0;
  }
  if (warp == 0 && tid == 0) stop = 1;
  if (warp == 1 && tma_kb_per_mma > 0 && ptx::elect_one()) {
    // background weight stream: 16 KB bulk copies into a 4-slot ring at 128 KB.. while the chain runs
    uint32_t it = 0;
    while (!stop) {
      const int s = it & 3;
      if (it >= 4) while (!ptx::mbar_try_wait(ptx::smem_u32(&tbar[s]), ((it >> 2) - 1) & 1)) {}
      ptx::mbar_arrive_expect_tx(ptx::smem_u32(&tbar[s]), 16384);
      ptx::bulk_g2s(sbase + 131072 + s * 16384, gsrc + (it & 63) * 16384, 16384, ptx::smem_u32(&tbar[s]));
      ++it;
    }
    for (uint32_t k = (it > 4 ? it - 4 : 0); k < it; ++k) while (!ptx::mbar_try_wait(ptx::smem_u32(&tbar[k & 3]), (k >> 2) & 1)) {}
    out[1] = it;
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16); long long h[2];
  uint8_t* g; cudaMalloc(&g, 1 << 20); cudaMemset(g, 0, 1 << 20);
  cudaFuncSetAttribute(chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
  for (int ts : {0, 1})
    for (int N : {64, 128, 256})
      for (int tma : {0, 1}) {
        const int count = 2048, same_b = 0;
        cudaMemset(d, 0, 16);
        chain<<<1, 128, 196608>>>(N, count, ts, same_b, d, g, tma);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%s N=%3d %s: %6.1f clk/MMA (ideal %3d)  background copies: %lld x 16 KB = %.1f B/clk  %s\n", ts ? "TS" : "SS", N,
               tma ? "with bulk-copy stream" : "alone                ", double(h[0]) / (3.0 * count) * 3.0, 128 * N * 16 / 4096, h[1],
               h[1] * 16384.0 / (double(h[0]) * 3), e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
