// Microbenchmark: which SM pipes the candidate epilogue instructions use on B200 and how they overlap.
// For each pair (A, B): time A alone, B alone and A+B interleaved in one warp; t(A+B) ~ max => different pipes,
// ~ sum => shared pipe / issue bound.  Also the hybrid tanh epilogue (MUFU for a fraction of the columns, an
// fp32x2 FMA-pipe polynomial for the rest) end to end: tcgen05.ld -> tanh -> bf16 pack -> tcgen05.st.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../neural-network-based-viterbi-decoding_b200/csrc -o pipe_bench pipe_bench.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include "ptx.cuh"
using namespace nvd;

__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

enum { OP_TANH, OP_FFMA, OP_FFMA2, OP_LOP3, OP_VIMNMX3, OP_FMNMX, OP_PRMT, OP_IMADW, OP_F2FP, OP_I2FP, OP_HFMA2, OP_NONE };

// state: two 32-bit registers per chain
template <int OP>
__device__ __forceinline__ void step(uint32_t& a, uint32_t& b, uint32_t k0, uint32_t k1) {
  if constexpr (OP == OP_TANH) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(a))); a = __float_as_uint(y); }
  else if constexpr (OP == OP_FFMA) { float y; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(__uint_as_float(a)), "f"(__uint_as_float(k0)), "f"(__uint_as_float(k1))); a = __float_as_uint(y); }
  else if constexpr (OP == OP_FFMA2) { uint64_t v = (uint64_t(b) << 32) | a, k = (uint64_t(k1) << 32) | k0; v = fma2(v, k, k); a = uint32_t(v); b = uint32_t(v >> 32); }
  else if constexpr (OP == OP_LOP3) { asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(a) : "r"(a), "r"(k0), "r"(k1)); }
  else if constexpr (OP == OP_VIMNMX3) { a = uint32_t(max(max(int(a), int(k0) + int(b)), int(b))); }
  else if constexpr (OP == OP_FMNMX) { float y; asm volatile("{.reg .f32 t; abs.f32 t, %1; min.f32 %0, t, %2;}" : "=f"(y) : "f"(__uint_as_float(a)), "f"(__uint_as_float(k0))); a = __float_as_uint(y); }
  else if constexpr (OP == OP_PRMT) { asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(a) : "r"(a), "r"(k0)); }
  else if constexpr (OP == OP_IMADW) { uint64_t y; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(y) : "r"(a), "r"(k0)); a = uint32_t(y) ^ uint32_t(y >> 32); }
  else if constexpr (OP == OP_F2FP) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(a) : "f"(__uint_as_float(a)), "f"(__uint_as_float(k0))); }
  else if constexpr (OP == OP_I2FP) { float y; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(y) : "r"(a)); a = __float_as_uint(y); }
  else if constexpr (OP == OP_HFMA2) { asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(a) : "r"(a), "r"(k0), "r"(k1)); }
}

template <int A, int NA, int B, int NB>
__global__ void __launch_bounds__(512) pair_bench(uint32_t* out, int iters, uint32_t k0, uint32_t k1) {
  constexpr int ILP = 6;
  uint32_t xa[ILP], xb[ILP], ya[ILP], yb[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { xa[i] = 0x3c000000u + threadIdx.x * 77 + i; xb[i] = xa[i] ^ 0x1234u; ya[i] = xa[i] + 99; yb[i] = xb[i] + 7; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
#pragma unroll
      for (int n = 0; n < NA; ++n) step<A>(xa[i], xb[i], k0, k1);
#pragma unroll
      for (int n = 0; n < NB; ++n) step<B>(ya[i], yb[i], k0, k1);
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s ^= xa[i] ^ xb[i] ^ ya[i] ^ yb[i];
  if (s == 0x12345678u) out[0] = s;
}

template <int A, int NA, int B, int NB>
float run_pair(int sms) {
  uint32_t* d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2048;
  pair_bench<A, NA, B, NB><<<sms, 512>>>(d, 16, 0x3f7f0000u, 0x3a000000u); cudaDeviceSynchronize();
  cudaEventRecord(e0); pair_bench<A, NA, B, NB><<<sms, 512>>>(d, iters, 0x3f7f0000u, 0x3a000000u); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); cudaFree(d);
  // clk per (warp-level) loop body iteration element per SMSP: 4 warps per SMSP, ILP 6
  return ms * 1e-3f * 1.965e9f / (iters * 6.0f * 4.0f);
}
#define PAIR(A, NA, B, NB) do { float ta = run_pair<A, NA, OP_NONE, 0>(sms), tb = run_pair<OP_NONE, 0, B, NB>(sms), tab = run_pair<A, NA, B, NB>(sms); \
  printf("%-10s x%d + %-10s x%d : A %6.2f  B %6.2f  A+B %6.2f clk/SMSP per (warp x body)   [sum %6.2f max %6.2f]\n", #A, NA, #B, NB, ta, tb, tab, ta + tb, ta > tb ? ta : tb); } while (0)

// ---- hybrid epilogue -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack2(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
struct PolyK { uint64_t k[6]; };
// tanh(x) ~ sign(x) * u * Q(u), u = min(|x|, 3.5), Q of degree 5 (max abs error 1.1e-3)
__device__ __forceinline__ uint32_t poly_pair(uint32_t xa, uint32_t xb, const PolyK& K) {
  float ua, ub;
  asm("{.reg .f32 t; abs.f32 t, %1; min.f32 %0, t, 0f40600000;}" : "=f"(ua) : "f"(__uint_as_float(xa)));
  asm("{.reg .f32 t; abs.f32 t, %1; min.f32 %0, t, 0f40600000;}" : "=f"(ub) : "f"(__uint_as_float(xb)));
  const uint64_t u = (uint64_t(__float_as_uint(ub)) << 32) | __float_as_uint(ua);
  uint64_t p = fma2(u, K.k[5], K.k[4]);
  p = fma2(p, u, K.k[3]); p = fma2(p, u, K.k[2]); p = fma2(p, u, K.k[1]); p = fma2(p, u, K.k[0]);
  p = mul2(p, u);
  uint32_t pa = uint32_t(p), pb = uint32_t(p >> 32);
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(pa) : "r"(xa), "r"(pa));   // (x & sign) | p
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(pb) : "r"(xb), "r"(pb));
  return pack2(__uint_as_float(pa), __uint_as_float(pb));
}
// NPOLY of every 8 pairs go to the FMA pipe; STORE 0: tcgen05.st, 1: st.shared.v4, 2: none
template <int NPOLY, int STORE>
__global__ void __launch_bounds__(512, 1) epi(int iters, float* out, PolyK K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, r = tid & 127;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0), g = warp >> 2;
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = slot + g * 128 + (uint32_t((warp & 3) * 32) << 16);
  const uint32_t a_next = ptx::smem_u32(smem) + g * 32768;
  uint32_t accu = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 128; c += 32) {
      uint32_t v[32];
      ptx::tmem_ld32(tm + c, v); ptx::tmem_wait_ld();
      uint32_t w[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if ((e & 7) < NPOLY) w[e] = poly_pair(v[2 * e], v[2 * e + 1], K);
        else w[e] = pack2(ptx::tanh_approx(__uint_as_float(v[2 * e])), ptx::tanh_approx(__uint_as_float(v[2 * e + 1])));
      }
      if (STORE == 0) { ptx::tmem_st16(tm + (c >> 1), w); }
      else if (STORE == 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_next + ((c >> 3) + q) * 2048 + r * 16), "r"(w[4 * q]), "r"(w[4 * q + 1]), "r"(w[4 * q + 2]), "r"(w[4 * q + 3]) : "memory");
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) accu ^= w[e];
      }
    }
    if (STORE == 0) ptx::tmem_wait_st();
  }
  if (accu == 0x12345u) out[0] = 1.f;
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(slot, 512);
}
template <int NPOLY, int STORE>
void run_epi(const char* name, int sms, const PolyK& K) {
  float* d; cudaMalloc(&d, 4);
  cudaFuncSetAttribute(epi<NPOLY, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  epi<NPOLY, STORE><<<sms, 512, 4 * 32768>>>(10, d, K); cudaDeviceSynchronize();
  const int iters = 2000;
  cudaEventRecord(e0); epi<NPOLY, STORE><<<sms, 512, 4 * 32768>>>(iters, d, K); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-44s %8.3f ms  %6.2f tanh/clk/SM   %s\n", name, ms, 512.0 * 128 * iters / (ms * 1e-3) / 1.965e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs; per-op figures assume 1965 MHz\n", p.name, sms);
  PAIR(OP_FFMA, 4, OP_FFMA2, 4);
  PAIR(OP_LOP3, 4, OP_VIMNMX3, 4);
  PAIR(OP_LOP3, 4, OP_FMNMX, 4);
  PAIR(OP_LOP3, 4, OP_PRMT, 4);
  PAIR(OP_LOP3, 4, OP_F2FP, 4);
  PAIR(OP_FFMA, 4, OP_F2FP, 4);
  PAIR(OP_TANH, 2, OP_F2FP, 2);
  PAIR(OP_FFMA2, 4, OP_IMADW, 4);
  PAIR(OP_FFMA, 4, OP_IMADW, 4);
  PAIR(OP_LOP3, 4, OP_IMADW, 4);
  PAIR(OP_TANH, 1, OP_FFMA2, 3);
  PAIR(OP_TANH, 1, OP_FFMA2, 4);
  PAIR(OP_TANH, 1, OP_LOP3, 4);
  PAIR(OP_TANH, 1, OP_I2FP, 2);
  PAIR(OP_LOP3, 4, OP_I2FP, 2);
  PAIR(OP_FFMA, 4, OP_I2FP, 2);
  PAIR(OP_FFMA2, 4, OP_HFMA2, 4);
  PAIR(OP_FFMA2, 4, OP_LOP3, 4);
  PolyK K;
  const float q[6] = {1.0f, -0.01f, -0.3f, 0.15f, -0.03f, 0.002f};   // placeholder magnitudes (timing only)
  for (int i = 0; i < 6; ++i) { uint32_t u; memcpy(&u, &q[i], 4); K.k[i] = (uint64_t(u) << 32) | u; }
  run_epi<0, 0>("ldtm + 8/8 MUFU + pack + sttm", sms, K);
  run_epi<0, 1>("ldtm + 8/8 MUFU + pack + sts", sms, K);
  run_epi<0, 2>("ldtm + 8/8 MUFU + pack", sms, K);
  run_epi<2, 0>("ldtm + 6/8 MUFU + 2/8 poly + sttm", sms, K);
  run_epi<3, 0>("ldtm + 5/8 MUFU + 3/8 poly + sttm", sms, K);
  run_epi<4, 0>("ldtm + 4/8 MUFU + 4/8 poly + sttm", sms, K);
  run_epi<5, 0>("ldtm + 3/8 MUFU + 5/8 poly + sttm", sms, K);
  run_epi<8, 0>("ldtm + 8/8 poly + sttm", sms, K);
  run_epi<3, 1>("ldtm + 5/8 MUFU + 3/8 poly + sts", sms, K);
  run_epi<4, 2>("ldtm + 4/8 MUFU + 4/8 poly (no store)", sms, K);
  return 0;
}
