// Microbenchmark: MUFU-class throughput on B200 (ops / clk / SM) for the candidates of the tanh epilogue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>

template <int OP>
__device__ __forceinline__ float op(float x) {
  float y;
  if constexpr (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if constexpr (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if constexpr (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if constexpr (OP == 3) asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if constexpr (OP == 4) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if constexpr (OP == 5) { uint32_t a = __float_as_uint(x), b; asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(b) : "r"(a)); y = __uint_as_float(b); }
  else if constexpr (OP == 6) { uint32_t a = __float_as_uint(x), b; asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(b) : "r"(a)); y = __uint_as_float(b); }
  else if constexpr (OP == 7) asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if constexpr (OP == 8) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if constexpr (OP == 9) { // FMA-pipe odd polynomial, 5 FFMA + 2 FMNMX + 2 FMUL
    float c = fminf(fmaxf(x, -3.5f), 3.5f); float t = c * c;
    float p = fmaf(t, -1.1e-5f, 4.1e-4f); p = fmaf(p, t, -6.3e-3f); p = fmaf(p, t, 5.2e-2f); p = fmaf(p, t, -0.31f); p = fmaf(p, t, 1.0f);
    y = p * c; }
  else if constexpr (OP == 10) { // ex2-based tanh: 1 - 2/(1+e^{2x}) : EX2 + RCP + 2 FMA
    float e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
    float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f)); y = fmaf(-2.0f, r, 1.0f); }
  else if constexpr (OP == 11) { // packed bf16x2 HFMA2 polynomial (2 elements per op)
    __nv_bfloat162 c = *reinterpret_cast<__nv_bfloat162*>(&x); __nv_bfloat162 t = __hmul2(c, c);
    __nv_bfloat162 k0 = __floats2bfloat162_rn(-6.3e-3f, -6.3e-3f), k1 = __floats2bfloat162_rn(5.2e-2f, 5.2e-2f), k2 = __floats2bfloat162_rn(-0.31f, -0.31f), k3 = __floats2bfloat162_rn(1.f, 1.f);
    __nv_bfloat162 p = __hfma2(t, k0, k1); p = __hfma2(p, t, k2); p = __hfma2(p, t, k3); p = __hmul2(p, c);
    y = *reinterpret_cast<float*>(&p); }
  else if constexpr (OP == 12) { // packed f16x2 HFMA2 polynomial
    __half2 c = *reinterpret_cast<__half2*>(&x); __half2 t = __hmul2(c, c);
    __half2 k0 = __floats2half2_rn(-6.3e-3f, -6.3e-3f), k1 = __floats2half2_rn(5.2e-2f, 5.2e-2f), k2 = __floats2half2_rn(-0.31f, -0.31f), k3 = __floats2half2_rn(1.f, 1.f);
    __half2 p = __hfma2(t, k0, k1); p = __hfma2(p, t, k2); p = __hfma2(p, t, k3); p = __hmul2(p, c);
    y = *reinterpret_cast<float*>(&p); }
  else if constexpr (OP == 13) { // F2FP pack alone: 2 elements per op
    uint32_t b; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b) : "f"(x), "f"(x + 1.0f)); y = __uint_as_float(b | 0x3f000000u); }
  else if constexpr (OP == 14) { // 2 tanh + F2FP pack (the current epilogue), 2 elements
    float a, c; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(a) : "f"(x)); asm volatile("tanh.approx.f32 %0, %1;" : "=f"(c) : "f"(x + 0.5f));
    uint32_t b; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b) : "f"(a), "f"(c)); y = __uint_as_float(b); }
  else if constexpr (OP == 15) { // 2 tanh + integer rounding + PRMT pack, 2 elements
    float a, c; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(a) : "f"(x)); asm volatile("tanh.approx.f32 %0, %1;" : "=f"(c) : "f"(x + 0.5f));
    uint32_t ua = __float_as_uint(a) + 0x8000u, uc = __float_as_uint(c) + 0x8000u;
    y = __uint_as_float(__byte_perm(ua, uc, 0x7632)); }
  else if constexpr (OP == 16) { // I2FP
    y = __uint2float_rn(__float_as_uint(x)) * 1e-9f; }
  else if constexpr (OP == 17) { // FSETP/FSEL style select chain element
    y = (x > 0.5f) ? x * 0.5f : x + 0.25f; }
  return y;
}

template <int OP, int ILP>
__global__ void __launch_bounds__(512) bench(float* out, int iters) {
  float v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = 0.001f * (threadIdx.x + i + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = op<OP>(v[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  if (s == 12345.678f) out[0] = s;
}

template <int OP>
void run(const char* name, int sms, double clk_hz, int elems_per_op) {
  constexpr int ILP = 8;
  const int iters = 4096;
  float* d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<OP, ILP><<<sms, 512>>>(d, 64);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  bench<OP, ILP><<<sms, 512>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = double(sms) * 512 * ILP * iters * elems_per_op;
  double per_clk_sm = ops / (ms * 1e-3) / clk_hz / sms;
  printf("%-28s %8.3f ms  %7.2f elem/clk/SM (at %.0f MHz)\n", name, ms, per_clk_sm, clk_hz / 1e6);
  cudaFree(d);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double clk = khz * 1e3;
  printf("%s, %d SMs, clock attr %.0f MHz\n", p.name, sms, clk / 1e6);
  run<0>("tanh.approx.f32", sms, clk, 1);
  run<1>("ex2.approx.f32", sms, clk, 1);
  run<2>("rcp.approx.f32", sms, clk, 1);
  run<3>("sin.approx.f32", sms, clk, 1);
  run<4>("lg2.approx.f32", sms, clk, 1);
  run<7>("sqrt.approx.f32", sms, clk, 1);
  run<8>("rsqrt.approx.f32", sms, clk, 1);
  run<5>("tanh.approx.bf16x2", sms, clk, 2);
  run<6>("tanh.approx.f16x2", sms, clk, 2);
  run<9>("poly tanh fp32 (9 ops)", sms, clk, 1);
  run<10>("ex2+rcp tanh", sms, clk, 1);
  run<11>("poly bf16x2 HFMA2 (5 ops)", sms, clk, 2);
  run<12>("poly f16x2 HFMA2 (5 ops)", sms, clk, 2);
  run<13>("F2FP pack bf16x2", sms, clk, 2);
  run<14>("2 tanh + F2FP pack", sms, clk, 2);
  run<15>("2 tanh + IADD/PRMT pack", sms, clk, 2);
  run<16>("I2FP u32->f32", sms, clk, 1);
  return 0;
}
