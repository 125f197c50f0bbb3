// K3 (fp32 parity path): generic dense layer  out = act(in @ W + b)  on the FP32 FMA pipe, and
// the first-index argmax hard decision.
//
// Replaces, for ANY layer widths, the TensorFlow ops behind multilayer_perceptron
// (/root/reference/src/fnn_viterbi.py:495-508), the bitwise net (fnn_viterbi_bitwise.py:388,397),
// the sliding-window net (fnn_viterbi_seq_predict.py:408-412) and tf.argmax (fnn_viterbi.py:606-610).
// This is the path held to 1e-5 against the oracle (no tensor-core mode has an fp32 mantissa, and
// tanh.approx is ~2^-11): FFMA accumulation in ascending k with per-16 partial sums, libdevice
// tanhf.  The throughput path is the tcgen05 kernel in k_mlp_tc.cu.
#include <algorithm>

#include "common.cuh"

namespace nvd {

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<int32_t>(int32_t v) { return static_cast<float>(v); }

template <typename InT>
__global__ void __launch_bounds__(256) dense_act_kernel(const InT* __restrict__ A, const float* __restrict__ W,
                                                        const float* __restrict__ bias, float* __restrict__ out,
                                                        int64_t B, int K, int N, int act) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * BM;
  const int col0 = blockIdx.y * BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + 256 * q;
      const int r = idx >> 4, kk = idx & 15;
      const int64_t gr = row0 + r;
      const int gk = k0 + kk;
      As[kk][r] = (gr < B && gk < K) ? to_f32<InT>(A[gr * K + gk]) : 0.0f;
      const int kb = idx >> 6, c = idx & 63;
      const int gkb = k0 + kb, gc = col0 + c;
      Bs[kb][c] = (gkb < K && gc < N) ? W[static_cast<int64_t>(gkb) * N + gc] : 0.0f;
    }
    __syncthreads();
    float part[4][4] = {};
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gr = row0 + ty * 4 + i;
    if (gr >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = col0 + tx * 4 + j;
      if (gc >= N) continue;
      float v = acc[i][j] + bias[gc];
      if (act == NVD_ACT_TANH) v = tanhf(v);
      else if (act == NVD_ACT_RELU) v = fmaxf(v, 0.0f);
      out[gr * N + gc] = v;
    }
  }
}

// one warp per row; first index on ties (tf.argmax / np.argmax semantics)
template <typename HardT>
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ x, int64_t B, int N,
                                                          HardT* __restrict__ hard) {
  const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* r = x + row * N;
  float best = -INFINITY;
  int bi = 0x7FFFFFFF;
  for (int c = lane; c < N; c += 32) {
    const float v = r[c];
    if (v > best) { best = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xFFFFFFFFu, best, o);
    const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (bi == 0x7FFFFFFF) bi = 0;  // all -inf / NaN row
  if (lane == 0) hard[row] = static_cast<HardT>(bi);
}

template <typename InT>
int run_layer(nvd_ctx* ctx, const InT* in, const float* W, const float* b, float* out, int64_t B, int K,
              int N, int act) {
  dim3 grid(static_cast<unsigned>((B + BM - 1) / BM), static_cast<unsigned>((N + BN - 1) / BN));
  dense_act_kernel<InT><<<grid, 256, 0, ctx->stream>>>(in, W, b, out, B, K, N, act);
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

}  // namespace

int launch_mlp_fp32(nvd_ctx* ctx, const MlpModel& m, const void* y, int y_dtype, int64_t B, float* soft_out,
                    void* hard_out, int hard_bytes) {
  if (B == 0) return NVD_OK;
  int maxdim = 0;
  for (int l = 1; l <= m.n_layers; ++l) maxdim = std::max(maxdim, m.dims[l]);
  // chunk rows so the two ping-pong activation buffers stay <= 512 MiB
  int64_t chunk = (static_cast<int64_t>(512) << 20) / (2LL * maxdim * 4);
  chunk = std::max<int64_t>(1024, std::min<int64_t>(chunk, B));
  chunk = (chunk + BM - 1) / BM * BM;
  NVD_TRY(ensure_scratch(ctx, static_cast<size_t>(2) * chunk * maxdim * 4));
  float* buf[2] = {static_cast<float*>(ctx->scratch), static_cast<float*>(ctx->scratch) + chunk * maxdim};
  const size_t in_elem = (y_dtype == NVD_BF16) ? 2 : 4;
  for (int64_t r0 = 0; r0 < B; r0 += chunk) {
    const int64_t rows = std::min<int64_t>(chunk, B - r0);
    const char* yin = static_cast<const char*>(y) + static_cast<size_t>(r0) * m.dims[0] * in_elem;
    const float* cur = nullptr;
    for (int l = 0; l < m.n_layers; ++l) {
      const bool last = (l == m.n_layers - 1);
      float* dst = (last && soft_out) ? soft_out + r0 * m.dims[m.n_layers] : buf[l & 1];
      if (l == 0) {
        if (y_dtype == NVD_F32)
          NVD_TRY(run_layer(ctx, reinterpret_cast<const float*>(yin), m.W[0], m.b[0], dst, rows, m.dims[0], m.dims[1], m.acts[0]));
        else if (y_dtype == NVD_BF16)
          NVD_TRY(run_layer(ctx, reinterpret_cast<const __nv_bfloat16*>(yin), m.W[0], m.b[0], dst, rows, m.dims[0], m.dims[1], m.acts[0]));
        else
          NVD_TRY(run_layer(ctx, reinterpret_cast<const int32_t*>(yin), m.W[0], m.b[0], dst, rows, m.dims[0], m.dims[1], m.acts[0]));
      } else {
        NVD_TRY(run_layer(ctx, cur, m.W[l], m.b[l], dst, rows, m.dims[l], m.dims[l + 1], m.acts[l]));
      }
      cur = dst;
    }
    if (hard_out) {
      const int N = m.dims[m.n_layers];
      const unsigned grid = static_cast<unsigned>((rows * 32 + 255) / 256);
      if (hard_bytes == 1)
        argmax_rows_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(cur, rows, N, static_cast<uint8_t*>(hard_out) + r0);
      else if (hard_bytes == 2)
        argmax_rows_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(cur, rows, N, static_cast<uint16_t*>(hard_out) + r0);
      else
        argmax_rows_kernel<int32_t><<<grid, 256, 0, ctx->stream>>>(cur, rows, N, static_cast<int32_t*>(hard_out) + r0);
      NVD_CHECK_LAUNCH();
    }
  }
  return NVD_OK;
}

}  // namespace nvd
