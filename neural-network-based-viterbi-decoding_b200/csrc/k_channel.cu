// K2: fused BPSK + AWGN channel on a counter-based Philox generator.
//
// Replaces modulate_awgn (/root/reference/src/fnn_viterbi.py:224-253: x = float32(c), 0 -> -1,
// y = x + np.random.normal(0, sigma)) and the per-frame Python loops that call it
// (:380-388; fnn_viterbi_mixed_snr.py:452-460; fnn_viterbi_bitwise.py:309-317).  The noise of
// symbol j of global frame f is Philox4x32-10(key = seed, counter = (f, j/4, snr_idx|stream)),
// so a frame is reproducible from (seed, SNR index, frame index) on any GPU count.
//
// One thread = one (frame, 4-symbol block): 1 Philox call, 2 Box–Muller pairs, 4 outputs.
// Consecutive threads write consecutive 4-symbol groups -> coalesced.  Algorithmic HBM traffic
// per frame: ceil(n/8) B read + n*s B written (s = 2 bf16, 4 fp32/int32): 30 B / 58 B at n=14.
#include "common.cuh"
#include "frame_gen.cuh"

namespace nvd {

template <int kDtype>
struct OutT;
template <>
struct OutT<NVD_F32> { using type = float; };
template <>
struct OutT<NVD_BF16> { using type = __nv_bfloat16; };
template <>
struct OutT<NVD_I32_TRUNC> { using type = int32_t; };

template <int kDtype>
__device__ __forceinline__ typename OutT<kDtype>::type cast_out(float v) {
  if constexpr (kDtype == NVD_F32) return v;
  else if constexpr (kDtype == NVD_BF16) return __float2bfloat16_rn(v);
  else return static_cast<int32_t>(v);  // truncation toward zero, as a numpy int32 row stores it
}

template <int kDtype>
__global__ void __launch_bounds__(256) channel_kernel(uint64_t seed, const uint8_t* __restrict__ coded,
                                                      int n, int S, int nblk, int64_t B, float sigma,
                                                      uint32_t stream_word, uint64_t frame0,
                                                      typename OutT<kDtype>::type* __restrict__ y) {
  using T = typename OutT<kDtype>::type;
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= B * nblk) return;
  const int64_t f = t / nblk;
  const int blk = static_cast<int>(t - f * nblk);
  const int j0 = 4 * blk;
  const uint32_t bits = (static_cast<uint32_t>(__ldg(coded + f * S + (j0 >> 3))) >> (j0 & 7)) & 0xFu;
  float z[4];
  noise4(seed, frame0 + f, blk, stream_word, z);
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = fmaf(sigma, z[k], ((bits >> k) & 1u) ? 1.0f : -1.0f);
  T* row = y + f * n + j0;
  if (j0 + 4 <= n && (n & 1) == 0) {
    // even n: rows start 4 B (bf16) / 8 B (fp32, int32) aligned -> two vector stores
    if constexpr (kDtype == NVD_BF16) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]);
      const __nv_bfloat162 hi = __floats2bfloat162_rn(v[2], v[3]);
      reinterpret_cast<__nv_bfloat162*>(row)[0] = lo;
      reinterpret_cast<__nv_bfloat162*>(row)[1] = hi;
    } else if constexpr (kDtype == NVD_F32) {
      reinterpret_cast<float2*>(row)[0] = make_float2(v[0], v[1]);
      reinterpret_cast<float2*>(row)[1] = make_float2(v[2], v[3]);
    } else {
      reinterpret_cast<int2*>(row)[0] = make_int2(cast_out<kDtype>(v[0]), cast_out<kDtype>(v[1]));
      reinterpret_cast<int2*>(row)[1] = make_int2(cast_out<kDtype>(v[2]), cast_out<kDtype>(v[3]));
    }
  } else {
    for (int k = 0; k < 4 && j0 + k < n; ++k) row[k] = cast_out<kDtype>(v[k]);
  }
}

int launch_channel(nvd_ctx* ctx, const uint8_t* coded, int n, int64_t B, float sigma, int snr_idx,
                   uint64_t frame0, int out_dtype, void* y) {
  if (B == 0) return NVD_OK;
  const int S = (n + 7) / 8, nblk = (n + 3) / 4;
  const int64_t threads = B * nblk;
  const unsigned grid = static_cast<unsigned>((threads + 255) / 256);
  const uint32_t sw = noise_stream_word(snr_idx);
  switch (out_dtype) {
    case NVD_F32:
      channel_kernel<NVD_F32><<<grid, 256, 0, ctx->stream>>>(ctx->seed, coded, n, S, nblk, B, sigma,
                                                             sw, frame0, static_cast<float*>(y));
      break;
    case NVD_BF16:
      channel_kernel<NVD_BF16><<<grid, 256, 0, ctx->stream>>>(
          ctx->seed, coded, n, S, nblk, B, sigma, sw, frame0, static_cast<__nv_bfloat16*>(y));
      break;
    case NVD_I32_TRUNC:
      channel_kernel<NVD_I32_TRUNC><<<grid, 256, 0, ctx->stream>>>(
          ctx->seed, coded, n, S, nblk, B, sigma, sw, frame0, static_cast<int32_t*>(y));
      break;
    default:
      set_error("nvd_channel: unknown out_dtype %d", out_dtype);
      return NVD_ERR_INVALID;
  }
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

}  // namespace nvd
