// K2: fused BPSK + AWGN channel on a counter-based Philox generator.
//
// Replaces modulate_awgn (/root/reference/src/fnn_viterbi.py:224-253: x = float32(c), 0 -> -1,
// y = x + np.random.normal(0, sigma)) and the per-frame Python loops that call it
// (:380-388; fnn_viterbi_mixed_snr.py:452-460; fnn_viterbi_bitwise.py:309-317).  The noise of
// symbol j of global frame f is Philox4x32-10(key = seed, counter = (f, j/8, snr_idx|stream)) — eight symbols per
// Philox call (frame_gen.cuh) — so a frame is reproducible from (seed, SNR index, frame index) on any GPU count.
//
// One thread = one (frame, 8-symbol block): 1 Philox call, 4 Box–Muller pairs, 8 outputs, one coded byte in.
// Consecutive threads write consecutive 8-symbol groups -> coalesced.  Algorithmic HBM traffic
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
template <>
struct OutT<NVD_BF16_TRUNC> { using type = __nv_bfloat16; };

template <int kDtype>
__device__ __forceinline__ typename OutT<kDtype>::type cast_out(float v) {
  if constexpr (kDtype == NVD_F32) return v;
  else if constexpr (kDtype == NVD_BF16) return __float2bfloat16_rn(v);
  else if constexpr (kDtype == NVD_BF16_TRUNC) return __float2bfloat16_rn(truncf(v));   // the same integer, exact in bf16
  else return static_cast<int32_t>(v);  // truncation toward zero, as a numpy int32 row stores it
}

// kSmall: B * nblk < 2^32, so the (frame, block) split is a 32-bit division instead of a 64-bit one (which costs
// more instructions than the Philox rounds).  The round keys arrive precomputed (PhiloxKeys, frame_gen.cuh).
template <int kDtype, bool kSmall>
__global__ void __launch_bounds__(256) channel_kernel(const PhiloxKeys rk, const uint8_t* __restrict__ coded,
                                                      int n, int S, int nblk, int64_t B, float sigma,
                                                      uint32_t stream_word, uint64_t frame0, int64_t row_stride,
                                                      typename OutT<kDtype>::type* __restrict__ y) {
  using T = typename OutT<kDtype>::type;
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= B * nblk) return;
  int64_t f;
  if (kSmall) f = static_cast<uint32_t>(t) / static_cast<uint32_t>(nblk);
  else f = t / nblk;
  const int blk = static_cast<int>(t - f * nblk);
  const int j0 = kNormalsPerBlock * blk;
  const uint32_t bits = __ldg(coded + f * S + blk);        // the block's eight code bits are byte `blk` of the packed row
  float v[8];
  block_symbols<4>(rk, frame0 + f, blk, stream_word, bm_scale(sigma), bits, v);
  T* row = y + f * row_stride + j0;
  if (j0 + 8 <= n && (n & 1) == 0 && (row_stride & 1) == 0) {
    // even n: rows start 4 B (bf16) / 8 B (fp32, int32) aligned -> four vector stores
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      if constexpr (kDtype == NVD_BF16) reinterpret_cast<__nv_bfloat162*>(row + k)[0] = __floats2bfloat162_rn(v[k], v[k + 1]);
      else if constexpr (kDtype == NVD_BF16_TRUNC) reinterpret_cast<__nv_bfloat162*>(row + k)[0] = __floats2bfloat162_rn(truncf(v[k]), truncf(v[k + 1]));
      else if constexpr (kDtype == NVD_F32) reinterpret_cast<float2*>(row + k)[0] = make_float2(v[k], v[k + 1]);
      else reinterpret_cast<int2*>(row + k)[0] = make_int2(cast_out<kDtype>(v[k]), cast_out<kDtype>(v[k + 1]));
    }
  } else {
    for (int k = 0; k < 8 && j0 + k < n; ++k) row[k] = cast_out<kDtype>(v[k]);
  }
}

// Short frames (n <= 16 symbols, n even: the class decoders, L <= 8): one thread = one frame, the same
// noise generation as the fused pipeline — 2 Philox blocks and 7-8 Box-Muller pairs per thread
// with no per-block index arithmetic.
// A warp's 32 rows are contiguous in y, so its 4- / 8-byte stores fill whole sectors between them.
// kN: compile-time symbol count (14 for the reference L = 7: the unused Box-Muller pair and every bound check
// disappear), 0 = run-time n
template <int kDtype, int kN>
__global__ void __launch_bounds__(256) channel_frame_kernel(const PhiloxKeys rk, const uint8_t* __restrict__ coded,
                                                            int n_rt, int64_t B, float sigma, uint32_t stream_word,
                                                            uint64_t frame0, typename OutT<kDtype>::type* __restrict__ y) {
  using T = typename OutT<kDtype>::type;
  const int n = kN ? kN : n_rt;
  // a warp's 32 rows go through a shared staging block so that they leave as full 16-byte vectors: per-thread
  // 4-byte stores at a 28-byte stride cost seven partial-sector writes per sector in L2
  __shared__ __align__(16) T stage[8][32 * 16];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t f = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t f_warp = f - lane;
  if (f_warp >= B) return;
  const bool live = f < B;
  const int S = (n + 7) >> 3;
  uint32_t bits = live ? __ldg(coded + f * S) : 0u;
  if (live && S > 1) bits |= static_cast<uint32_t>(__ldg(coded + f * S + 1)) << 8;
  T* row = &stage[wib][lane * n];
  float v[16];
  if (kN == 14) {
    frame14_symbols(rk, frame0 + f, stream_word, bm_scale(sigma), bits, v);
  } else {
    block_symbols<4>(rk, frame0 + f, 0u, stream_word, bm_scale(sigma), bits, v);
    if (n > 8) block_symbols<4>(rk, frame0 + f, 1u, stream_word, bm_scale(sigma), bits >> 8, v + 8);
  }
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    if (j >= n) break;
    const float a = v[j], b = v[j + 1];
    if constexpr (kDtype == NVD_BF16) reinterpret_cast<__nv_bfloat162*>(row + j)[0] = __floats2bfloat162_rn(a, b);
    else if constexpr (kDtype == NVD_BF16_TRUNC) reinterpret_cast<__nv_bfloat162*>(row + j)[0] = __floats2bfloat162_rn(truncf(a), truncf(b));
    else if constexpr (kDtype == NVD_F32) reinterpret_cast<float2*>(row + j)[0] = make_float2(a, b);
    else reinterpret_cast<int2*>(row + j)[0] = make_int2(cast_out<kDtype>(a), cast_out<kDtype>(b));
  }
  __syncwarp();
  const int64_t rows = (B - f_warp < 32) ? (B - f_warp) : 32;
  const int bytes = static_cast<int>(rows) * n * static_cast<int>(sizeof(T));
  char* dst = reinterpret_cast<char*>(y + f_warp * n);     // 32 * n * sizeof(T) is a multiple of 16: vector-aligned
  const char* src = reinterpret_cast<const char*>(&stage[wib][0]);
  for (int o = lane * 16; o + 16 <= bytes; o += 32 * 16) *reinterpret_cast<uint4*>(dst + o) = *reinterpret_cast<const uint4*>(src + o);
  for (int o = (bytes & ~15) + lane * 2; o + 2 <= bytes; o += 64)   // tail of a partial last warp, 2-byte granules
    *reinterpret_cast<uint16_t*>(dst + o) = *reinterpret_cast<const uint16_t*>(src + o);
}

int launch_channel(nvd_ctx* ctx, const uint8_t* coded, int n, int64_t B, float sigma, int snr_idx,
                   uint64_t frame0, int out_dtype, void* y, int64_t row_stride) {
  if (B == 0) return NVD_OK;
  if (row_stride <= 0) row_stride = n;            // dense rows; a padded stride leaves the pad elements untouched
  const int S = (n + 7) / 8, nblk = (n + kNormalsPerBlock - 1) / kNormalsPerBlock;
  const int64_t threads = B * nblk;
  const unsigned grid = static_cast<unsigned>((threads + 255) / 256);
  const uint32_t sw = noise_stream_word(snr_idx);
  const PhiloxKeys rk = make_philox_keys(ctx->seed);
  if (n <= 16 && (n & 1) == 0 && row_stride == n && reinterpret_cast<uintptr_t>(y) % 16 == 0) {   // the frame kernel stores 16-byte vectors
    const unsigned fgrid = static_cast<unsigned>((B + 255) / 256);
#define NVD_LAUNCH_FRAME(DT, CT)                                                                                          \
  do {                                                                                                                    \
    if (n == 14) channel_frame_kernel<DT, 14><<<fgrid, 256, 0, ctx->stream>>>(rk, coded, n, B, sigma, sw, frame0, static_cast<CT*>(y)); \
    else channel_frame_kernel<DT, 0><<<fgrid, 256, 0, ctx->stream>>>(rk, coded, n, B, sigma, sw, frame0, static_cast<CT*>(y));         \
  } while (0)
    switch (out_dtype) {
      case NVD_F32: NVD_LAUNCH_FRAME(NVD_F32, float); break;
      case NVD_BF16: NVD_LAUNCH_FRAME(NVD_BF16, __nv_bfloat16); break;
      case NVD_I32_TRUNC: NVD_LAUNCH_FRAME(NVD_I32_TRUNC, int32_t); break;
      case NVD_BF16_TRUNC: NVD_LAUNCH_FRAME(NVD_BF16_TRUNC, __nv_bfloat16); break;
      default:
        set_error("nvd_channel: unknown out_dtype %d", out_dtype);
        return NVD_ERR_INVALID;
    }
    NVD_CHECK_LAUNCH();
    return NVD_OK;
  }
  const bool small = threads < (int64_t(1) << 32);
#define NVD_LAUNCH_CHANNEL(DT, CT)                                                                                     \
  do {                                                                                                                 \
    if (small) channel_kernel<DT, true><<<grid, 256, 0, ctx->stream>>>(rk, coded, n, S, nblk, B, sigma, sw, frame0,    \
                                                                       row_stride, static_cast<CT*>(y));                \
    else channel_kernel<DT, false><<<grid, 256, 0, ctx->stream>>>(rk, coded, n, S, nblk, B, sigma, sw, frame0,         \
                                                                  row_stride, static_cast<CT*>(y));                     \
  } while (0)
  switch (out_dtype) {
    case NVD_F32: NVD_LAUNCH_CHANNEL(NVD_F32, float); break;
    case NVD_BF16: NVD_LAUNCH_CHANNEL(NVD_BF16, __nv_bfloat16); break;
    case NVD_I32_TRUNC: NVD_LAUNCH_CHANNEL(NVD_I32_TRUNC, int32_t); break;
    case NVD_BF16_TRUNC: NVD_LAUNCH_CHANNEL(NVD_BF16_TRUNC, __nv_bfloat16); break;
    default:
      set_error("nvd_channel: unknown out_dtype %d", out_dtype);
      return NVD_ERR_INVALID;
  }
#undef NVD_LAUNCH_CHANNEL
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

}  // namespace nvd
