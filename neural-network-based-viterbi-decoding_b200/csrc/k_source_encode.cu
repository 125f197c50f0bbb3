// K0 (source) and K1 (bit-packed convolutional encoder).
//
// K1 replaces the per-frame Python loops of encode_57 / encode_133171
// (/root/reference/src/fnn_viterbi.py:122-217) and, once the host has prepended the six state
// bits, encode_133171_with_state (/root/reference/src/fnn_viterbi_bitwise.py:141-215).
// Frames are byte-packed, time-ordered little-endian (include/nvd.h).  The encoder is a
// shift-XOR on packed words; the 6 bits of state that cross a byte/word boundary travel by
// warp shuffle from the lane that holds the previous bytes (vector path) — lane 0 alone goes
// back to global memory for them.
//
// HBM-bound integer work: algorithmic traffic ceil(L/8) B read + ceil(2L/8) B written per frame
// (3 B at L=7).  Vector path: 16 B loaded / 32 B stored per thread, fully coalesced.
#include "common.cuh"
#include "frame_gen.cuh"

namespace nvd {

// ---------------------------------------------------------------------------------------------
// K0: info words.  L <= 16: frame f carries class (f mod 2^L), stored time-packed.
//     L  > 16: Philox stream kStreamInfo, 128 bits per block.
// ---------------------------------------------------------------------------------------------
// L <= 8, 16 frames (one 16-byte store) per thread
__global__ void __launch_bounds__(256) source_small16_kernel(int L, uint64_t frame0, int64_t n16, uint4* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n16) return;
  const uint32_t mask = (1u << L) - 1u;
  const uint32_t base = static_cast<uint32_t>(frame0 + 16 * static_cast<uint64_t>(i));
  // four frames per 32-bit word: pack the four class indices into bytes, reverse all 32 bits at once (ONE bit-reverse
  // on the XU pipe instead of four: ncu showed this kernel XU-bound at 82-92 %), put the bytes back in order and drop
  // each byte's 8 - L low bits
  uint32_t w[4];
  const uint32_t keep = 0x01010101u * (0xFFu >> (8 - L));
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) c |= ((base + 4 * q + k) & mask) << (8 * k);
    w[q] = (__byte_perm(__brev(c), 0u, 0x0123u) >> (8 - L)) & keep;
  }
  out[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(256) source_small_kernel(int L, uint64_t frame0, int64_t B,
                                                           uint8_t* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;  // 4 frames each
  const uint32_t mask = (1u << L) - 1u;
  if (L <= 8) {
    const int64_t f = 4 * i;
    if (f + 3 < B && (B & 3) == 0) {
      uint32_t w = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        w |= bitrev_L(static_cast<uint32_t>(frame0 + f + k) & mask, L) << (8 * k);
      reinterpret_cast<uint32_t*>(out)[i] = w;
    } else {
      for (int k = 0; k < 4 && f + k < B; ++k)
        out[f + k] = static_cast<uint8_t>(bitrev_L(static_cast<uint32_t>(frame0 + f + k) & mask, L));
    }
  } else {  // 9..16 bits: two bytes per frame
    for (int k = 0; k < 4; ++k) {
      const int64_t f = 4 * i + k;
      if (f < B)
        reinterpret_cast<uint16_t*>(out)[f] =
            static_cast<uint16_t>(bitrev_L(static_cast<uint32_t>(frame0 + f) & mask, L));
    }
  }
}

__global__ void __launch_bounds__(256) source_philox_kernel(uint64_t seed, int L, int S, int nblk,
                                                            uint64_t frame0, int64_t B,
                                                            uint8_t* __restrict__ out) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= B * nblk) return;
  const int64_t f = t / nblk;
  const int blk = static_cast<int>(t - f * nblk);
  const uint4 x = frame_words(seed, frame0 + f, blk, kStreamInfo);
  uint8_t* row = out + f * S;
  if ((S & 15) == 0 && 128 * (blk + 1) <= L && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    reinterpret_cast<uint4*>(row)[blk] = x;          // a whole block inside 16-byte aligned rows: one vector store
    return;
  }
  const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int byte = 16 * blk + k;
    if (byte < S) {
      uint32_t v = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu;
      const int valid = L - 8 * byte;
      if (valid < 8) v &= (1u << valid) - 1u;
      row[byte] = static_cast<uint8_t>(v);
    }
  }
}

int launch_source(nvd_ctx* ctx, int L, uint64_t frame0, int64_t B, uint8_t* info) {
  if (B == 0) return NVD_OK;
  if (L <= 8 && reinterpret_cast<uintptr_t>(info) % 16 == 0 && B >= 16) {
    const int64_t n16 = B / 16;
    source_small16_kernel<<<static_cast<unsigned>((n16 + 255) / 256), 256, 0, ctx->stream>>>(L, frame0, n16, reinterpret_cast<uint4*>(info));
    NVD_CHECK_LAUNCH();
    const int64_t rest = B - 16 * n16;     // tail frames: the scalar kernel, 4-byte aligned start
    if (rest > 0)
      source_small_kernel<<<static_cast<unsigned>(((rest + 3) / 4 + 255) / 256), 256, 0, ctx->stream>>>(L, frame0 + 16 * n16, rest, info + 16 * n16);
  } else if (L <= 16) {
    const int64_t threads = (B + 3) / 4;
    source_small_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, ctx->stream>>>(
        L, frame0, B, info);
  } else {
    const int S = (L + 7) / 8, nblk = (L + 127) / 128;
    const int64_t threads = B * nblk;
    source_philox_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->seed, L, S, nblk, frame0, B, info);
  }
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

// n_prefix (<= 8) time-packed bits in front of every frame: the six start-state bits that encode_133171_with_state
// prepends (/root/reference/src/fnn_viterbi_bitwise.py:162-163).  One output byte per thread.
__global__ void __launch_bounds__(256) prepend_bits_kernel(const uint8_t* __restrict__ in, int L, int S_in, int S_out,
                                                           uint32_t prefix_bits, int n_prefix, int64_t B, uint8_t* __restrict__ out) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= B * S_out) return;
  const int64_t f = t / S_out;
  const int b = static_cast<int>(t - f * S_out);
  const uint8_t* row = in + f * S_in;
  auto info_byte = [&](int i) -> uint32_t {       // masked to the L valid bits
    if (i < 0 || i >= S_in) return 0u;
    uint32_t v = row[i];
    const int valid = L - 8 * i;
    return valid < 8 ? (v & ((1u << valid) - 1u)) : v;
  };
  uint32_t v;
  if (b == 0) v = prefix_bits | (info_byte(0) << n_prefix);
  else v = ((info_byte(b - 1) | (info_byte(b) << 8)) >> (8 - n_prefix));
  out[t] = static_cast<uint8_t>(v);
}

int launch_prepend_bits(nvd_ctx* ctx, const uint8_t* info, int L, uint32_t prefix_bits, int n_prefix, int64_t B, uint8_t* out) {
  if (B == 0) return NVD_OK;
  const int S_in = (L + 7) / 8, S_out = (L + n_prefix + 7) / 8;
  const int64_t threads = B * S_out;
  prepend_bits_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, ctx->stream>>>(info, L, S_in, S_out, prefix_bits, n_prefix, B, out);
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

// teacher-forced decoder inputs of the recurrent decoder, [B, T_dec] fp32: EOS (1.0), the L info bits as -1/+1, zeros
// (prepare_batch, /root/reference/src/rnn_viterbi.py:362-371)
__global__ void __launch_bounds__(256) teacher_inputs_kernel(const uint8_t* __restrict__ info, int info_stride, int L, int T_dec,
                                                             int64_t B, float* __restrict__ dec_in) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= B * T_dec) return;
  const int64_t f = t / T_dec;
  const int j = static_cast<int>(t - f * T_dec);
  float v = 0.0f;
  if (j == 0) v = 1.0f;
  else if (j <= L) v = ((info[f * info_stride + ((j - 1) >> 3)] >> ((j - 1) & 7)) & 1u) ? 1.0f : -1.0f;
  dec_in[t] = v;
}

int launch_teacher_inputs(nvd_ctx* ctx, const uint8_t* info, int info_stride, int L, int T_dec, int64_t B, float* dec_in) {
  if (B == 0) return NVD_OK;
  const int64_t threads = B * T_dec;
  teacher_inputs_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, ctx->stream>>>(info, info_stride, L, T_dec, B, dec_in);
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

// ---------------------------------------------------------------------------------------------
// K1 vector path: S_out == 2*S_in, 16 input bytes per thread.
// ---------------------------------------------------------------------------------------------
template <int kSinIsOne>
__global__ void __launch_bounds__(256) encode_vec16_kernel(int code, int L, int S_in, int64_t n16,
                                                           const uint4* __restrict__ in,
                                                           uint4* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const bool active = i < n16;
  uint4 w = make_uint4(0, 0, 0, 0);
  if (active) w = __ldg(in + i);
  // previous byte of the stream: top byte of the left neighbour's last word (warp shuffle);
  // lane 0 re-reads it from global memory.
  uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, w.w, 1) >> 24;
  if ((threadIdx.x & 31) == 0)
    prev = (active && i > 0) ? reinterpret_cast<const uint8_t*>(in)[16 * i - 1] : 0u;
  if (!active) return;

  const uint32_t wi[4] = {w.x, w.y, w.z, w.w};
  uint32_t wo[8];
  int b = kSinIsOne ? 0 : static_cast<int>((16 * i) % S_in);  // byte position inside its frame
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    uint32_t u = (wi[k >> 2] >> (8 * (k & 3))) & 0xFFu;
    const int valid = kSinIsOne ? L : min(8, L - 8 * b);
    if (valid < 8) u &= (1u << valid) - 1u;
    const bool start = kSinIsOne ? true : (b == 0);
    const uint32_t c16 = encode16(u, start ? 0u : (prev >> 2), code, start, valid);
    if (k & 1)
      wo[k >> 1] |= c16 << 16;
    else
      wo[k >> 1] = c16;
    prev = (wi[k >> 2] >> (8 * (k & 3))) & 0xFFu;
    if (!kSinIsOne) b = (b + 1 == S_in) ? 0 : b + 1;
  }
  out[2 * i] = make_uint4(wo[0], wo[1], wo[2], wo[3]);
  out[2 * i + 1] = make_uint4(wo[4], wo[5], wo[6], wo[7]);
}

// K1 table path, frames of one byte (L <= 8): the code word is a function of the info byte alone (zero start state), so
// each block builds the 256-entry table with the same encode16 once and every thread translates 16 bytes with 16
// shared-memory lookups — the shift-XOR path spends ~25 instructions per byte and is instruction-bound, not HBM-bound.
__global__ void __launch_bounds__(256) encode_lut16_kernel(int code, int L, int64_t n16, const uint4* __restrict__ in,
                                                           uint4* __restrict__ out) {
  __shared__ uint16_t lut[256];
  {
    const uint32_t u = threadIdx.x & ((L < 8) ? ((1u << L) - 1u) : 0xFFu);
    lut[threadIdx.x] = static_cast<uint16_t>(encode16(u, 0u, code, true, L));
  }
  __syncthreads();
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n16) return;
  const uint4 w = __ldg(in + i);
  const uint32_t wi[4] = {w.x, w.y, w.z, w.w};
  uint32_t wo[8];
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    const uint32_t a = lut[(wi[k >> 2] >> (8 * (k & 3))) & 0xFFu];
    const uint32_t b = lut[(wi[k >> 2] >> (8 * ((k & 3) + 1))) & 0xFFu];
    wo[k >> 1] = a | (b << 16);
  }
  out[2 * i] = make_uint4(wo[0], wo[1], wo[2], wo[3]);
  out[2 * i + 1] = make_uint4(wo[4], wo[5], wo[6], wo[7]);
}

// K1 byte path: any L; one input byte per thread over [g_begin, g_end).
__global__ void __launch_bounds__(256) encode_bytes_kernel(int code, int L, int S_in, int S_out,
                                                           int64_t g_begin, int64_t g_end,
                                                           const uint8_t* __restrict__ in,
                                                           uint8_t* __restrict__ out) {
  const int64_t g = g_begin + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (g >= g_end) return;
  const int64_t f = g / S_in;
  const int b = static_cast<int>(g - f * S_in);
  uint32_t u = in[g];
  const int valid = min(8, L - 8 * b);
  if (valid < 8) u &= (1u << valid) - 1u;
  const uint32_t hist = b ? (static_cast<uint32_t>(in[g - 1]) >> 2) : 0u;
  const uint32_t c16 = encode16(u, hist, code, b == 0, valid);
  uint8_t* o = out + f * S_out + 2 * b;
  o[0] = static_cast<uint8_t>(c16);
  if (2 * b + 1 < S_out) o[1] = static_cast<uint8_t>(c16 >> 8);
}

int launch_encode(nvd_ctx* ctx, int code, const uint8_t* info, int L, int64_t B, uint8_t* coded) {
  if (B == 0) return NVD_OK;
  const int S_in = (L + 7) / 8, S_out = (2 * L + 7) / 8;
  const int64_t total = B * S_in;
  int64_t done = 0;
  const bool aligned = (reinterpret_cast<uintptr_t>(info) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(coded) % 16 == 0);
  if (S_out == 2 * S_in && aligned && total >= 16) {
    const int64_t n16 = total / 16;
    const unsigned grid = static_cast<unsigned>((n16 + 255) / 256);
    if (S_in == 1)
      encode_lut16_kernel<<<grid, 256, 0, ctx->stream>>>(code, L, n16, reinterpret_cast<const uint4*>(info),
                                                         reinterpret_cast<uint4*>(coded));
    else
      encode_vec16_kernel<0><<<grid, 256, 0, ctx->stream>>>(
          code, L, S_in, n16, reinterpret_cast<const uint4*>(info), reinterpret_cast<uint4*>(coded));
    NVD_CHECK_LAUNCH();
    done = n16 * 16;
  }
  if (done < total) {
    const int64_t rest = total - done;
    encode_bytes_kernel<<<static_cast<unsigned>((rest + 255) / 256), 256, 0, ctx->stream>>>(
        code, L, S_in, S_out, done, total, info, coded);
    NVD_CHECK_LAUNCH();
  }
  return NVD_OK;
}

}  // namespace nvd
