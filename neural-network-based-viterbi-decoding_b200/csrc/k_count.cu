// K5: fused bit-error / frame-error popcount reduction.
//
// Replaces calculate_error (/root/reference/src/fnn_viterbi.py:619-674: FER = mean(pred != label),
// BER = sum of per-bit XORs of the MSB-first bit rows) and calculate_ber /
// calculate_hamming_distance_decimal (/root/reference/src/fnn_viterbi_mixed_snr.py:303-328,634-655:
// bin(a ^ b).count('1')), plus the inline counts of fnn_viterbi_bitwise.py:508-509 and
// rnn_viterbi.py:477-491.  Both reduce to sum popcount(pred ^ label) and sum [pred != label].
//
// HBM-bound byte work: 2*stride bytes per frame.  Byte frames (L <= 8) take 16 frames per thread
// as two 128-bit loads; counts are reduced warp -> block -> one 64-bit atomic per counter per CTA.
#include <algorithm>

#include "common.cuh"
#include "frame_gen.cuh"

namespace nvd {

namespace {

__device__ __forceinline__ void block_accumulate(uint32_t be, uint32_t fe, uint64_t* counters,
                                                 uint64_t add_frames, uint64_t add_bits) {
  __shared__ uint32_t sbe[8], sfe[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    be += __shfl_xor_sync(0xFFFFFFFFu, be, o);
    fe += __shfl_xor_sync(0xFFFFFFFFu, fe, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sbe[w] = be; sfe[w] = fe; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tb = 0, tf = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) { tb += sbe[i]; tf += sfe[i]; }
    if (tb) atomicAdd(reinterpret_cast<unsigned long long*>(counters), static_cast<unsigned long long>(tb));
    if (tf) atomicAdd(reinterpret_cast<unsigned long long*>(counters + 1), static_cast<unsigned long long>(tf));
    if (blockIdx.x == 0 && add_frames) {
      atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2), static_cast<unsigned long long>(add_frames));
      atomicAdd(reinterpret_cast<unsigned long long*>(counters + 3), static_cast<unsigned long long>(add_bits));
    }
  }
}

// per-byte: MSB-first class index (L <= 8 bits) -> time-packed byte
__device__ __forceinline__ uint32_t brev_bytes(uint32_t w, int L) {
  const uint32_t r = __byte_perm(__brev(w), 0u, 0x0123u);
  return (r >> (8 - L)) & (0x01010101u * (0xFFu >> (8 - L)));
}

__global__ void __launch_bounds__(256) count_bytes16_kernel(const uint4* __restrict__ pred,
                                                            const uint4* __restrict__ label, int64_t n16,
                                                            int L, int msb, uint64_t* counters,
                                                            uint64_t add_frames, uint64_t add_bits) {
  uint32_t be = 0, fe = 0;
  const uint32_t mask = 0x01010101u * ((1u << L) - 1u);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n16;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 p = __ldg(pred + i), l = __ldg(label + i);
    const uint32_t pw[4] = {p.x, p.y, p.z, p.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t pv = msb ? brev_bytes(pw[k], L) : pw[k];
      const uint32_t x = (pv ^ lw[k]) & mask;
      be += __popc(x);
      fe += __popc(__vcmpne4(x, 0u) & 0x01010101u);
    }
  }
  block_accumulate(be, fe, counters, add_frames, add_bits);
}

__global__ void __launch_bounds__(256) count_generic_kernel(const uint8_t* __restrict__ pred,
                                                            const uint8_t* __restrict__ label, int S, int L,
                                                            int64_t f_begin, int64_t B, int msb,
                                                            uint64_t* counters, uint64_t add_frames,
                                                            uint64_t add_bits) {
  uint32_t be = 0, fe = 0;
  for (int64_t f = f_begin + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; f < B;
       f += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint8_t* p = pred + f * S;
    const uint8_t* l = label + f * S;
    uint32_t any = 0;
    if (msb) {  // S <= 4: little-endian integers
      uint32_t pv = 0, lv = 0;
      for (int b = 0; b < S; ++b) { pv |= static_cast<uint32_t>(p[b]) << (8 * b); lv |= static_cast<uint32_t>(l[b]) << (8 * b); }
      const uint32_t m = (L >= 32) ? 0xFFFFFFFFu : ((1u << L) - 1u);
      const uint32_t x = (bitrev_L(pv & m, L) ^ lv) & m;
      be += __popc(x);
      any = x;
    } else {
      for (int b = 0; b < S; ++b) {
        uint32_t x = static_cast<uint32_t>(p[b] ^ l[b]);
        const int valid = L - 8 * b;
        if (valid < 8) x &= (1u << valid) - 1u;
        be += __popc(x);
        any |= x;
      }
    }
    fe += any ? 1u : 0u;
  }
  block_accumulate(be, fe, counters, add_frames, add_bits);
}

// decision byte (0 / 1) against ONE bit of a packed info row: the per-bit predictor's test
// (/root/reference/src/fnn_viterbi_bitwise.py:456-457 label, :504-509 decision and error rate)
__global__ void __launch_bounds__(256) count_bit_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ info,
                                                        int info_stride, int bit_position, int64_t B, uint64_t* counters) {
  uint32_t e = 0;
  for (int64_t f = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; f < B; f += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint32_t lab = (static_cast<uint32_t>(__ldg(info + f * info_stride + (bit_position >> 3))) >> (bit_position & 7)) & 1u;
    e += (static_cast<uint32_t>(__ldg(pred + f)) ^ lab) & 1u;
  }
  block_accumulate(e, e, counters, static_cast<uint64_t>(B), static_cast<uint64_t>(B));
}

// packed rows with different strides (the recurrent decoder writes T_dec decision bits per frame, the label row holds
// the L info bits; only the first L are scored, /root/reference/src/rnn_viterbi.py:477-491)
__global__ void __launch_bounds__(256) count_rows_kernel(const uint8_t* __restrict__ pred, int pred_stride,
                                                         const uint8_t* __restrict__ label, int label_stride, int L, int64_t B,
                                                         uint64_t* counters) {
  uint32_t be = 0, fe = 0;
  const int S = (L + 7) >> 3;
  for (int64_t f = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; f < B; f += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint8_t* p = pred + f * pred_stride;
    const uint8_t* l = label + f * label_stride;
    uint32_t any = 0;
    for (int b = 0; b < S; ++b) {
      uint32_t x = static_cast<uint32_t>(p[b] ^ l[b]);
      const int valid = L - 8 * b;
      if (valid < 8) x &= (1u << valid) - 1u;
      be += __popc(x);
      any |= x;
    }
    fe += any ? 1u : 0u;
  }
  block_accumulate(be, fe, counters, static_cast<uint64_t>(B), static_cast<uint64_t>(B) * static_cast<uint64_t>(L));
}

}  // namespace

int launch_count_bit(nvd_ctx* ctx, const uint8_t* pred, const uint8_t* info, int info_stride, int bit_position, int64_t B,
                     uint64_t* counters) {
  if (B == 0) return NVD_OK;
  const int grid = static_cast<int>(std::min<int64_t>((B + 255) / 256, ctx->sm_count * 8));
  count_bit_kernel<<<grid, 256, 0, ctx->stream>>>(pred, info, info_stride, bit_position, B, counters);
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

int launch_count_rows(nvd_ctx* ctx, const uint8_t* pred, int pred_stride, const uint8_t* label, int label_stride, int L, int64_t B,
                      uint64_t* counters) {
  if (B == 0) return NVD_OK;
  const int grid = static_cast<int>(std::min<int64_t>((B + 255) / 256, ctx->sm_count * 8));
  count_rows_kernel<<<grid, 256, 0, ctx->stream>>>(pred, pred_stride, label, label_stride, L, B, counters);
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

int launch_count(nvd_ctx* ctx, const void* pred, const void* label, int stride, int L, int64_t B, int flags,
                 uint64_t* counters) {
  if (B == 0) return NVD_OK;
  const int msb = (flags & NVD_COUNT_PRED_IS_MSB_CLASS) ? 1 : 0;
  int64_t done = 0;
  uint64_t tot_frames = static_cast<uint64_t>(B), tot_bits = static_cast<uint64_t>(B) * static_cast<uint64_t>(L);
  const bool aligned = (reinterpret_cast<uintptr_t>(pred) % 16 == 0) && (reinterpret_cast<uintptr_t>(label) % 16 == 0);
  const int max_grid = ctx->sm_count * 8;
  if (stride == 1 && aligned && B >= 16) {
    const int64_t n16 = B / 16;
    const int grid = static_cast<int>(std::min<int64_t>((n16 + 255) / 256, max_grid));
    count_bytes16_kernel<<<grid, 256, 0, ctx->stream>>>(static_cast<const uint4*>(pred),
                                                        static_cast<const uint4*>(label), n16, L, msb, counters,
                                                        tot_frames, tot_bits);
    tot_frames = tot_bits = 0;
    NVD_CHECK_LAUNCH();
    done = n16 * 16;
  }
  if (done < B) {
    const int grid = static_cast<int>(std::min<int64_t>((B - done + 255) / 256, max_grid));
    count_generic_kernel<<<grid, 256, 0, ctx->stream>>>(static_cast<const uint8_t*>(pred),
                                                        static_cast<const uint8_t*>(label), stride, L, done, B,
                                                        msb, counters, tot_frames, tot_bits);
    NVD_CHECK_LAUNCH();
  }
  return NVD_OK;
}

}  // namespace nvd
